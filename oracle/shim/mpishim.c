/* mpishim.c -- fork + shared-memory implementation of oracle/shim/mpi.h
 *
 * TEST INFRASTRUCTURE ONLY (see mpi.h).  Not linked into the product.
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

enum { SHIM_MAX_RANKS = 64, SHIM_SCRATCH = 1 << 16 };

struct shim_shared
    {
    pthread_barrier_t barrier;
    int flag; /* rank-0 status word for collective file ops */
    unsigned char scratch[SHIM_SCRATCH];
    };

struct shim_file
    {
    int fd;
    off_t pos; /* individual file pointer (MPI_File_seek/read/write) */
    };

static struct shim_shared* g_sh = NULL;
static int g_rank = 0;
static int g_size = 1;
static pid_t g_children[SHIM_MAX_RANKS];

double shim_wtime(void)
    {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
    }

void* shim_shared_alloc(size_t bytes)
    {
    void* p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    return p == MAP_FAILED ? NULL : p;
    }

static void shim_sync(void)
    {
    if (g_size > 1)
        pthread_barrier_wait(&g_sh->barrier);
    }

int MPI_Init(int* argc, char*** argv)
    {
    (void)argc;
    (void)argv;
    const char* np = getenv("PGSD_SHIM_NP");
    g_size = np ? atoi(np) : 1;
    if (g_size < 1 || g_size > SHIM_MAX_RANKS)
        {
        fprintf(stderr, "mpishim: bad PGSD_SHIM_NP\n");
        exit(2);
        }
    g_sh = (struct shim_shared*)shim_shared_alloc(sizeof(struct shim_shared));
    if (!g_sh)
        {
        perror("mpishim: mmap");
        exit(2);
        }
    pthread_barrierattr_t attr;
    pthread_barrierattr_init(&attr);
    pthread_barrierattr_setpshared(&attr, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&g_sh->barrier, &attr, (unsigned)g_size);
    fflush(stdout);
    fflush(stderr);
    g_rank = 0;
    for (int r = 1; r < g_size; r++)
        {
        pid_t pid = fork();
        if (pid < 0)
            {
            perror("mpishim: fork");
            exit(2);
            }
        if (pid == 0)
            {
            g_rank = r;
            break;
            }
        g_children[r] = pid;
        }
    return MPI_SUCCESS;
    }

int MPI_Finalize(void)
    {
    shim_sync();
    fflush(stdout);
    fflush(stderr);
    if (g_rank != 0)
        _exit(0);
    int bad = 0;
    for (int r = 1; r < g_size; r++)
        {
        int st = 0;
        waitpid(g_children[r], &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0)
            bad = 1;
        }
    if (bad)
        fprintf(stderr, "mpishim: a rank exited abnormally\n");
    return MPI_SUCCESS;
    }

int MPI_Comm_rank(MPI_Comm comm, int* rank)
    {
    (void)comm;
    *rank = g_rank;
    return MPI_SUCCESS;
    }

int MPI_Comm_size(MPI_Comm comm, int* size)
    {
    (void)comm;
    *size = g_size;
    return MPI_SUCCESS;
    }

int MPI_Barrier(MPI_Comm comm)
    {
    (void)comm;
    shim_sync();
    return MPI_SUCCESS;
    }

static size_t dt_size(MPI_Datatype dt) { return (size_t)(dt & 0xff); }
static int dt_kind(MPI_Datatype dt) { return (dt >> 8) & 0xff; }

int MPI_Bcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm comm)
    {
    (void)comm;
    size_t bytes = (size_t)count * dt_size(dt);
    if (g_size == 1)
        return MPI_SUCCESS;
    if (bytes > SHIM_SCRATCH)
        {
        fprintf(stderr, "mpishim: Bcast too large\n");
        abort();
        }
    if (g_rank == root)
        memcpy(g_sh->scratch, buf, bytes);
    shim_sync();
    if (g_rank != root)
        memcpy(buf, g_sh->scratch, bytes);
    shim_sync();
    return MPI_SUCCESS;
    }

int MPI_Allgather(const void* send, int scount, MPI_Datatype sdt, void* recv, int rcount,
                  MPI_Datatype rdt, MPI_Comm comm)
    {
    (void)comm;
    (void)rcount;
    (void)rdt;
    size_t bytes = (size_t)scount * dt_size(sdt);
    if (bytes * (size_t)g_size > SHIM_SCRATCH)
        {
        fprintf(stderr, "mpishim: Allgather too large\n");
        abort();
        }
    if (g_size == 1)
        {
        if (send != MPI_IN_PLACE)
            memcpy(recv, send, bytes);
        return MPI_SUCCESS;
        }
    const void* src = (send == MPI_IN_PLACE) ? (const char*)recv + bytes * (size_t)g_rank : send;
    memcpy(g_sh->scratch + bytes * (size_t)g_rank, src, bytes);
    shim_sync();
    memcpy(recv, g_sh->scratch, bytes * (size_t)g_size);
    shim_sync();
    return MPI_SUCCESS;
    }

#define SHIM_REDUCE(T)                                                      \
    {                                                                       \
    T* out = (T*)recv;                                                      \
    const T* all = (const T*)tmp;                                           \
    for (int i = 0; i < count; i++)                                         \
        {                                                                   \
        T acc = all[i];                                                     \
        for (int r = 1; r < g_size; r++)                                    \
            {                                                               \
            T v = all[(size_t)r * (size_t)count + (size_t)i];               \
            if (op == MPI_SUM)                                              \
                acc = (T)(acc + v);                                         \
            else if (op == MPI_MIN)                                         \
                acc = v < acc ? v : acc;                                    \
            else                                                            \
                acc = v > acc ? v : acc;                                    \
            }                                                               \
        out[i] = acc;                                                       \
        }                                                                   \
    }

int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype dt, MPI_Op op,
                  MPI_Comm comm)
    {
    size_t es = dt_size(dt);
    size_t bytes = (size_t)count * es;
    unsigned char tmp[SHIM_SCRATCH / 4];
    if (bytes * (size_t)g_size > sizeof(tmp))
        {
        fprintf(stderr, "mpishim: Allreduce too large\n");
        abort();
        }
    const void* src = (send == MPI_IN_PLACE) ? recv : send;
    unsigned char mine[256];
    if (bytes > sizeof(mine))
        {
        fprintf(stderr, "mpishim: Allreduce too large\n");
        abort();
        }
    memcpy(mine, src, bytes);
    MPI_Allgather(mine, count, dt, tmp, count, dt, comm);
    int kind = dt_kind(dt);
    /* reduce in the DECLARED type: the reference relies on unsigned wrap-around
       for its {s,-s} MIN trick (pgsd.c:174-202). */
    if (kind == 0 && es == 1)
        SHIM_REDUCE(uint8_t)
    else if (kind == 0 && es == 2)
        SHIM_REDUCE(uint16_t)
    else if (kind == 0 && es == 4)
        SHIM_REDUCE(uint32_t)
    else if (kind == 0 && es == 8)
        SHIM_REDUCE(uint64_t)
    else if (kind == 1 && es == 4)
        SHIM_REDUCE(int32_t)
    else if (kind == 1 && es == 8)
        SHIM_REDUCE(int64_t)
    else if (kind == 2 && es == 8)
        SHIM_REDUCE(double)
    else
        {
        fprintf(stderr, "mpishim: Allreduce datatype %d unsupported\n", dt);
        abort();
        }
    return MPI_SUCCESS;
    }

/* only reached from pgsd_bcast_index_entry (pgsd.c:152-172), which nothing on the path calls */
int MPI_Type_create_struct(int n, const int* blocklens, const MPI_Aint* displs,
                           const MPI_Datatype* types, MPI_Datatype* newtype)
    {
    (void)n;
    (void)blocklens;
    (void)displs;
    (void)types;
    *newtype = SHIM_DT(0, 32);
    return MPI_SUCCESS;
    }
int MPI_Type_commit(MPI_Datatype* dt)
    {
    (void)dt;
    return MPI_SUCCESS;
    }
int MPI_Type_free(MPI_Datatype* dt)
    {
    (void)dt;
    return MPI_SUCCESS;
    }

int MPI_File_open(MPI_Comm comm, const char* fname, int amode, MPI_Info info, MPI_File* fh)
    {
    (void)comm;
    (void)info;
    *fh = NULL;
    int base = (amode & MPI_MODE_RDWR) ? O_RDWR : O_RDONLY;
    int fd = -1;
    if (g_rank == 0)
        {
        int fl = base;
        if (amode & MPI_MODE_CREATE)
            fl |= O_CREAT;
        if (amode & MPI_MODE_EXCL)
            fl |= O_EXCL;
        fd = open(fname, fl, 0644);
        g_sh->flag = (fd >= 0) ? 1 : 0;
        }
    shim_sync();
    int ok = g_sh->flag;
    if (ok && g_rank != 0)
        fd = open(fname, base, 0644);
    shim_sync();
    if (!ok || fd < 0)
        return MPI_ERR_OTHER;
    struct shim_file* f = (struct shim_file*)calloc(1, sizeof(struct shim_file));
    f->fd = fd;
    f->pos = 0;
    *fh = f;
    return MPI_SUCCESS;
    }

int MPI_File_close(MPI_File* fh)
    {
    if (fh == NULL || *fh == NULL)
        return MPI_ERR_OTHER;
    int rc = close((*fh)->fd);
    free(*fh);
    *fh = NULL;
    return rc == 0 ? MPI_SUCCESS : MPI_ERR_OTHER;
    }

int MPI_File_set_size(MPI_File fh, MPI_Offset size)
    {
    if (fh == NULL)
        return MPI_ERR_OTHER;
    int rc = 0;
    if (g_rank == 0)
        rc = ftruncate(fh->fd, (off_t)size);
    shim_sync();
    return rc == 0 ? MPI_SUCCESS : MPI_ERR_OTHER;
    }

int MPI_File_get_size(MPI_File fh, MPI_Offset* size)
    {
    if (fh == NULL)
        return MPI_ERR_OTHER;
    struct stat st;
    if (fstat(fh->fd, &st) != 0)
        return MPI_ERR_OTHER;
    *size = (MPI_Offset)st.st_size;
    return MPI_SUCCESS;
    }

int MPI_File_seek(MPI_File fh, MPI_Offset off, int whence)
    {
    if (fh == NULL)
        return MPI_ERR_OTHER;
    if (whence == MPI_SEEK_SET)
        fh->pos = (off_t)off;
    else if (whence == MPI_SEEK_END)
        {
        struct stat st;
        if (fstat(fh->fd, &st) != 0)
            return MPI_ERR_OTHER;
        fh->pos = st.st_size + (off_t)off;
        }
    else
        fh->pos += (off_t)off;
    return MPI_SUCCESS;
    }

static int shim_pread_all(int fd, void* buf, size_t n, off_t off)
    {
    char* p = (char*)buf;
    while (n > 0)
        {
        ssize_t k = pread(fd, p, n, off);
        if (k < 0)
            {
            if (errno == EINTR)
                continue;
            return -1;
            }
        if (k == 0)
            break; /* short read at EOF, like MPI-IO: not an error */
        p += k;
        off += k;
        n -= (size_t)k;
        }
    return 0;
    }

static int shim_pwrite_all(int fd, const void* buf, size_t n, off_t off)
    {
    const char* p = (const char*)buf;
    while (n > 0)
        {
        ssize_t k = pwrite(fd, p, n, off);
        if (k < 0)
            {
            if (errno == EINTR)
                continue;
            return -1;
            }
        p += k;
        off += k;
        n -= (size_t)k;
        }
    return 0;
    }

/* counts are int on purpose (reference quirk Q11: size_t sizes are narrowed at the call) */
int MPI_File_read_at(MPI_File fh, MPI_Offset off, void* buf, int count, MPI_Datatype dt,
                     MPI_Status* st)
    {
    (void)st;
    if (fh == NULL || count < 0)
        return MPI_ERR_OTHER;
    return shim_pread_all(fh->fd, buf, (size_t)count * dt_size(dt), (off_t)off) == 0
               ? MPI_SUCCESS
               : MPI_ERR_OTHER;
    }

int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void* buf, int count, MPI_Datatype dt,
                      MPI_Status* st)
    {
    (void)st;
    if (fh == NULL || count < 0)
        return MPI_ERR_OTHER;
    return shim_pwrite_all(fh->fd, buf, (size_t)count * dt_size(dt), (off_t)off) == 0
               ? MPI_SUCCESS
               : MPI_ERR_OTHER;
    }

int MPI_File_read(MPI_File fh, void* buf, int count, MPI_Datatype dt, MPI_Status* st)
    {
    if (fh == NULL)
        return MPI_ERR_OTHER;
    int rc = MPI_File_read_at(fh, (MPI_Offset)fh->pos, buf, count, dt, st);
    fh->pos += (off_t)((size_t)count * dt_size(dt));
    return rc;
    }

int MPI_File_write(MPI_File fh, const void* buf, int count, MPI_Datatype dt, MPI_Status* st)
    {
    if (fh == NULL)
        return MPI_ERR_OTHER;
    int rc = MPI_File_write_at(fh, (MPI_Offset)fh->pos, buf, count, dt, st);
    fh->pos += (off_t)((size_t)count * dt_size(dt));
    return rc;
    }
