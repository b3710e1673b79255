// comm.cpp -- host-side transports of the rank communicator (single / host callback / shm).
// The NCCL transport lives in device.cu next to the K2 scan kernel.
#include "comm.h"
#include "device.h"

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <vector>

namespace pgsdb
{
uint64_t g_collectives = 0;

void scan_sizes_host(const uint64_t* g, int P, size_t n, int rank, SizeScan* out)
    {
    for (size_t i = 0; i < n; i++)
        {
        uint64_t excl = 0, total = 0, mx = 0;
        for (int r = 0; r < P; r++)
            {
            uint64_t v = g[(size_t)r * n + i];
            if (r < rank)
                excl += v;
            total += v;
            if (v > mx)
                mx = v;
            }
        out[i].excl = excl;
        out[i].total = total;
        out[i].maxv = mx;
        out[i].first = g[i];
        }
    }

int Comm::allgather_scan(const uint64_t* send, SizeScan* out, size_t n)
    {
    if (n == 0)
        return 0;
    std::vector<uint64_t> g((size_t)nprocs * n);
    int rc = allgather(send, g.data(), n);
    if (rc != 0)
        return rc;
    scan_sizes_host(g.data(), nprocs, n, rank, out);
    return 0;
    }

int Comm::barrier()
    {
    if (nprocs == 1)
        return 0;
    uint64_t one = 1;
    std::vector<uint64_t> g((size_t)nprocs);
    return allgather(&one, g.data(), 1);
    }

namespace
    {
class SingleComm : public Comm
    {
    public:
    int allgather(const uint64_t* send, uint64_t* recv, size_t n) override
        {
        memcpy(recv, send, n * sizeof(uint64_t));
        return 0;
        }
    };

class HostComm : public Comm
    {
    public:
    int (*fn)(void*, const uint64_t*, uint64_t*, size_t) = nullptr;
    void* ctx = nullptr;
    int allgather(const uint64_t* send, uint64_t* recv, size_t n) override
        {
        g_collectives++;
        return fn(ctx, send, recv, n);
        }
    };

// Single-host transport: a POSIX shm segment holding a process-shared barrier and one
// slot of SLOT_WORDS uint64 per rank.
class ShmComm : public Comm
    {
    public:
    enum
        {
        SLOT_WORDS = 4096
        };
    struct Seg
        {
        volatile uint64_t ready;
        pthread_barrier_t barrier;
        uint64_t slots[1];
        };
    Seg* seg = nullptr;
    size_t seg_bytes = 0;
    std::string name;

    ~ShmComm() override
        {
        if (seg)
            {
            pthread_barrier_wait(&seg->barrier);
            munmap((void*)seg, seg_bytes);
            if (rank == 0)
                shm_unlink(name.c_str());
            }
        }

    int allgather(const uint64_t* send, uint64_t* recv, size_t n) override
        {
        g_collectives++;
        size_t done = 0;
        while (done < n)
            {
            size_t k = n - done < (size_t)SLOT_WORDS ? n - done : (size_t)SLOT_WORDS;
            memcpy((void*)(seg->slots + (size_t)rank * SLOT_WORDS), send + done, k * 8);
            pthread_barrier_wait(&seg->barrier);
            for (int r = 0; r < nprocs; r++)
                memcpy(recv + (size_t)r * n + done, (const void*)(seg->slots + (size_t)r * SLOT_WORDS),
                       k * 8);
            pthread_barrier_wait(&seg->barrier);
            done += k;
            }
        return 0;
        }
    };

SingleComm g_single;
Comm* g_comm = &g_single;
    } // namespace

Comm* comm() { return g_comm; }

static int g_comm_users = 0; // open file handles that hold a pointer to the communicator
void comm_acquire() { g_comm_users++; }
void comm_release()
    {
    if (g_comm_users > 0)
        g_comm_users--;
    }
int comm_users() { return g_comm_users; }

bool comm_replace(Comm* c)
    {
    if (g_comm_users > 0)
        {
        // open files keep a pointer to the communicator they were opened under
        if (c != nullptr && c != &g_single)
            delete c;
        return false;
        }
    if (g_comm != &g_single)
        delete g_comm;
    g_comm = c ? c : &g_single;
    return true;
    }

const char* comm_kind_name()
    {
    switch (g_comm->kind)
        {
        case CommKind::Host: return "host";
        case CommKind::Shm: return "shm";
        case CommKind::Nccl: return "nccl";
        default: return "single";
        }
    }

Comm* make_host_comm(int rank, int nprocs, int (*fn)(void*, const uint64_t*, uint64_t*, size_t),
                     void* ctx)
    {
    HostComm* c = new HostComm;
    c->rank = rank;
    c->nprocs = nprocs;
    c->kind = CommKind::Host;
    c->fn = fn;
    c->ctx = ctx;
    return c;
    }

Comm* make_shm_comm(int rank, int nprocs, const char* name, std::string& err)
    {
    size_t bytes = sizeof(ShmComm::Seg) + (size_t)nprocs * ShmComm::SLOT_WORDS * 8;
    int fd = -1;
    if (rank == 0)
        {
        shm_unlink(name);
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0)
            {
            err = std::string("shm_open/ftruncate failed: ") + strerror(errno);
            if (fd >= 0)
                close(fd);
            return nullptr;
            }
        }
    else
        {
        // wait (up to 60 s) for rank 0 to create and size the segment
        for (int i = 0; i < 60000; i++)
            {
            fd = shm_open(name, O_RDWR, 0600);
            if (fd >= 0)
                {
                struct stat st;
                if (fstat(fd, &st) == 0 && (size_t)st.st_size >= bytes)
                    break;
                close(fd);
                fd = -1;
                }
            struct timespec ts = { 0, 1000000 };
            nanosleep(&ts, nullptr);
            }
        if (fd < 0)
            {
            err = "timed out waiting for the shm segment";
            return nullptr;
            }
        }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED)
        {
        err = std::string("mmap failed: ") + strerror(errno);
        return nullptr;
        }
    ShmComm* c = new ShmComm;
    c->rank = rank;
    c->nprocs = nprocs;
    c->kind = CommKind::Shm;
    c->seg = (ShmComm::Seg*)p;
    c->seg_bytes = bytes;
    c->name = name;
    if (rank == 0)
        {
        pthread_barrierattr_t attr;
        pthread_barrierattr_init(&attr);
        pthread_barrierattr_setpshared(&attr, PTHREAD_PROCESS_SHARED);
        pthread_barrier_init(&c->seg->barrier, &attr, (unsigned)nprocs);
        __sync_synchronize();
        c->seg->ready = 0x50475344u;
        }
    else
        {
        for (int i = 0; i < 60000 && c->seg->ready != 0x50475344u; i++)
            {
            struct timespec ts = { 0, 1000000 };
            nanosleep(&ts, nullptr);
            }
        if (c->seg->ready != 0x50475344u)
            {
            err = "timed out waiting for the shm barrier";
            munmap(p, bytes);
            c->seg = nullptr;
            delete c;
            return nullptr;
            }
        }
    pthread_barrier_wait(&c->seg->barrier);
    return c;
    }

// geometry of the distributed reorder (device.h); host-only so that CPU tests and callers can use it
int dist_plan(uint64_t n_global, int nranks, DistPlan* out)
    {
    if (n_global == 0 || n_global >= 0xffffffffull || nranks < 1 || nranks > 8)
        return -2;
    int tg = 0;
    while (tg < 32 && ((n_global - 1) >> tg) != 0)
        tg++;
    int L = 10;                 // SLOT_MIN_BITS
    while (tg - L > 15)         // SLOT_MAX_BUCKET_BITS
        L++;
    if (L > 12)                 // SLOT_MAX_BITS
        return -2;
    out->L = L;
    out->cap = 1u << L;
    out->nbp = 1u << (tg > L ? tg - L : 0);
    out->nb_used = (uint32_t)((n_global + out->cap - 1) / out->cap);
    out->nbr = (out->nb_used + (uint32_t)nranks - 1) / (uint32_t)nranks;
    return 0;
    }
} // namespace pgsdb
