// pgsd_replay -- runs an op script (tests/opscript.py) against libpgsd_b200 at P ranks.
//
// The product-side twin of oracle/ref_driver.c's script mode: same op file, same blob, same log
// and read-dump formats, so a parity test runs both and compares the .gsd files, the logs and
// the dumps byte for byte.  Ranks are forked processes joined by the library's shm communicator.
//
//   pgsd_replay <ops.txt> <blob.bin> <out_prefix> [--np P] [--device] [--soa] [--auto-offset]
//
//   --device       chunk bytes are first copied to device memory; pgsd_write_chunk gets device
//                  pointers (K1 vector copy -> arena -> K3 staging), reads land in device memory
//   --soa          with --device: chunks with 2 <= M <= 8 are split into M device columns and go
//                  through pgsd_b200_write_chunk_soa (K1 interleave)
//   --auto-offset  split chunks pass PGSD_B200_OFFSET_AUTO instead of the caller-side prefix sum
#include "pgsd.h"
#include "pgsd_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>
#include <vector>

static int g_rank = 0, g_np = 1;
static bool g_device = false, g_soa = false, g_auto = false;

static const unsigned char* map_blob(const char* path, size_t* len)
    {
    int fd = open(path, O_RDONLY);
    if (fd < 0)
        {
        perror(path);
        exit(2);
        }
    struct stat st;
    fstat(fd, &st);
    *len = (size_t)st.st_size;
    if (st.st_size == 0)
        {
        close(fd);
        return (const unsigned char*)"";
        }
    void* p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED)
        {
        perror("mmap blob");
        exit(2);
        }
    return (const unsigned char*)p;
    }

static void split_rows(uint64_t n, uint64_t* rows)
    {
    for (int r = 0; r < g_np; r++)
        rows[r] = n / (uint64_t)g_np + ((uint64_t)r < n % (uint64_t)g_np ? 1 : 0);
    }

static void dump_bytes(const char* prefix, int k, int rank_tag, const void* p, size_t n)
    {
    char path[4096];
    if (rank_tag < 0)
        snprintf(path, sizeof(path), "%s.read%d", prefix, k);
    else
        snprintf(path, sizeof(path), "%s.read%d.r%d", prefix, k, rank_tag);
    FILE* f = fopen(path, "wb");
    if (!f)
        {
        perror(path);
        exit(2);
        }
    if (n)
        fwrite(p, 1, n, f);
    fclose(f);
    }

static void die_cuda(const char* what)
    {
    fprintf(stderr, "pgsd_replay rank %d: %s: %s\n", g_rank, what, pgsd_b200_last_error());
    exit(3);
    }

// write one chunk through the selected path
static int write_one(pgsd_handle* h, const char* name, int type, uint64_t N, uint32_t M, uint64_t Ng,
                     uint64_t offset_elems, uint64_t global_size, bool all, const unsigned char* src)
    {
    const size_t es = pgsd_sizeof_type((pgsd_type)type);
    if (!g_device || N == 0 || es == 0)
        return pgsd_write_chunk(h, name, (pgsd_type)type, N, M, Ng, M, offset_elems, global_size, all, 0,
                                N == 0 ? nullptr : src);
    const size_t bytes = N * M * es;
    void* d = nullptr;
    if (pgsd_b200_malloc(&d, bytes + 64) != 0)
        die_cuda("malloc");
    int rc;
    if (g_soa && M >= 2 && M <= 8)
        {
        // de-interleave on the host into M columns, upload, let K1 interleave them again
        std::vector<unsigned char> soa(bytes);
        for (uint64_t i = 0; i < N; i++)
            for (uint32_t j = 0; j < M; j++)
                memcpy(&soa[((size_t)j * N + i) * es], src + ((size_t)i * M + j) * es, es);
        if (pgsd_b200_memcpy(d, soa.data(), bytes, 1) != 0)
            die_cuda("memcpy H2D");
        pgsd_b200_column cols[8];
        for (uint32_t j = 0; j < M; j++)
            {
            cols[j].base = (const char*)d + (size_t)j * N * es;
            cols[j].stride = 1;
            }
        rc = pgsd_b200_write_chunk_soa(h, name, (pgsd_type)type, N, M, Ng, M, offset_elems, all,
                                       (pgsd_type)type, cols);
        }
    else
        {
        if (pgsd_b200_memcpy(d, src, bytes, 1) != 0)
            die_cuda("memcpy H2D");
        rc = pgsd_write_chunk(h, name, (pgsd_type)type, N, M, Ng, M, offset_elems, global_size, all, 0, d);
        }
    // the library consumed `d` in stream order (default stream); freeing synchronises
    pgsd_b200_synchronize();
    pgsd_b200_free(d);
    return rc;
    }

static int read_into(pgsd_handle* h, void* host, const pgsd_index_entry* e, uint64_t N, uint32_t M,
                     uint32_t off, bool all, size_t bytes)
    {
    if (!g_device || bytes == 0)
        return pgsd_read_chunk(h, host, e, N, M, off, all);
    void* d = nullptr;
    if (pgsd_b200_malloc(&d, bytes) != 0)
        die_cuda("malloc");
    int rc = pgsd_read_chunk(h, d, e, N, M, off, all);
    if (rc == 0 && pgsd_b200_memcpy(host, d, bytes, 2) != 0)
        die_cuda("memcpy D2H");
    pgsd_b200_free(d);
    return rc;
    }

static int run_script(const char* ops_path, const char* blob_path, const char* out_prefix)
    {
    size_t blob_len = 0;
    const unsigned char* blob = map_blob(blob_path, &blob_len);
    FILE* ops = fopen(ops_path, "r");
    if (!ops)
        {
        perror(ops_path);
        return 2;
        }
    char logpath[4096];
    snprintf(logpath, sizeof(logpath), "%s.log", out_prefix);
    FILE* log = (g_rank == 0) ? fopen(logpath, "w") : nullptr;

    pgsd_handle h;
    memset(&h, 0, sizeof(h));
    char line[8192];
    int opno = 0, nread = 0;
    while (fgets(line, sizeof(line), ops))
        {
        char cmd[64];
        int off = 0;
        if (sscanf(line, "%63s%n", cmd, &off) != 1 || cmd[0] == '#')
            continue;
        const char* rest = line + off;
        opno++;
        if (!strcmp(cmd, "create"))
            {
            char path[4096], app[256], schema[256];
            unsigned sv;
            int flags, excl;
            sscanf(rest, "%4095s %255s %255s %u %d %d", path, app, schema, &sv, &flags, &excl);
            int rc = pgsd_create_and_open(&h, path, app, schema, sv, (pgsd_open_flag)flags, excl);
            if (log)
                fprintf(log, "%d create %d\n", opno, rc);
            }
        else if (!strcmp(cmd, "open"))
            {
            char path[4096];
            int flags;
            sscanf(rest, "%4095s %d", path, &flags);
            int rc = pgsd_open(&h, path, (pgsd_open_flag)flags);
            if (log)
                fprintf(log, "%d open %d\n", opno, rc);
            }
        else if (!strcmp(cmd, "setbuf"))
            {
            unsigned long long v;
            sscanf(rest, "%llu", &v);
            int rc = pgsd_set_maximum_write_buffer_size(&h, v);
            if (log)
                fprintf(log, "%d setbuf %d %llu\n", opno, rc,
                        (unsigned long long)pgsd_get_maximum_write_buffer_size(&h));
            }
        else if (!strcmp(cmd, "setidx"))
            {
            unsigned long long v;
            sscanf(rest, "%llu", &v);
            int rc = pgsd_set_index_entries_to_buffer(&h, v);
            if (log)
                fprintf(log, "%d setidx %d %llu\n", opno, rc,
                        (unsigned long long)pgsd_get_index_entries_to_buffer(&h));
            }
        else if (!strcmp(cmd, "chunk"))
            {
            char name[4096], mode[8];
            int type, all, used = 0;
            unsigned M;
            unsigned long long Ng, boff;
            sscanf(rest, "%4095s %d %u %d %7s %llu %llu%n", name, &type, &M, &all, mode, &Ng, &boff, &used);
            rest += used;
            uint64_t rows[64];
            uint64_t N, offset_elems, global_size;
            size_t es = pgsd_sizeof_type((pgsd_type)type);
            const unsigned char* src = blob + boff;
            if (mode[0] == 'R')
                {
                N = Ng;
                offset_elems = 0;
                global_size = 0;
                }
            else
                {
                if (mode[0] == 'S')
                    split_rows(Ng, rows);
                else
                    for (int r = 0; r < g_np; r++)
                        {
                        unsigned long long v = 0;
                        int u = 0;
                        sscanf(rest, "%llu%n", &v, &u);
                        rest += u;
                        rows[r] = v;
                        }
                uint64_t start = 0;
                for (int r = 0; r < g_rank; r++)
                    start += rows[r];
                N = rows[g_rank];
                offset_elems = g_auto ? PGSD_B200_OFFSET_AUTO : start * M;
                global_size = Ng * M;
                src += start * M * es;
                }
            int rc = write_one(&h, name, type, N, M, Ng, offset_elems, global_size, all != 0, src);
            if (log)
                fprintf(log, "%d chunk %d\n", opno, rc);
            }
        else if (!strcmp(cmd, "end_frame"))
            {
            int rc = pgsd_end_frame(&h);
            if (log)
                fprintf(log, "%d end_frame %d\n", opno, rc);
            }
        else if (!strcmp(cmd, "flush"))
            {
            int rc = pgsd_flush(&h);
            if (log)
                fprintf(log, "%d flush %d\n", opno, rc);
            }
        else if (!strcmp(cmd, "close"))
            {
            int rc = pgsd_close(&h);
            if (log)
                fprintf(log, "%d close %d\n", opno, rc);
            }
        else if (!strcmp(cmd, "nframes"))
            {
            unsigned long long v = pgsd_get_nframes(&h);
            if (log)
                fprintf(log, "%d nframes %llu\n", opno, v);
            }
        else if (!strcmp(cmd, "nnames"))
            {
            unsigned long long v = pgsd_get_nnames(&h);
            if (log)
                fprintf(log, "%d nnames %llu\n", opno, v);
            }
        else if (!strcmp(cmd, "find"))
            {
            unsigned long long frame;
            char name[4096];
            sscanf(rest, "%llu %4095s", &frame, name);
            const pgsd_index_entry* e = pgsd_find_chunk(&h, frame, name);
            if (log)
                {
                if (e)
                    fprintf(log, "%d find 1 %llu %llu %lld %u %u %u %u\n", opno, (unsigned long long)e->frame,
                            (unsigned long long)e->N, (long long)e->location, e->M, (unsigned)e->id,
                            (unsigned)e->type, (unsigned)e->flags);
                else
                    fprintf(log, "%d find 0\n", opno);
                }
            }
        else if (!strcmp(cmd, "read"))
            {
            unsigned long long frame;
            char name[4096];
            int all;
            sscanf(rest, "%llu %4095s %d", &frame, name, &all);
            // every rank holds the index replica, so every rank may dereference the entry
            const pgsd_index_entry* e = pgsd_find_chunk(&h, frame, name);
            int k = nread++;
            if (!e)
                {
                if (log)
                    fprintf(log, "%d read notfound\n", opno);
                continue;
                }
            const pgsd_index_entry ent = *e;
            size_t es = pgsd_sizeof_type((pgsd_type)ent.type);
            if (!all)
                {
                size_t bytes = ent.N * ent.M * es;
                void* buf = malloc(bytes ? bytes : 1);
                int rc = read_into(&h, buf, &ent, 0, 0, 0, false, bytes);
                if (g_rank == 0)
                    dump_bytes(out_prefix, k, -1, buf, rc == 0 ? bytes : 0);
                if (log)
                    fprintf(log, "%d read %d %zu\n", opno, rc, bytes);
                free(buf);
                }
            else
                {
                uint64_t rows[64];
                split_rows(ent.N, rows);
                uint64_t start = 0;
                for (int r = 0; r < g_rank; r++)
                    start += rows[r];
                size_t bytes = rows[g_rank] * ent.M * es;
                void* buf = malloc(bytes ? bytes : 1);
                int rc = read_into(&h, buf, &ent, rows[g_rank], ent.M, (uint32_t)start, true, bytes);
                dump_bytes(out_prefix, k, g_rank, buf, rc == 0 ? bytes : 0);
                if (log)
                    fprintf(log, "%d read %d %zu\n", opno, rc, bytes);
                free(buf);
                }
            }
        else if (!strcmp(cmd, "match"))
            {
            char prefix[4096];
            sscanf(rest, "%4095s", prefix);
            const char* m = strcmp(prefix, "-") ? prefix : "";
            const char* found = pgsd_find_matching_chunk_name(&h, m, nullptr);
            if (log)
                fprintf(log, "%d match", opno);
            while (found)
                {
                if (log)
                    fprintf(log, " %s", found);
                found = pgsd_find_matching_chunk_name(&h, m, found);
                }
            if (log)
                fprintf(log, "\n");
            }
        else
            {
            if (g_rank == 0)
                fprintf(stderr, "pgsd_replay: unknown op '%s'\n", cmd);
            return 2;
            }
        }
    if (log)
        fclose(log);
    fclose(ops);
    return 0;
    }

int main(int argc, char** argv)
    {
    std::vector<const char*> pos;
    for (int i = 1; i < argc; i++)
        {
        if (!strcmp(argv[i], "--np") && i + 1 < argc)
            g_np = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--device"))
            g_device = true;
        else if (!strcmp(argv[i], "--soa"))
            g_soa = true;
        else if (!strcmp(argv[i], "--auto-offset"))
            g_auto = true;
        else
            pos.push_back(argv[i]);
        }
    if (pos.size() != 3 || g_np < 1 || g_np > 64)
        {
        fprintf(stderr, "usage: pgsd_replay <ops> <blob> <out_prefix> [--np P] [--device] [--soa] [--auto-offset]\n");
        return 2;
        }
    // fork before anything touches CUDA
    std::vector<pid_t> kids;
    for (int r = 1; r < g_np; r++)
        {
        pid_t p = fork();
        if (p < 0)
            {
            perror("fork");
            return 2;
            }
        if (p == 0)
            {
            g_rank = r;
            kids.clear();
            break;
            }
        kids.push_back(p);
        }
    char seg[128];
    snprintf(seg, sizeof(seg), "/pgsd_replay_%d", g_rank == 0 ? (int)getpid() : (int)getppid());
    if (g_np > 1 && pgsd_b200_comm_init_shm(g_rank, g_np, seg) != 0)
        {
        fprintf(stderr, "pgsd_replay rank %d: comm init failed: %s\n", g_rank, pgsd_b200_last_error());
        return 2;
        }
    if (g_device && !pgsd_b200_cuda_available())
        {
        fprintf(stderr, "pgsd_replay: --device needs a CUDA device (no CPU fallback)\n");
        return 3;
        }
    int rc = run_script(pos[0], pos[1], pos[2]);
    pgsd_b200_shutdown();
    pgsd_b200_comm_finalize();
    if (g_rank != 0)
        _exit(rc);
    for (pid_t p : kids)
        {
        int st = 0;
        waitpid(p, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0)
            rc = rc ? rc : 4;
        }
    return rc;
    }
