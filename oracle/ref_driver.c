/* ref_driver.c -- drives the UNMODIFIED reference libpgsd (compiled in place from
 * /root/reference/pgsd/pgsd/pgsd.c against oracle/shim/mpi.h) at P shim ranks.
 *
 * TEST INFRASTRUCTURE ONLY.  Built into oracle/_ref/ref_driver by oracle/build_ref.sh.
 * Used (a) to generate the golden .gsd files the parity tests compare against and
 * (b) as the CPU "reference" arm of bench.py.  The product never links or runs it.
 *
 * Two modes:
 *   ref_driver script <ops.txt> <blob.bin> <out_prefix>
 *   ref_driver bench  <file.gsd> <N> <frames> <blob.bin> [fsync]
 *
 * Script ops (one per line, whitespace separated), see tests/opscript.py for the
 * generator and pgsd_sph_b200/replay.py for the product-side interpreter of the same file:
 *   create <path> <application> <schema> <schema_version> <flags> <excl>
 *   open <path> <flags>
 *   setbuf <bytes> | setidx <entries>
 *   chunk <name> <type> <M> <all> <R|S|X> <N_global> <blob_off> [rows_0 .. rows_{P-1}]
 *   end_frame | flush | close | nframes | nnames
 *   find <frame> <name>
 *   read <frame> <name> <all>
 *   match <prefix|->           ("-" = empty prefix)
 * Call conventions follow SURVEY.md Appendix A.4 (taken from the disabled Python
 * writer hoomd.py:597-632 and benchmark-write.cc:33-45,85-130):
 *   R: every rank passes the whole array (N=N_global, offset=0, global_size=0)
 *   S: rows split floor(N/P) (+1 if rank < N%P), offset = row_start*M elements
 *   X: explicit rows per rank.
 */
#define _GNU_SOURCE
#include "pgsd.h"

#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

static int g_rank, g_np;

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
    void* p = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_SHARED, fd, 0);
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
    FILE* log = (g_rank == 0) ? fopen(logpath, "w") : NULL;

    struct pgsd_handle h;
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
            int rc = pgsd_create_and_open(&h, path, app, schema, sv, (enum pgsd_open_flag)flags, excl);
            if (log)
                fprintf(log, "%d create %d\n", opno, rc);
            }
        else if (!strcmp(cmd, "open"))
            {
            char path[4096];
            int flags;
            sscanf(rest, "%4095s %d", path, &flags);
            int rc = pgsd_open(&h, path, (enum pgsd_open_flag)flags);
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
            sscanf(rest, "%4095s %d %u %d %7s %llu %llu%n", name, &type, &M, &all, mode, &Ng, &boff,
                   &used);
            rest += used;
            uint64_t rows[64];
            uint64_t N, offset_elems, global_size;
            size_t es = pgsd_sizeof_type((enum pgsd_type)type);
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
                offset_elems = start * M;
                global_size = Ng * M;
                src += start * M * es;
                }
            const void* data = (N == 0) ? NULL : (const void*)src;
            int rc = pgsd_write_chunk(&h, name, (enum pgsd_type)type, N, M, Ng, M, offset_elems,
                                      global_size, all != 0, 0, data);
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
            const struct pgsd_index_entry* e = pgsd_find_chunk(&h, frame, name);
            if (log)
                {
                if (e)
                    fprintf(log, "%d find 1 %llu %llu %lld %u %u %u %u\n", opno,
                            (unsigned long long)e->frame, (unsigned long long)e->N,
                            (long long)e->location, e->M, (unsigned)e->id, (unsigned)e->type,
                            (unsigned)e->flags);
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
            const struct pgsd_index_entry* e = pgsd_find_chunk(&h, frame, name);
            /* only root may dereference the entry (reference quirk Q16, pgsd.c:2378);
               broadcast what the other ranks need, as benchmark-read.cc:89-99 does */
            unsigned long long meta[4] = { 0, 0, 0, 0 };
            if (g_rank == 0 && e)
                {
                meta[0] = 1;
                meta[1] = e->N;
                meta[2] = e->M;
                meta[3] = e->type;
                }
            MPI_Bcast(meta, 4, MPI_UNSIGNED_LONG_LONG, 0, MPI_COMM_WORLD);
            int k = nread++;
            if (!meta[0])
                {
                if (log)
                    fprintf(log, "%d read notfound\n", opno);
                continue;
                }
            size_t es = pgsd_sizeof_type((enum pgsd_type)meta[3]);
            if (!all)
                {
                size_t bytes = meta[1] * meta[2] * es;
                void* buf = malloc(bytes ? bytes : 1);
                int rc = pgsd_read_chunk(&h, buf, e, 0, 0, 0, false);
                if (g_rank == 0)
                    dump_bytes(out_prefix, k, -1, buf, rc == 0 ? bytes : 0);
                if (log)
                    fprintf(log, "%d read %d %zu\n", opno, rc, bytes);
                free(buf);
                }
            else
                {
                uint64_t rows[64];
                split_rows(meta[1], rows);
                uint64_t start = 0;
                for (int r = 0; r < g_rank; r++)
                    start += rows[r];
                size_t bytes = rows[g_rank] * meta[2] * es;
                void* buf = malloc(bytes ? bytes : 1);
                int rc = pgsd_read_chunk(&h, buf, e, rows[g_rank], (uint32_t)meta[2],
                                         (uint32_t)start, true);
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
            /* Non-root ranks hold no namelist (pgsd.c:1531-1607) and would dereference NULL
               (pgsd.c:2590), so only root may call; on a writable file the call flushes
               collectively (pgsd.c:2579-2586), which root cannot do alone: unsupported at P > 1. */
            if (g_np > 1 && h.open_flags != PGSD_OPEN_READONLY)
                {
                if (log)
                    fprintf(log, "%d match unsupported\n", opno);
                continue;
                }
            const char* found = g_rank == 0 ? pgsd_find_matching_chunk_name(&h, m, NULL) : NULL;
            if (log)
                fprintf(log, "%d match", opno);
            while (g_rank == 0 && found)
                {
                fprintf(log, " %s", found);
                found = pgsd_find_matching_chunk_name(&h, m, found);
                }
            /* non-root ranks hold no namelist (pgsd.c:1531-1607): only root iterates */
            if (log)
                fprintf(log, "\n");
            }
        else
            {
            if (g_rank == 0)
                fprintf(stderr, "ref_driver: unknown op '%s'\n", cmd);
            return 2;
            }
        }
    if (log)
        fclose(log);
    fclose(ops);
    return 0;
    }

/* bench mode: the HOOMD-schema frame of SURVEY.md section 8(d) written `frames` times.
   Input blob = SoA f32/u32 arrays of length N in the order
   pos_x pos_y pos_z vel_x vel_y vel_z density pressure typeid id. */
static int run_bench(const char* path, uint64_t N, int frames, const char* blob_path, int do_fsync, int nlogs)
    {
    size_t blob_len = 0;
    const unsigned char* blob = map_blob(blob_path, &blob_len);
    if (blob_len < N * 40)
        {
        fprintf(stderr, "ref_driver: blob too small\n");
        return 2;
        }
    const float* soa[8];
    for (int i = 0; i < 8; i++)
        soa[i] = (const float*)(blob + (size_t)i * N * 4);
    const uint32_t* typeid_g = (const uint32_t*)(blob + (size_t)8 * N * 4);
    const uint32_t* id_g = (const uint32_t*)(blob + (size_t)9 * N * 4);

    uint64_t rows[64];
    split_rows(N, rows);
    uint64_t start = 0;
    for (int r = 0; r < g_rank; r++)
        start += rows[r];
    uint64_t n = rows[g_rank];

    float* position = (float*)malloc(n * 12 + 16);
    float* velocity = (float*)malloc(n * 12 + 16);
    double* t_frame = (double*)calloc((size_t)frames, sizeof(double));

    struct pgsd_handle h;
    int rc = pgsd_create_and_open(&h, path, "pgsd-b200", "hoomd", pgsd_make_version(1, 4),
                                  PGSD_OPEN_READWRITE, 0);
    if (rc != 0)
        {
        fprintf(stderr, "ref_driver: create failed %d\n", rc);
        return 2;
        }
    const float box[6] = { 10.f, 10.f, 10.f, 0.f, 0.f, 0.f };
    for (int f = 0; f < frames; f++)
        {
        MPI_Barrier(MPI_COMM_WORLD);
        double t0 = shim_wtime();
        /* the reference's host "contiguity copy" (fl.pyx:571, hoomd.py:220-266): SoA -> (n,3) */
        for (uint64_t i = 0; i < n; i++)
            {
            position[3 * i + 0] = soa[0][start + i];
            position[3 * i + 1] = soa[1][start + i];
            position[3 * i + 2] = soa[2][start + i];
            velocity[3 * i + 0] = soa[3][start + i];
            velocity[3 * i + 1] = soa[4][start + i];
            velocity[3 * i + 2] = soa[5][start + i];
            }
        uint64_t step = 10ull * (uint64_t)f;
        uint8_t dim = 3;
        uint32_t Nu = (uint32_t)N;
        pgsd_write_chunk(&h, "configuration/step", PGSD_TYPE_UINT64, 1, 1, 1, 1, 0, 0, false, 0, &step);
        pgsd_write_chunk(&h, "configuration/dimensions", PGSD_TYPE_UINT8, 1, 1, 1, 1, 0, 0, false, 0, &dim);
        pgsd_write_chunk(&h, "configuration/box", PGSD_TYPE_FLOAT, 6, 1, 6, 1, 0, 0, false, 0, box);
        pgsd_write_chunk(&h, "particles/N", PGSD_TYPE_UINT32, 1, 1, 1, 1, 0, 0, false, 0, &Nu);
        pgsd_write_chunk(&h, "particles/position", PGSD_TYPE_FLOAT, n, 3, N, 3, start * 3, N * 3, true, 0, position);
        pgsd_write_chunk(&h, "particles/velocity", PGSD_TYPE_FLOAT, n, 3, N, 3, start * 3, N * 3, true, 0, velocity);
        pgsd_write_chunk(&h, "particles/typeid", PGSD_TYPE_UINT32, n, 1, N, 1, start, N, true, 0, typeid_g + start);
        pgsd_write_chunk(&h, "particles/density", PGSD_TYPE_FLOAT, n, 1, N, 1, start, N, true, 0, soa[6] + start);
        pgsd_write_chunk(&h, "particles/pressure", PGSD_TYPE_FLOAT, n, 1, N, 1, start, N, true, 0, soa[7] + start);
        pgsd_write_chunk(&h, "log/particles/id", PGSD_TYPE_UINT32, n, 1, N, 1, start, N, true, 0, id_g + start);
        for (int k = 0; k < nlogs; k++) /* BASELINE config 5: per-frame log scalars, root-owned, buffered */
            {
            char nm[64];
            snprintf(nm, sizeof(nm), "log/value/v%d", k);
            float v = (float)k;
            pgsd_write_chunk(&h, nm, PGSD_TYPE_FLOAT, 1, 1, 1, 1, 0, 0, false, 0, &v);
            }
        pgsd_end_frame(&h);
        if (do_fsync)
            {
            /* the shim's MPI_File is {int fd; ...}: first member */
            fsync(*(int*)h.fh);
            }
        MPI_Barrier(MPI_COMM_WORLD);
        t_frame[f] = shim_wtime() - t0;
        }
    pgsd_close(&h);
    if (g_rank == 0)
        {
        printf("{\"ranks\": %d, \"N\": %llu, \"frames\": %d, \"fsync\": %d, \"frame_s\": [", g_np,
               (unsigned long long)N, frames, do_fsync);
        for (int f = 0; f < frames; f++)
            printf("%s%.6f", f ? ", " : "", t_frame[f]);
        printf("]}\n");
        fflush(stdout);
        }
    free(position);
    free(velocity);
    free(t_frame);
    return 0;
    }

int main(int argc, char** argv)
    {
    if (argc < 2)
        {
        fprintf(stderr, "usage: ref_driver script <ops> <blob> <out_prefix> | bench <file> <N> <frames> <blob> [fsync|nofsync [n_log_scalars]]\n");
        return 2;
        }
    MPI_Init(NULL, NULL);
    MPI_Comm_rank(MPI_COMM_WORLD, &g_rank);
    MPI_Comm_size(MPI_COMM_WORLD, &g_np);
    int rc = 2;
    if (!strcmp(argv[1], "script") && argc >= 5)
        rc = run_script(argv[2], argv[3], argv[4]);
    else if (!strcmp(argv[1], "bench") && argc >= 6)
        rc = run_bench(argv[2], strtoull(argv[3], NULL, 10), atoi(argv[4]), argv[5],
                       argc >= 7 && !strcmp(argv[6], "fsync"), argc >= 8 ? atoi(argv[7]) : 0);
    MPI_Finalize();
    return rc;
    }
