// api_b200.cpp -- the pgsd_b200.h C ABI: communicator set-up, device-resident SoA chunk writes
// (K1), the size scan (K2), the particle-id reorder (K4 + K5) and accounting.  Thin extern "C"
// shims over comm.cpp / device.cu / kernels_*.cu / pgsd_file.cpp; no CUDA or torch types cross.
#include "../../include/pgsd_b200.h"
#include "comm.h"
#include "device.h"
#include "file_stage.h"
#include "read_ahead.h"

#include <fcntl.h>
#include <unistd.h>
#include "file_internal.h"

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace pgsdb;

static int replace_comm(Comm* c)
    {
    if (comm_replace(c))
        return PGSD_SUCCESS;
    set_last_error("the communicator cannot be replaced while files opened under it are still open");
    return PGSD_ERROR_INVALID_ARGUMENT;
    }

extern "C" {

// ------------------------------------------------------------------ communicator
int pgsd_b200_comm_init_host(int rank, int nprocs, pgsd_b200_allgather_fn fn, void* ctx)
    {
    if (nprocs < 1 || rank < 0 || rank >= nprocs || (nprocs > 1 && fn == nullptr))
        {
        set_last_error("comm_init_host: bad rank/nprocs or missing all-gather callback");
        return PGSD_ERROR_INVALID_ARGUMENT;
        }
    return replace_comm(nprocs == 1 ? nullptr : make_host_comm(rank, nprocs, fn, ctx));
    }

int pgsd_b200_comm_init_shm(int rank, int nprocs, const char* segment_name)
    {
    if (nprocs < 1 || rank < 0 || rank >= nprocs || segment_name == nullptr)
        {
        set_last_error("comm_init_shm: bad rank/nprocs/name");
        return PGSD_ERROR_INVALID_ARGUMENT;
        }
    if (nprocs == 1)
        return replace_comm(nullptr);
    std::string err;
    Comm* c = make_shm_comm(rank, nprocs, segment_name, err);
    if (!c)
        {
        set_last_error(err);
        return PGSD_ERROR_IO;
        }
    return replace_comm(c);
    }

int pgsd_b200_nccl_unique_id(void* out128)
    {
    if (out128 == nullptr)
        return PGSD_ERROR_INVALID_ARGUMENT;
    std::string err;
    if (nccl_unique_id(out128, err) != 0)
        {
        set_last_error(err);
        return PGSD_ERROR_IO;
        }
    return PGSD_SUCCESS;
    }

int pgsd_b200_comm_init_nccl(int rank, int nprocs, const void* unique_id128, int cuda_device)
    {
    if (nprocs < 1 || rank < 0 || rank >= nprocs || unique_id128 == nullptr)
        {
        set_last_error("comm_init_nccl: bad rank/nprocs/id");
        return PGSD_ERROR_INVALID_ARGUMENT;
        }
    std::string err;
    Comm* c = make_nccl_comm(rank, nprocs, unique_id128, cuda_device, err);
    if (!c)
        {
        set_last_error(err);
        return PGSD_ERROR_IO;
        }
    return replace_comm(c);
    }

int pgsd_b200_comm_finalize(void) { return replace_comm(nullptr); }

int pgsd_b200_comm_rank(void) { return comm()->rank; }
int pgsd_b200_comm_size(void) { return comm()->nprocs; }
const char* pgsd_b200_comm_kind(void) { return comm_kind_name(); }
int pgsd_b200_barrier(void) { return comm()->barrier() == 0 ? PGSD_SUCCESS : PGSD_ERROR_IO; }

int pgsd_b200_partition(uint64_t n_local, uint64_t* n_global, uint64_t* row_start)
    {
    SizeScan sc;
    if (comm()->allgather_scan(&n_local, &sc, 1) != 0)
        return PGSD_ERROR_IO;
    if (n_global)
        *n_global = sc.total;
    if (row_start)
        *row_start = sc.excl;
    return PGSD_SUCCESS;
    }

// ------------------------------------------------------------------ device
int pgsd_b200_cuda_available(void) { return dev_cuda_available() ? 1 : 0; }
int pgsd_b200_device_init(int cuda_device) { return dev_init(cuda_device); }
int pgsd_b200_set_stream(void* cuda_stream)
    {
    dev_set_user_stream(cuda_stream);
    return PGSD_SUCCESS;
    }
const char* pgsd_b200_last_error(void) { return last_error().c_str(); }

int pgsd_b200_configure_staging(uint32_t n_slots, uint64_t slot_bytes, uint32_t writer_threads)
    {
    return dev_configure_staging(n_slots, slot_bytes, writer_threads);
    }

static bool columns_on_device(const struct pgsd_b200_column* cols, uint32_t M, bool* mixed)
    {
    int ndev = 0;
    for (uint32_t j = 0; j < M; j++)
        if (dev_is_device_pointer(cols[j].base))
            ndev++;
    *mixed = ndev != 0 && ndev != (int)M;
    return ndev == (int)M;
    }

int pgsd_b200_write_chunk_soa(struct pgsd_handle* handle, const char* name, enum pgsd_type dst_type,
                              uint64_t N, uint32_t M, uint64_t N_global, uint32_t M_global,
                              uint64_t offset, bool all, enum pgsd_type src_type,
                              const struct pgsd_b200_column* cols)
    {
    if (M == 0 || M > 8 || (N > 0 && cols == nullptr))
        return PGSD_ERROR_INVALID_ARGUMENT;
    Column c[8];
    bool on_device = true, mixed = false;
    if (N > 0)
        {
        for (uint32_t j = 0; j < M; j++)
            {
            if (cols[j].base == nullptr)
                return PGSD_ERROR_INVALID_ARGUMENT;
            c[j].base = cols[j].base;
            c[j].stride = cols[j].stride;
            }
        on_device = columns_on_device(cols, M, &mixed);
        if (mixed)
            {
            set_last_error("write_chunk_soa: columns must be all device or all host pointers");
            return PGSD_ERROR_INVALID_ARGUMENT;
            }
        }
    return file_write_chunk_device(handle, name, (int)dst_type, N, M, N_global, M_global, offset, all,
                                   (int)src_type, c, !on_device);
    }

int pgsd_b200_write_chunks_soa(struct pgsd_handle* handle, int n_chunks,
                               const struct pgsd_b200_chunk_desc* chunks)
    {
    if (n_chunks < 0 || (n_chunks > 0 && chunks == nullptr))
        return PGSD_ERROR_INVALID_ARGUMENT;
    std::vector<DeviceChunk> dc((size_t)n_chunks);
    std::vector<Column> cols((size_t)n_chunks * 8);
    for (int i = 0; i < n_chunks; i++)
        {
        const pgsd_b200_chunk_desc& d = chunks[i];
        if (d.M == 0 || d.M > 8 || (d.N > 0 && d.cols == nullptr))
            return PGSD_ERROR_INVALID_ARGUMENT;
        bool on_device = true, mixed = false;
        Column* c = &cols[(size_t)i * 8];
        if (d.N > 0)
            {
            for (uint32_t j = 0; j < d.M; j++)
                {
                if (d.cols[j].base == nullptr)
                    return PGSD_ERROR_INVALID_ARGUMENT;
                c[j].base = d.cols[j].base;
                c[j].stride = d.cols[j].stride;
                }
            on_device = columns_on_device(d.cols, d.M, &mixed);
            if (mixed)
                {
                set_last_error("write_chunks_soa: columns must be all device or all host pointers");
                return PGSD_ERROR_INVALID_ARGUMENT;
                }
            }
        dc[(size_t)i] = DeviceChunk { d.name, (int)d.dst_type, (int)d.src_type, d.N, d.M, d.N_global, d.M_global,
                                      d.offset, d.all, c, !on_device };
        }
    return file_write_chunks_device(handle, n_chunks, dc.data());
    }

int pgsd_b200_pack_soa(void* dst_device, enum pgsd_type dst_type, uint64_t N, uint32_t M,
                       enum pgsd_type src_type, const struct pgsd_b200_column* cols_device,
                       void* cuda_stream)
    {
    if (M == 0 || M > 8 || cols_device == nullptr)
        return PGSD_ERROR_INVALID_ARGUMENT;
    Column c[8];
    for (uint32_t j = 0; j < M; j++)
        {
        c[j].base = cols_device[j].base;
        c[j].stride = cols_device[j].stride;
        }
    return dev_pack(dst_device, (int)dst_type, N, M, (int)src_type, c, cuda_stream);
    }

int pgsd_b200_scan_sizes(const uint64_t* sizes, int P, int C, int rank, uint64_t* excl,
                         uint64_t* total, uint64_t* maxv)
    {
    if (C < 0)
        return PGSD_ERROR_INVALID_ARGUMENT;
    std::vector<SizeScan> out((size_t)C);
    int rc = dev_scan_sizes(sizes, P, C, rank, out.data());
    if (rc != 0)
        return rc;
    for (int c = 0; c < C; c++)
        {
        if (excl)
            excl[c] = out[(size_t)c].excl;
        if (total)
            total[c] = out[(size_t)c].total;
        if (maxv)
            maxv[c] = out[(size_t)c].maxv;
        }
    return PGSD_SUCCESS;
    }

// ------------------------------------------------------------------ reorder
static int to_fields(int nfields, const struct pgsd_b200_field* in, std::vector<ReorderField>& out)
    {
    if (nfields < 0 || (nfields > 0 && in == nullptr))
        return PGSD_ERROR_INVALID_ARGUMENT;
    out.resize((size_t)nfields);
    for (int i = 0; i < nfields; i++)
        {
        out[(size_t)i].in = in[i].in;
        out[(size_t)i].out = in[i].out;
        out[(size_t)i].row_bytes = in[i].row_bytes;
        }
    return PGSD_SUCCESS;
    }

int pgsd_b200_sort_ids(uint64_t n, const uint32_t* keys_device, uint32_t* keys_sorted_device,
                       uint32_t* perm_device, void* cuda_stream)
    {
    return dev_sort_ids(n, keys_device, keys_sorted_device, perm_device, cuda_stream);
    }

int pgsd_b200_gather(uint64_t n, const uint32_t* perm_device, int nfields,
                     const struct pgsd_b200_field* fields_device, void* cuda_stream)
    {
    std::vector<ReorderField> f;
    int rc = to_fields(nfields, fields_device, f);
    if (rc != 0)
        return rc;
    return dev_gather(n, perm_device, nfields, f.data(), cuda_stream);
    }

int pgsd_b200_reorder_device(uint64_t n, const uint32_t* keys_device, uint32_t* keys_sorted_device,
                             uint32_t* perm_device, int nfields,
                             const struct pgsd_b200_field* fields_device, void* cuda_stream)
    {
    std::vector<ReorderField> f;
    int rc = to_fields(nfields, fields_device, f);
    if (rc != 0)
        return rc;
    return dev_reorder(n, keys_device, keys_sorted_device, perm_device, nfields, f.data(), cuda_stream);
    }

int pgsd_b200_reorder_host(uint64_t n, const uint32_t* keys_host, uint32_t* keys_sorted_host,
                           uint32_t* perm_host, int nfields, const struct pgsd_b200_field* fields_host)
    {
    std::vector<ReorderField> f;
    int rc = to_fields(nfields, fields_host, f);
    if (rc != 0)
        return rc;
    return dev_reorder_host(n, keys_host, keys_sorted_host, perm_host, nfields, f.data());
    }

int pgsd_b200_reorder_distributed_plan(uint64_t n_global, int nranks, int rank, uint64_t* id_first, uint64_t* max_rows)
    {
    DistPlan pl;
    if (rank < 0 || rank >= nranks || id_first == nullptr || max_rows == nullptr || dist_plan(n_global, nranks, &pl) != 0)
        return PGSD_ERROR_INVALID_ARGUMENT;
    *id_first = (uint64_t)rank * pl.nbr * pl.cap;
    *max_rows = (uint64_t)pl.nbr * pl.cap;
    return PGSD_SUCCESS;
    }

int pgsd_b200_reorder_distributed(uint64_t n_local, const uint32_t* keys_device, uint64_t out_capacity,
                                  uint64_t* n_out, uint64_t* id_first, uint32_t* keys_sorted_device,
                                  int nfields, const struct pgsd_b200_field* fields_device, void* cuda_stream)
    {
    std::vector<ReorderField> f;
    int rc = to_fields(nfields, fields_device, f);
    if (rc != 0)
        return rc;
    return dev_reorder_distributed(n_local, keys_device, out_capacity, n_out, id_first, keys_sorted_device, nfields,
                                   f.data(), cuda_stream);
    }

// ------------------------------------------------------------------ accounting
int pgsd_b200_get_stats(struct pgsd_b200_stats* out)
    {
    if (out == nullptr)
        return PGSD_ERROR_INVALID_ARGUMENT;
    const DevStats& s = dev_stats();
    out->kernel_launches = s.kernel_launches;
    out->h2d_bytes = s.h2d_bytes;
    out->d2h_bytes = s.d2h_bytes;
    out->file_bytes_written = s.file_bytes_written;
    out->file_bytes_read = s.file_bytes_read;
    out->collectives = g_collectives;
    out->commit_wait_s = s.commit_wait_s;
    out->d2h_busy_s = s.d2h_busy_s;
    out->file_busy_s = s.file_busy_s;
    out->pieces = s.pieces;
    return PGSD_SUCCESS;
    }

int pgsd_b200_reset_stats(void)
    {
    dev_stats() = DevStats();
    g_collectives = 0;
    return PGSD_SUCCESS;
    }

// ------------------------------------------------------------------ raw device helpers
// For callers without a CUDA runtime of their own (the ctypes host layer, the replay tool, bench.py).
int pgsd_b200_malloc(void** p, uint64_t bytes) { return dev_malloc(p, bytes); }
int pgsd_b200_free(void* p) { return dev_free(p); }
int pgsd_b200_host_alloc(void** p, uint64_t bytes) { return dev_host_alloc(p, bytes); }
int pgsd_b200_host_free(void* p) { return dev_host_free(p); }
int pgsd_b200_memcpy(void* dst, const void* src, uint64_t bytes, int kind) { return dev_memcpy(dst, src, bytes, kind); }
int pgsd_b200_synchronize(void) { return dev_synchronize(); }
int pgsd_b200_drain(void) { return dev_drain(); }
int pgsd_b200_shutdown(void)
    {
    dev_shutdown();
    return PGSD_SUCCESS;
    }
int pgsd_b200_timer_create(void** t) { return dev_timer_create(t); }
int pgsd_b200_timer_start(void* t) { return dev_timer_start(t); }
int pgsd_b200_timer_stop(void* t, float* ms) { return dev_timer_stop(t, ms); }
int pgsd_b200_timer_destroy(void* t) { return dev_timer_destroy(t); }
int pgsd_b200_flush_l2(void) { return dev_flush_l2(); }
int pgsd_b200_pack_profiling(int on)
    {
    dev_pack_profiling(on != 0);
    return 0;
    }
int pgsd_b200_pack_last_ms(float* ms) { return ms ? dev_pack_last_ms(ms) : PGSD_ERROR_INVALID_ARGUMENT; }
int pgsd_b200_reorder_profiling(int on)
    {
    dev_reorder_profiling(on != 0);
    return 0;
    }
int pgsd_b200_reorder_phase_ms(float* out4) { return out4 ? dev_reorder_phase_ms(out4) : PGSD_ERROR_INVALID_ARGUMENT; }
int pgsd_b200_file_stage_write(int fd, const void* buf, uint64_t off, uint64_t len, int mode)
    {
    if (fd < 0 || (len > 0 && buf == nullptr) || mode < 0 || mode > 2)
        return PGSD_ERROR_INVALID_ARGUMENT;
    const bool use_mmap = mode == 2 || (mode == 0 && file_is_tmpfs(fd));
    uint64_t left = 0;
    return file_write_piece(fd, (const char*)buf, off, len, use_mmap, &left) ? PGSD_SUCCESS : PGSD_ERROR_IO;
    }

int pgsd_b200_file_stage_ceiling(const char* path, uint64_t off, uint64_t bytes, double* seconds, int* threads_used, int* mapped)
    {
    if (path == nullptr || seconds == nullptr)
        return PGSD_ERROR_INVALID_ARGUMENT;
    const int fd = open(path, O_RDWR | O_CREAT, 0644);
    if (fd < 0)
        return PGSD_ERROR_IO;
    int writers = 8, pw = 2, mode = 0;
    dev_file_stage_config(&writers, &pw, &mode);
    const bool use_mmap = mode == 2 || (mode == 0 && file_is_tmpfs(fd));
    const int threads = use_mmap ? writers : (pw < writers ? pw : writers);
    *seconds = file_stage_ceiling(fd, off, bytes, dev_slot_bytes(), threads, use_mmap);
    if (threads_used)
        *threads_used = threads;
    if (mapped)
        *mapped = use_mmap ? 1 : 0;
    close(fd);
    return *seconds < 0 ? PGSD_ERROR_IO : PGSD_SUCCESS;
    }

int pgsd_b200_read_ahead_stats(uint64_t* hits, uint64_t* issued, uint64_t* dropped)
    {
    dev_read_ahead_stats(hits, issued, dropped);
    return PGSD_SUCCESS;
    }

// The same state machine on host memory (read_ahead.cpp has no CUDA in it): what the CPU suite hammers from several
// threads.  Staging = malloc, a read = pread, the copy out of staging = memcpy.
namespace
    {
bool hra_read_now(int fd, void* dst, uint64_t bytes, uint64_t off)
    {
    uint64_t got = 0;
    while (got < bytes)
        {
        const ssize_t k = pread(fd, (char*)dst + got, bytes - got, (off_t)(off + got));
        if (k < 0 && errno == EINTR)
            continue;
        if (k <= 0)
            return false;
        got += (uint64_t)k;
        }
    return true;
    }
bool hra_alloc(void** p, uint64_t bytes)
    {
    *p = malloc(bytes);
    return *p != nullptr;
    }
void hra_release(void* p) { free(p); }
bool hra_copy(void* dst, const void* src, uint64_t bytes)
    {
    memcpy(dst, src, bytes);
    return true;
    }
ReadAhead g_host_ra(ReadAheadOps { hra_read_now, hra_alloc, hra_release, hra_copy, nullptr });
    } // namespace

int pgsd_b200_read_ahead_host_read(int fd, void* host_dst, uint64_t bytes, uint64_t off)
    {
    if (fd < 0 || (bytes > 0 && host_dst == nullptr))
        return PGSD_ERROR_INVALID_ARGUMENT;
    if (bytes == 0)
        return PGSD_SUCCESS;
    return g_host_ra.read(fd, host_dst, bytes, off) ? PGSD_SUCCESS : PGSD_ERROR_IO;
    }
int pgsd_b200_read_ahead_host_reset(void)
    {
    g_host_ra.reset();
    return PGSD_SUCCESS;
    }
int pgsd_b200_read_ahead_host_stats(uint64_t* hits, uint64_t* issued, uint64_t* dropped)
    {
    g_host_ra.stats(hits, issued, dropped);
    return PGSD_SUCCESS;
    }

int pgsd_b200_selftest(int which)
    {
    if (which == 0)
        return dev_selftest_mbar_timeout();
    return PGSD_ERROR_INVALID_ARGUMENT;
    }

} // extern "C"
