// device.h -- boundary between the host file layer (pgsd_file.cpp, api_b200.cpp) and the CUDA
// side (device.cu: context, frame arena, pinned staging ring + writer threads, NCCL transport,
// K2 scan; kernels_pack.cu: K1; kernels_sort.cu: K4/K5).  No CUDA types cross this header so
// the host translation units compile with a plain C++ compiler.
#pragma once
#include "comm.h"

#include <cstddef>
#include <cstdint>
#include <string>

namespace pgsdb
{
struct Column
    {
    const void* base;
    int64_t stride; // elements of the source type
    };

struct ReorderField
    {
    const void* in;
    void* out;
    uint32_t row_bytes;
    };

struct DevStats
    {
    uint64_t kernel_launches = 0;
    uint64_t h2d_bytes = 0;
    uint64_t d2h_bytes = 0;
    uint64_t file_bytes_written = 0;
    uint64_t file_bytes_read = 0;
    double commit_wait_s = 0;
    double d2h_busy_s = 0;  // sum over staged pieces of their D2H copy time (CUDA events on the copy streams)
    double file_busy_s = 0; // sum over staged pieces of the writer thread's time inside the file write
    uint64_t pieces = 0;    // staged pieces
    };
DevStats& dev_stats();

void set_last_error(const std::string& s);
const std::string& last_error();

bool dev_cuda_available();
int dev_init(int device); // idempotent; device < 0 keeps the current device. 0 on success
bool dev_is_device_pointer(const void* p);
void dev_set_user_stream(void* s);
void* dev_user_stream();
int dev_configure_staging(uint32_t n_slots, uint64_t slot_bytes, uint32_t writer_threads);
void dev_file_stage_config(int* writers, int* pwrite_threads, int* mode); // writer threads, pwrite pieces in flight, FileMode
uint64_t dev_slot_bytes();
void dev_shutdown(); // drain, stop threads, free device/pinned memory

size_t type_size(int pgsd_type); // 0 for unknown
bool cast_supported(int src_type, int dst_type);

// ---- K1 into the frame arena ------------------------------------------------------------
// One chunk to pack: dst (N, M) of dst_type from M columns of src_type.  Columns are device
// pointers, or host pointers when host_columns is set (copied to the device first).
struct PackRequest
    {
    int dst_type;
    int src_type;
    uint64_t N;
    uint32_t M;
    const Column* cols;
    bool host_columns;
    void* arena_ptr; // out: packed chunk in the current frame arena (NULL when N == 0)
    };
// Packs all requests with ONE K1 launch on the user stream into fresh allocations of the frame arena *frame
// (NULL: a free arena is acquired -- waiting for one when max_frames packed frames are still on their way to the
// file -- and returned).  The arena belongs to the calling file handle until it is submitted or abandoned.
int dev_arena_pack(PackRequest* reqs, int n, void** frame);

// ---- K3: arena -> pinned ring -> pwrite(fd) on writer threads --------------------------
struct WriteJob
    {
    const void* dev_ptr;
    uint64_t bytes;
    uint64_t file_off;
    };
// Queue the jobs of the frame being assembled and retire its arena; returns at once (the
// arena memory is recycled when its last byte is on its way to the file).
int dev_frame_submit(int fd, const WriteJob* jobs, int njobs, void* frame);
void dev_frame_abandon(void* frame);
int dev_drain(); // wait for every queued file write; PGSD_ERROR_IO (-1) if one failed
// queue a small host write behind the staged frames (copied; written in submission order by one thread); false when
// the staging threads are not running: write synchronously then
bool dev_async_host_write(int fd, const void* buf, uint64_t n, uint64_t off);
int dev_copy_to_host(void* host_dst, const void* dev_src, uint64_t bytes); // synchronous
// file -> pinned double buffer -> device (read path)
// read_only: the handle cannot write -- equally sized reads at a constant stride are fetched ahead (device.cu)
int dev_read_file_to_device(int fd, void* dev_dst, uint64_t bytes, uint64_t file_off, bool read_only);
void dev_read_ahead_reset(); // drop what was fetched ahead (every open / close of a handle)
void dev_read_ahead_stats(uint64_t* hits, uint64_t* issued, uint64_t* dropped);

// ---- K1 alone (no arena, no file)
int dev_pack(void* dst_dev, int dst_type, uint64_t N, uint32_t M, int src_type, const Column* cols,
             void* stream);

// ---- K2 on a host-provided matrix (test hook); the NCCL transport uses the same kernel
int dev_scan_sizes(const uint64_t* sizes, int P, int C, int rank, SizeScan* out);

// ---- K4 / K5
int dev_sort_ids(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm,
                 void* stream);
int dev_gather(uint64_t n, const uint32_t* perm, int nfields, const ReorderField* fields,
               void* stream);
int dev_reorder(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm, int nfields,
                const ReorderField* fields, void* stream);
int dev_reorder_host(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm,
                     int nfields, const ReorderField* fields);
// Geometry of the distributed reorder, a pure function of (rows in total, ranks): 2^L ids per bucket, nbp =
// histogram size (power of two), nb_used = ceil(N / 2^L) buckets hold ids, every rank owns nbr consecutive ones.
struct DistPlan
    {
    int L;
    uint32_t cap, nbp, nb_used, nbr;
    };
int dist_plan(uint64_t n_global, int nranks, DistPlan* out); // 0, or -2 when the frame does not fit the geometry
// one frame partitioned over the ranks (kernels_slot.cu); 1 = ids not unique / not dense, nothing written
int dev_reorder_distributed(uint64_t n_local, const uint32_t* keys, uint64_t out_capacity, uint64_t* n_out,
                            uint64_t* id_first, uint32_t* keys_sorted, int nfields, const ReorderField* fields,
                            void* stream);

int dev_selftest_mbar_timeout(); // kernels_cluster.cu; 0: the bounded mbarrier wait timed out and was reported

// phase timing of the last bucketed reorder (CUDA events): census, bucket pass, pair passes, gather
void dev_pack_profiling(bool on); // K1 launch of the last frame write
int dev_pack_last_ms(float* ms);
void dev_reorder_profiling(bool on);
int dev_reorder_phase_ms(float* out4);

// ---- plain device / pinned memory helpers for callers without a CUDA runtime of their own
int dev_malloc(void** p, uint64_t bytes);
int dev_free(void* p);
int dev_host_alloc(void** p, uint64_t bytes);
int dev_host_free(void* p);
int dev_memcpy(void* dst, const void* src, uint64_t bytes, int kind); // 1 H2D, 2 D2H, 3 D2D
int dev_synchronize();
// CUDA-event timing on the user stream (bench / tests)
int dev_timer_create(void** t);
int dev_timer_start(void* t);
int dev_timer_stop(void* t, float* ms);
int dev_timer_destroy(void* t);
int dev_flush_l2();
} // namespace pgsdb
