/* pgsd_b200.h -- B200 extensions of the PGSD C ABI (libpgsd_b200.so).
 *
 * Everything the reference gets from MPI or does on the host around its hot path, and
 * that has no slot in pgsd.h, enters here: the rank communicator, device-resident SoA
 * chunk writes (K1), the size allgather + exclusive scan (K2), and the particle-id
 * reorder of a decoded frame (K4 radix sort + K5 gather).  Plain pointers and sizes
 * only; no torch / numpy / MPI types.  Each entry cites the reference code it replaces
 * (paths relative to /root/reference/).
 */
#ifndef PGSD_B200_EXT_H
#define PGSD_B200_EXT_H

#include "pgsd.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ communicator
 * Replaces MPI_Init + MPI_COMM_WORLD (ref: pgsd.c:1487-1488 caches rank/nprocs; the
 * benchmarks call MPI_Init, benchmark-write.cc:23).  One process per GPU.  Exactly one
 * transport is active per process; pgsd_open/pgsd_create_and_open capture it.
 *   nccl : sizes are all-gathered with ncclAllGather over NVLink and scanned on the
 *          device (K2).  This is the production transport.
 *   host : the caller supplies an all-gather of uint64 vectors (e.g. gloo / an MPI the
 *          host application already has); used for host-pointer-only operation and for
 *          the CPU test-suite.
 *   shm  : built-in single-host transport over a POSIX shared-memory segment.
 */
typedef int (*pgsd_b200_allgather_fn)(void* ctx, const uint64_t* send, uint64_t* recv,
                                      size_t n_per_rank);

int pgsd_b200_comm_init_host(int rank, int nprocs, pgsd_b200_allgather_fn fn, void* ctx);
int pgsd_b200_comm_init_shm(int rank, int nprocs, const char* segment_name);
/* fills 128 bytes; call on rank 0 and ship the bytes to the other ranks out of band */
int pgsd_b200_nccl_unique_id(void* out128);
int pgsd_b200_comm_init_nccl(int rank, int nprocs, const void* unique_id128, int cuda_device);
int pgsd_b200_comm_finalize(void);
int pgsd_b200_comm_rank(void);
int pgsd_b200_comm_size(void);
const char* pgsd_b200_comm_kind(void); /* "single" | "host" | "shm" | "nccl" */
int pgsd_b200_barrier(void);

/* Caller-side partition offsets: all-gather n_local over the ranks, return the global
   row count and this rank's exclusive prefix.  Replaces the MPI_Allgather + prefix loop
   every multi-rank caller of the reference writes by hand
   (ref: benchmark-write.cc:33-45, benchmark-read.cc:64-76, fl.pyx:596-598). */
int pgsd_b200_partition(uint64_t n_local, uint64_t* n_global, uint64_t* row_start);

/* ------------------------------------------------------------------ device
 * CUDA state is created lazily by the first call that needs it. */
int pgsd_b200_cuda_available(void);     /* 1 if a CUDA device can be used */
int pgsd_b200_device_init(int cuda_device);
/* Stream on which the caller produces device data; K1 pack kernels are launched on it
   so that "data may be reused as soon as the call returns" holds in stream order
   (ref semantics: pgsd.c:521, :2229 finish with the caller's buffer before returning). */
int pgsd_b200_set_stream(void* cuda_stream);
const char* pgsd_b200_last_error(void);

/* pinned staging ring used between the device arena and pwrite (K3);
   defaults: 8 slots x 16 MiB, 8 writer threads.  Call before the first device write. */
int pgsd_b200_configure_staging(uint32_t n_slots, uint64_t slot_bytes, uint32_t writer_threads);

/* Use as `offset` to let the library place this rank's slice at the exclusive prefix
   (over ranks) of N*M elements, computed by K2 at frame commit. */
#define PGSD_B200_OFFSET_AUTO UINT64_MAX

/* One source column of a chunk: element j of row i is read from
   ((const src_type*)base)[i * stride].  SoA: M columns with stride 1.  AoS (e.g. HOOMD's
   Scalar4 pos): M columns base+j with stride 4.  Already packed (N,M): base+j, stride M. */
struct pgsd_b200_column
    {
    const void* base;
    int64_t stride; /* in elements of src_type */
    };

/* K1 -- pgsd_write_chunk for a chunk that still has to be packed and dtype-cast:
   dst[i*M + j] = (dst_type) cols[j][i*stride_j].  Columns are CUDA device pointers (hot
   path) or host pointers (copied host->device through the pinned ring first).
   Replaces the host cast + contiguity copy in front of the reference's write
   (ref: fl.pyx:571 numpy.ascontiguousarray, hoomd.py:206-270 ParticleData.validate)
   followed by pgsd_write_chunk (pgsd.c:2072-2259).  Same return codes as
   pgsd_write_chunk.  Every cast between the ten pgsd types is done as numpy.astype does it on x86-64
   (float -> integer truncates toward zero). */
int pgsd_b200_write_chunk_soa(struct pgsd_handle* handle,
                              const char* name,
                              enum pgsd_type dst_type,
                              uint64_t N,
                              uint32_t M,
                              uint64_t N_global,
                              uint32_t M_global,
                              uint64_t offset,
                              bool all,
                              enum pgsd_type src_type,
                              const struct pgsd_b200_column* cols);

/* A whole frame's SoA chunks in one call: ONE K1 launch packs all of them (<= 16 per launch),
   then each is recorded exactly as pgsd_b200_write_chunk_soa would, in array order. */
struct pgsd_b200_chunk_desc
    {
    const char* name;
    enum pgsd_type dst_type;
    enum pgsd_type src_type;
    uint64_t N;
    uint32_t M;
    uint64_t N_global; /* UINT64_MAX: let the library sum N over the ranks at frame commit */
    uint32_t M_global;
    uint64_t offset; /* elements, or PGSD_B200_OFFSET_AUTO */
    bool all;
    const struct pgsd_b200_column* cols; /* M columns, all device or all host */
    };
int pgsd_b200_write_chunks_soa(struct pgsd_handle* handle, int n_chunks,
                               const struct pgsd_b200_chunk_desc* chunks);

/* K1 alone: pack device columns into a device buffer on `cuda_stream` (no file). */
int pgsd_b200_pack_soa(void* dst_device,
                       enum pgsd_type dst_type,
                       uint64_t N,
                       uint32_t M,
                       enum pgsd_type src_type,
                       const struct pgsd_b200_column* cols_device,
                       void* cuda_stream);

/* K2 alone: sizes[P][C] (host) -> per chunk: exclusive prefix over ranks < rank, sum and
   max over ranks, computed by the device scan kernel.
   Replaces pgsd.c:1126 + :1150-1152 (Allgather + prefix), :1162/:2242 (SUM), :2157 (MAX). */
int pgsd_b200_scan_sizes(const uint64_t* sizes, int P, int C, int rank,
                         uint64_t* excl, uint64_t* total, uint64_t* maxv);

/* ------------------------------------------------------------------ reorder (K4 + K5)
 * Oracle definition (BASELINE.json north_star): o = numpy.argsort(keys, kind='stable');
 * out_f = in_f[o] for every field.  Bit-exact: payload is moved, never computed on. */
struct pgsd_b200_field
    {
    const void* in;     /* n rows of row_bytes */
    void* out;          /* n rows of row_bytes, must not alias in */
    uint32_t row_bytes; /* 1..1024, e.g. 12 for an (N,3) float32 field; multiples of 4 take the fast path */
    };

/* K4: stable LSD radix sort of (key, original index) pairs on the device.
   keys_sorted and perm may be NULL if not wanted. */
int pgsd_b200_sort_ids(uint64_t n, const uint32_t* keys_device, uint32_t* keys_sorted_device,
                       uint32_t* perm_device, void* cuda_stream);

/* K5: out_f[i] = in_f[perm[i]] for all fields in one launch. */
int pgsd_b200_gather(uint64_t n, const uint32_t* perm_device, int nfields,
                     const struct pgsd_b200_field* fields_device, void* cuda_stream);

/* K4+K5 on device-resident data.  Unique keys with at most 27 varying bits (particle ids) and word-sized
   fields of at most 31 words per row in total take the slot path: two passes over interleaved records, no
   ranking (kernels_slot.cu; workspace n * (row bytes + 4), cached between calls); the call then returns
   after the stream has been synchronised, because a device flag decides whether duplicates were found.
   Anything else takes the stable general path: frames of >= 1 Mi rows (PGSD_B200_BUCKET_MIN_ROWS) first
   get a bucket pass -- rows grouped by the top key bits into a workspace copy (n * (row bytes + 8) bytes)
   so that the gather is L2-local -- then LSD passes on (key, index) and the gather. */
int pgsd_b200_reorder_device(uint64_t n, const uint32_t* keys_device, uint32_t* keys_sorted_device,
                             uint32_t* perm_device, int nfields,
                             const struct pgsd_b200_field* fields_device, void* cuda_stream);

/* Host buffers in, host buffers out: pinned H2D, K4+K5, pinned D2H inside the call. */
int pgsd_b200_reorder_host(uint64_t n, const uint32_t* keys_host, uint32_t* keys_sorted_host,
                           uint32_t* perm_host, int nfields,
                           const struct pgsd_b200_field* fields_host);

/* K4+K5 for ONE frame whose rows are partitioned over the ranks (one process per GPU; SURVEY.md section 8e,
   config 3 read-back: rank r has read file partition r -- pgsd_read_chunk(..., all=true), pgsd.c:2497-2508 --
   and the frame must come out in particle-id order).  Collective over the communicator, at most 8 ranks.
   Rank r ends up with the rows whose ids lie in [*id_first, *id_first + S), S = ceil(ceil(N / C) / ranks) * C
   with C = 1024 (2048 / 4096 for more than 32 Mi / 64 Mi rows in total), in id order: *n_out rows written to
   keys_sorted_device and to every fields[i].out (capacity out_capacity rows each; S always suffices).
   The exchange is not a separate collective: a kernel stores every record straight into the owner's memory
   (CUDA IPC mapping, i.e. NVLink) -- appended to the owner's inbox in contiguous runs (default) or, with
   PGSD_B200_DIST_MODE=fused, at its final place in the owner's bucketed copy -- and the owner finishes locally.
   Returns 0, a negative pgsd error, or 1 on EVERY rank when the ids are not unique or not all below
   ceil(N / C) * C (dense ids 0..N-1 qualify): nothing was written, gather the frame to one GPU and use
   pgsd_b200_reorder_device.  row_bytes must be multiples of 4; keys 16-byte aligned. */
/* Host-only: which ids rank `rank` of `nranks` will own for a frame of n_global rows -- [*id_first,
   *id_first + *max_rows) -- so that callers can size their outputs before the collective call. */
int pgsd_b200_reorder_distributed_plan(uint64_t n_global, int nranks, int rank, uint64_t* id_first,
                                       uint64_t* max_rows);
int pgsd_b200_reorder_distributed(uint64_t n_local, const uint32_t* keys_device, uint64_t out_capacity,
                                  uint64_t* n_out, uint64_t* id_first, uint32_t* keys_sorted_device,
                                  int nfields, const struct pgsd_b200_field* fields_device, void* cuda_stream);

/* ------------------------------------------------------------------ accounting */
struct pgsd_b200_stats
    {
    uint64_t kernel_launches; /* CUDA kernels launched by this library */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t file_bytes_written;
    uint64_t file_bytes_read;
    uint64_t collectives; /* all-gathers issued through the communicator */
    double commit_wait_s; /* host time blocked in frame commits (D2H + pwrite drain) */
    double d2h_busy_s;    /* sum over staged pieces of their D2H copy time (CUDA events on the copy streams) */
    double file_busy_s;   /* sum over staged pieces of the writer thread's time inside the file write */
    uint64_t pieces;      /* staged pieces (16 MiB each unless configured otherwise) */
    };
int pgsd_b200_get_stats(struct pgsd_b200_stats* out);
int pgsd_b200_reset_stats(void);

/* ------------------------------------------------------------------ raw device helpers
 * For host layers without a CUDA runtime of their own (the Python file layer, the replay tool,
 * bench.py): device / pinned allocation, synchronous copies (kind: 1 H2D, 2 D2H, 3 D2D),
 * CUDA-event timers on the stream set with pgsd_b200_set_stream, and an L2 flush (writes a
 * 256 MiB buffer).  pgsd_b200_drain waits for every queued file write of this rank. */
int pgsd_b200_malloc(void** p, uint64_t bytes);
int pgsd_b200_free(void* p);
int pgsd_b200_host_alloc(void** p, uint64_t bytes);
int pgsd_b200_host_free(void* p);
int pgsd_b200_memcpy(void* dst, const void* src, uint64_t bytes, int kind);
int pgsd_b200_synchronize(void);
int pgsd_b200_drain(void);
int pgsd_b200_shutdown(void);
int pgsd_b200_timer_create(void** t);
int pgsd_b200_timer_start(void* t);
int pgsd_b200_timer_stop(void* t, float* ms);
int pgsd_b200_timer_destroy(void* t);
int pgsd_b200_flush_l2(void);
/* Per-phase device time of the last pgsd_b200_reorder_device call that took the bucketed path
   (CUDA events on the caller's stream): out4 = {key census, bucket pass, pair passes, gather} in ms.
   Measurement hook for bench.py; off by default. */
int pgsd_b200_reorder_profiling(int on);
/* Device time of the K1 launch of the last pgsd_b200_write_chunk(s)_soa / device pgsd_write_chunk
   (CUDA events recorded on the caller's stream directly around the launch). */
int pgsd_b200_pack_profiling(int on);
int pgsd_b200_pack_last_ms(float* ms);
int pgsd_b200_reorder_phase_ms(float* out4);
/* K3 file stage without the device (file_stage.cpp), for tests and for measuring the host-side ceiling:
   pgsd_b200_file_stage_write puts one piece into fd the way a writer thread does (mode 0: auto = mappings on
   tmpfs, pwrite elsewhere; 1: pwrite; 2: mappings); pgsd_b200_file_stage_ceiling writes `bytes` bytes of host
   memory at [off, off + bytes) of `path` with the library's own piece size / thread counts / mode and returns the
   seconds it took (bench.py reports file_ceiling_GBps from it, on the same target as the timed frames). */
int pgsd_b200_file_stage_write(int fd, const void* buf, uint64_t off, uint64_t len, int mode);
int pgsd_b200_file_stage_ceiling(const char* path, uint64_t off, uint64_t bytes, double* seconds, int* threads_used, int* mapped);
/* Read-ahead of pgsd_read_chunk into device memory (read-only handles; replaces nothing in the reference -- its
   pgsd_read_chunk, pgsd.c:2436-2537, is one blocking MPI_File_read_at per call): after three equally sized reads
   at a constant file stride the next two ranges are fetched into device staging buffers in the background.
   Counters since the library was loaded: calls served from staging, ranges fetched ahead, fetched ranges never
   used.  Opt-in: PGSD_B200_READ_AHEAD=1 (read per call). */
int pgsd_b200_read_ahead_stats(uint64_t* hits, uint64_t* issued, uint64_t* dropped);
/* The same read-ahead state machine on HOST memory (staging = malloc, a read = pread), always on: for the CPU
   test-suite, which reads files through it from several threads (tests/test_read_ahead_host.py). */
int pgsd_b200_read_ahead_host_read(int fd, void* host_dst, uint64_t bytes, uint64_t off);
int pgsd_b200_read_ahead_host_reset(void);
int pgsd_b200_read_ahead_host_stats(uint64_t* hits, uint64_t* issued, uint64_t* dropped);
/* Device self-tests of failure paths that valid inputs never reach.  which = 0: a kernel waits on an mbarrier
   whose bulk copy never arrives; returns 0 when the bounded wait gave up and reported it (the reorder kernels'
   "a bulk copy did not complete" path), > 0 otherwise. */
int pgsd_b200_selftest(int which);

#ifdef __cplusplus
}
#endif
#endif /* PGSD_B200_EXT_H */
