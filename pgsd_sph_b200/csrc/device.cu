// device.cu -- CUDA context of libpgsd_b200: frame arena, K3 staging pipeline (pinned ring,
// stager + writer threads), NCCL transport with the K2 size scan, and small helpers.
//
// K3 replaces the reference's blocking MPI_File_write_at of each rank's chunk bytes
// (/root/reference/pgsd/pgsd/pgsd.c:2229 direct chunks, :1154 buffered chunks): packed chunks
// stay in a device arena, are copied to pinned slots with cudaMemcpyAsync on side streams and
// written with pwrite by writer threads, so packing frame k+1 overlaps the file write of k.
// K2 replaces MPI_Allgather + prefix loop + MPI_Allreduce SUM/MAX (pgsd.c:1126,1150-1152,
// 1162,2157,2242): one ncclAllGather of the frame's u64 size vector + a device scan.
#include "device_internal.h"
#include "file_stage.h"
#include "read_ahead.h"

#include <nvtx3/nvToolsExt.h> // header-only; ranges cost nothing unless a profiler (nsys) is attached
#include <nccl.h> // types only; the library is dlopen()ed so that CPU-only hosts can load us

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <dlfcn.h>
#include <errno.h>
#include <map>
#include <mutex>
#include <thread>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/statvfs.h>
#include <sys/uio.h>
#include <unistd.h>
#include <vector>

namespace pgsdb
{
// ------------------------------------------------------------------------------ errors / stats
static std::string g_last_error;
static std::mutex g_err_mutex;
void set_last_error(const std::string& s)
    {
    std::lock_guard<std::mutex> g(g_err_mutex);
    g_last_error = s;
    }
const std::string& last_error() { return g_last_error; }

static DevStats g_stats;
DevStats& dev_stats() { return g_stats; }

#define CUDA_TRY(call, rcode)                                                                   \
    do                                                                                          \
        {                                                                                       \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            {                                                                                   \
            set_last_error(std::string(#call) + ": " + cudaGetErrorString(e__));                \
            cudaGetLastError();                                                                 \
            return (rcode);                                                                     \
            }                                                                                   \
        } while (0)

// ------------------------------------------------------------------------------ context
namespace
    {
struct Block
    {
    char* ptr = nullptr;
    size_t cap = 0;
    size_t used = 0;
    };

struct ArenaFrame
    {
    std::vector<Block> blocks;
    std::atomic<long> outstanding { 0 }; // pinned-slot writes not yet finished
    cudaEvent_t packed = nullptr;        // recorded on the user stream after the frame's K1 launches
    cudaEvent_t k1[2] = { nullptr, nullptr }; // trace: around the frame's last K1 launch
    bool k1_timed = false;
    bool assembling = false;             // owned by one writable file handle until it is submitted or abandoned
    uint64_t seq = 0;                    // submit order (trace)
    };

struct Slot
    {
    char* host = nullptr;
    cudaEvent_t t0 = nullptr; // recorded on the copy stream before the piece's D2H copy ...
    cudaEvent_t t1 = nullptr; // ... and after it (the writer thread waits on this one)
    };

// PGSD_B200_TRACE=<file>: every K1 launch, D2H piece and file piece with its interval on one clock (host
// steady_clock; device intervals are placed through a calibrated base event), written as a Chrome trace at
// shutdown.  Evidence for the K1(k+1) || D2H(k) || file(k-1) overlap; off by default.
struct TraceEvent
    {
    char kind; // 'K' K1, 'D' D2H piece, 'F' file piece
    uint64_t frame;
    double t0_us, t1_us;
    uint64_t bytes;
    int lane; // stream / writer thread
    };

struct Seg // one chunk inside a bundled copy
    {
    uint64_t dev_off; // from the bundle's first byte
    uint64_t bytes;
    uint64_t file_off;
    };

struct StageJob
    {
    int fd;
    const char* dev;
    uint64_t bytes;
    uint64_t file_off;
    ArenaFrame* frame;
    // Small frames (BASELINE config 5: 4096 particles, 6 device chunks of 16-48 KB): the chunks of a frame are
    // neighbours in the arena, so ONE D2H copy of their span replaces one copy + two event records + a slot per chunk,
    // and the writer thread puts them into the file with pwritev (file-contiguous runs in one call).
    std::vector<Seg> segs;
    };

struct WriteItem
    {
    int slot;
    int fd;
    uint64_t file_off;
    uint64_t bytes;
    ArenaFrame* frame;
    int stream; // copy stream the piece was staged on (trace)
    std::vector<Seg> segs;
    std::vector<char> host_bytes; // slot < 0: bytes handed over by the file layer (index entries, write buffer, names)
    };

struct Ctx
    {
    bool inited = false;
    bool failed = false;
    int device = 0;
    int sm_count = 148;
    cudaStream_t user = nullptr; // legacy default stream unless the caller sets one
    cudaStream_t copy[2] = { nullptr, nullptr };
    cudaStream_t aux = nullptr;

    // staging configuration
    uint32_t n_slots = 8;
    uint64_t slot_bytes = 16ull << 20;
    uint32_t n_writers = 8;
    FileMode file_mode = FileMode::Auto; // PGSD_B200_FILE_MODE: auto (mappings on tmpfs only) | pwrite | mmap
    uint32_t pwrite_threads = 4; // pwrite-mode pieces in flight: buffered writes to one file serialise on the inode lock,
                                 // more threads only add contention (ext4, whole pipeline at 64 Mi-particle frames: 3.74 /
                                 // 4.06 / 4.21 / 4.13 GB/s at 1 / 2 / 4 / 8; a fresh 8 GB file alone: 6.7 at 1-2, 5.2 at 8-16)
    uint32_t pwrite_active = 0;
    std::condition_variable cv_pwrite;
    uint32_t max_frames = 3;     // frames in flight (packed, not yet in the file): large frames
    uint64_t last_frame_bytes = 0; // arena bytes of the frame submitted last: small frames may queue deeper (see acquire_frame)
    uint64_t frame_seq = 0;

    // trace
    bool trace_on = false;
    std::string trace_path;
    std::vector<TraceEvent> trace;
    cudaEvent_t trace_base = nullptr;
    std::chrono::steady_clock::time_point trace_t0;

    std::vector<Slot> slots;
    std::vector<int> free_slots;
    std::deque<StageJob> stage_q;
    std::deque<WriteItem> write_q;
    // Everything small goes through ONE thread in submission order: bundled frames and the file layer's own small
    // writes (rank 0's index entries, every rank's write buffer).  Buffered writes to one file serialise on the inode
    // lock, so several threads issuing 100-byte to 100-KB writes only queue on that lock, and the caller's
    // pgsd_end_frame queued behind them (measured at 4096-particle frames: 65 us of an 85-us frame inside
    // end_frame with 8 writer threads, 46 us with 2).  FIFO order also keeps overlapping index writes in order.
    std::deque<WriteItem> serial_q;
    std::condition_variable cv_serial;
    std::thread serial_writer;
    std::mutex mu;
    std::condition_variable cv_stage, cv_write, cv_slot, cv_done;
    std::thread stager;
    std::vector<std::thread> writers;
    bool threads_running = false;
    bool stop = false;
    long jobs_outstanding = 0; // stage jobs + write items in flight
    std::atomic<bool> io_error { false };

    std::vector<ArenaFrame*> frames;

    // reorder_host scratch
    char* scratch = nullptr;
    size_t scratch_bytes = 0;
    };
Ctx g;

double trace_now_us()
    {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - g.trace_t0).count();
    }
// device interval [e0, e1] on the host clock (both events have completed)
bool trace_device_interval(cudaEvent_t e0, cudaEvent_t e1, double* t0_us, double* t1_us)
    {
    float a = 0.f, b = 0.f;
    if (cudaEventElapsedTime(&a, g.trace_base, e0) != cudaSuccess || cudaEventElapsedTime(&b, g.trace_base, e1) != cudaSuccess)
        {
        cudaGetLastError();
        return false;
        }
    *t0_us = 1e3 * a;
    *t1_us = 1e3 * b;
    return true;
    }
void trace_dump()
    {
    static bool dumped = false;
    if (!g.trace_on || g.trace_path.empty() || (dumped && g.trace.empty()))
        return; // (shutdown and the exit hook both come here: the second call must not truncate the file)
    dumped = true;
    FILE* f = fopen(g.trace_path.c_str(), "w");
    if (!f)
        return;
    fprintf(f, "{\"traceEvents\":[\n");
    bool first = true;
    for (const TraceEvent& e : g.trace)
        {
        const char* name = e.kind == 'K' ? "K1 pack" : e.kind == 'D' ? "D2H piece" : "file piece";
        const int tid = e.kind == 'K' ? 0 : e.kind == 'D' ? 10 + e.lane : 100 + e.lane;
        fprintf(f, "%s{\"name\":\"%s f%llu\",\"cat\":\"%c\",\"ph\":\"X\",\"pid\":%d,\"tid\":%d,\"ts\":%.1f,\"dur\":%.1f,"
                   "\"args\":{\"frame\":%llu,\"bytes\":%llu}}",
                first ? "" : ",\n", name, (unsigned long long)e.frame, e.kind, g.device, tid, e.t0_us, e.t1_us - e.t0_us,
                (unsigned long long)e.frame, (unsigned long long)e.bytes);
        first = false;
        }
    fprintf(f, "\n]}\n");
    fclose(f);
    g.trace.clear();
    }

// One pinned piece -> file: see file_stage.cpp for the two ways and why.
void writer_main(int widx)
    {
    cudaSetDevice(g.device);
    const bool serial = widx < 0;
    std::deque<WriteItem>& q = serial ? g.serial_q : g.write_q;
    std::condition_variable& cv = serial ? g.cv_serial : g.cv_write;
    for (;;)
        {
        WriteItem it;
            {
            std::unique_lock<std::mutex> lk(g.mu);
            cv.wait(lk, [&q] { return g.stop || !q.empty(); });
            if (q.empty())
                return;
            it = std::move(q.front());
            q.pop_front();
            }
        if (it.slot < 0)
            {
            // bytes of the file layer: plain positional write
            uint64_t left = 0;
            const auto t0 = std::chrono::steady_clock::now();
            const bool okh = file_write_piece(it.fd, it.host_bytes.data(), it.file_off, it.host_bytes.size(), false, &left);
            const double busy = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (!okh)
                g.io_error = true;
                {
                std::lock_guard<std::mutex> lk(g.mu);
                g_stats.file_bytes_written += it.host_bytes.size() - left;
                g_stats.file_busy_s += busy;
                g.jobs_outstanding--;
                }
            g.cv_done.notify_all();
            continue;
            }
        Slot& sl = g.slots[it.slot];
        bool ok = cudaEventSynchronize(sl.t1) == cudaSuccess;
        float d2h_ms = 0.f;
        if (ok && cudaEventElapsedTime(&d2h_ms, sl.t0, sl.t1) != cudaSuccess)
            {
            cudaGetLastError();
            d2h_ms = 0.f;
            }
        double dt0 = 0, dt1 = 0;
        const bool traced = g.trace_on && ok && trace_device_interval(sl.t0, sl.t1, &dt0, &dt1);
        uint64_t left = it.bytes;
        double busy = 0, ft0 = 0, ft1 = 0;
        if (ok)
            {
            const bool use_mmap = g.file_mode == FileMode::Mmap || (g.file_mode == FileMode::Auto && file_is_tmpfs(it.fd));
            if (!use_mmap)
                {
                std::unique_lock<std::mutex> lk(g.mu);
                g.cv_pwrite.wait(lk, [] { return g.pwrite_active < g.pwrite_threads; });
                g.pwrite_active++;
                }
            nvtxRangePushA("pgsd K3 file piece");
            const auto t0 = std::chrono::steady_clock::now();
            if (g.trace_on)
                ft0 = trace_now_us();
            if (it.segs.empty())
                ok = file_write_piece(it.fd, sl.host, it.file_off, it.bytes, use_mmap, &left);
            else
                {
                // bundled small chunks: file-contiguous runs leave with one pwritev each
                left = 0;
                size_t i = 0;
                while (ok && i < it.segs.size())
                    {
                    struct iovec iov[16];
                    int k = 0;
                    uint64_t run = 0;
                    const uint64_t at = it.segs[i].file_off;
                    while (i < it.segs.size() && k < 16 && it.segs[i].file_off == at + run)
                        {
                        iov[k].iov_base = sl.host + it.segs[i].dev_off;
                        iov[k].iov_len = (size_t)it.segs[i].bytes;
                        run += it.segs[i].bytes;
                        k++;
                        i++;
                        }
                    const ssize_t w = pwritev(it.fd, iov, k, (off_t)at);
                    if (w != (ssize_t)run)
                        {
                        // short or failed vector write: the plain path finishes (or reports) every segment
                        for (int q = 0; q < k && ok; q++)
                            {
                            uint64_t l2 = 0;
                            uint64_t o = at;
                            for (int r = 0; r < q; r++)
                                o += iov[r].iov_len;
                            ok = file_write_piece(it.fd, (const char*)iov[q].iov_base, o, iov[q].iov_len, false, &l2);
                            left += l2;
                            }
                        }
                    }
                }
            if (g.trace_on)
                ft1 = trace_now_us();
            busy = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            nvtxRangePop();
            if (!use_mmap)
                {
                    {
                    std::lock_guard<std::mutex> lk(g.mu);
                    g.pwrite_active--;
                    }
                g.cv_pwrite.notify_one();
                }
            }
        if (!ok)
            g.io_error = true;
            {
            std::lock_guard<std::mutex> lk(g.mu);
            g.free_slots.push_back(it.slot);
            g_stats.file_bytes_written += it.bytes - left;
            g_stats.file_busy_s += busy;
            g_stats.d2h_busy_s += 1e-3 * d2h_ms;
            g_stats.pieces++;
            if (traced)
                {
                g.trace.push_back(TraceEvent { 'D', it.frame->seq, dt0, dt1, it.bytes, it.stream });
                g.trace.push_back(TraceEvent { 'F', it.frame->seq, ft0, ft1, it.bytes, widx });
                }
            it.frame->outstanding--;
            g.jobs_outstanding--;
            }
        g.cv_slot.notify_one();
        g.cv_done.notify_all();
        }
    }

void stager_main()
    {
    cudaSetDevice(g.device);
    unsigned rr = 0;
    for (;;)
        {
        StageJob job;
            {
            std::unique_lock<std::mutex> lk(g.mu);
            g.cv_stage.wait(lk, [] { return g.stop || !g.stage_q.empty(); });
            if (g.stage_q.empty())
                return;
            job = std::move(g.stage_q.front());
            g.stage_q.pop_front();
            }
        uint64_t done = 0;
        if (!job.segs.empty())
            {
            int slot;
                {
                std::unique_lock<std::mutex> lk(g.mu);
                g.cv_slot.wait(lk, [] { return !g.free_slots.empty(); });
                slot = g.free_slots.back();
                g.free_slots.pop_back();
                g.jobs_outstanding++;
                job.frame->outstanding++;
                }
            const int si = (int)(rr++ & 1);
            cudaStream_t st = g.copy[si];
            bool ok = cudaStreamWaitEvent(st, job.frame->packed, 0) == cudaSuccess
                      && cudaEventRecord(g.slots[slot].t0, st) == cudaSuccess
                      && cudaMemcpyAsync(g.slots[slot].host, job.dev, job.bytes, cudaMemcpyDeviceToHost, st) == cudaSuccess
                      && cudaEventRecord(g.slots[slot].t1, st) == cudaSuccess;
            if (!ok)
                g.io_error = true;
            uint64_t payload = 0;
            for (const Seg& sg : job.segs)
                payload += sg.bytes;
                {
                std::lock_guard<std::mutex> lk(g.mu);
                g_stats.d2h_bytes += job.bytes;
                WriteItem wi { slot, job.fd, 0, payload, job.frame, si, {}, {} };
                wi.segs.swap(job.segs);
                g.serial_q.push_back(std::move(wi));
                }
            g.cv_serial.notify_one();
            done = job.bytes;
            }
        while (done < job.bytes)
            {
            // interior piece boundaries fall on page boundaries of the FILE (file_stage.cpp: page ownership)
            uint64_t len = done == 0 ? file_first_piece_len(job.file_off, job.bytes, g.slot_bytes)
                                     : (job.bytes - done < g.slot_bytes ? job.bytes - done : g.slot_bytes);
            int slot;
                {
                std::unique_lock<std::mutex> lk(g.mu);
                g.cv_slot.wait(lk, [] { return !g.free_slots.empty(); });
                slot = g.free_slots.back();
                g.free_slots.pop_back();
                g.jobs_outstanding++;
                job.frame->outstanding++;
                }
            const int si = (int)(rr++ & 1);
            cudaStream_t st = g.copy[si];
            nvtxRangePushA("pgsd K3 D2H piece (enqueue)");
            bool ok = cudaStreamWaitEvent(st, job.frame->packed, 0) == cudaSuccess
                      && cudaEventRecord(g.slots[slot].t0, st) == cudaSuccess
                      && cudaMemcpyAsync(g.slots[slot].host, job.dev + done, len, cudaMemcpyDeviceToHost, st)
                             == cudaSuccess
                      && cudaEventRecord(g.slots[slot].t1, st) == cudaSuccess;
            nvtxRangePop();
            if (!ok)
                g.io_error = true;
                {
                std::lock_guard<std::mutex> lk(g.mu);
                g_stats.d2h_bytes += len;
                g.write_q.push_back(WriteItem { slot, job.fd, job.file_off + done, len, job.frame, si, {}, {} });
                }
            g.cv_write.notify_one();
            done += len;
            }
            {
            std::lock_guard<std::mutex> lk(g.mu);
            job.frame->outstanding--; // the job's own reference
            g.jobs_outstanding--;
            }
        g.cv_done.notify_all();
        }
    }

void stop_threads_at_exit();

int start_threads()
    {
    if (g.threads_running)
        return 0;
    g.slots.resize(g.n_slots);
    for (uint32_t i = 0; i < g.n_slots; i++)
        {
        CUDA_TRY(cudaHostAlloc((void**)&g.slots[i].host, g.slot_bytes, cudaHostAllocDefault), -6);
        CUDA_TRY(cudaEventCreate(&g.slots[i].t0), -1);
        CUDA_TRY(cudaEventCreate(&g.slots[i].t1), -1);
        g.free_slots.push_back((int)i);
        }
    g.stop = false;
    static bool exit_hook = false;
    if (!exit_hook)
        {
        // joinable std::thread objects must not reach static destruction: finish queued file
        // writes and join at exit (registered after the CUDA runtime came up, so it runs first)
        atexit([]() { dev_drain(); stop_threads_at_exit(); trace_dump(); });
        exit_hook = true;
        }
    g.stager = std::thread(stager_main);
    g.serial_writer = std::thread(writer_main, -1);
    for (uint32_t i = 0; i < g.n_writers; i++)
        g.writers.emplace_back(writer_main, (int)i);
    g.threads_running = true;
    return 0;
    }

void stop_threads();
void stop_threads_at_exit() { stop_threads(); }

void stop_threads()
    {
    if (!g.threads_running)
        return;
        {
        std::lock_guard<std::mutex> lk(g.mu);
        g.stop = true;
        }
    g.cv_stage.notify_all();
    g.cv_write.notify_all();
    g.cv_serial.notify_all();
    g.stager.join();
    g.serial_writer.join();
    for (auto& t : g.writers)
        t.join();
    g.writers.clear();
    for (auto& s : g.slots)
        {
        if (s.host)
            cudaFreeHost(s.host);
        if (s.t0)
            cudaEventDestroy(s.t0);
        if (s.t1)
            cudaEventDestroy(s.t1);
        }
    g.slots.clear();
    g.free_slots.clear();
    g.threads_running = false;
    }

int arena_alloc(ArenaFrame* f, uint64_t bytes, void** out)
    {
    const size_t need = (bytes + 255) / 256 * 256;
    for (auto& b : f->blocks)
        if (b.cap - b.used >= need)
            {
            *out = b.ptr + b.used;
            b.used += need;
            return 0;
            }
    size_t total = 0;
    for (auto& b : f->blocks)
        total += b.cap;
    size_t cap = need > total ? need : total;
    if (cap < (1u << 20))
        cap = 1u << 20;
    Block nb;
    CUDA_TRY(cudaMalloc((void**)&nb.ptr, cap), -6);
    nb.cap = cap;
    nb.used = need;
    f->blocks.push_back(nb);
    *out = nb.ptr;
    return 0;
    }

// A frame arena for a file handle that starts assembling a frame.  At most max_frames packed frames wait for the
// file at a time: when that many are in flight the caller waits for the oldest to land (this is the back-pressure
// of the whole pipeline).  Frames that other handles are still assembling do not count -- they cannot finish by
// waiting -- so a new arena is created for the caller instead.
int acquire_frame(ArenaFrame** out)
    {
    ArenaFrame* got = nullptr;
    std::unique_lock<std::mutex> lk(g.mu);
    for (;;)
        {
        size_t in_flight = 0;
        for (ArenaFrame* f : g.frames)
            {
            if (f->assembling)
                continue;
            if (f->outstanding.load() == 0)
                {
                got = f;
                break;
                }
            in_flight++;
            }
        if (got)
            break;
        // Small frames (config 5: 164 KB) spend ~0.2 ms between K1 and the file however short they are; with 3 in flight
        // that latency, not any throughput, would set the frame rate.  Up to 64 frames / 256 MiB may queue instead.
        uint64_t allowed = g.max_frames;
        if (g.last_frame_bytes > 0)
            {
            const uint64_t k = (256ull << 20) / g.last_frame_bytes;
            allowed = k < g.max_frames ? g.max_frames : (k > 64 ? 64 : k);
            }
        if (in_flight < allowed)
            {
            ArenaFrame* f = new ArenaFrame;
            if (cudaEventCreateWithFlags(&f->packed, cudaEventDisableTiming) != cudaSuccess
                || cudaEventCreate(&f->k1[0]) != cudaSuccess || cudaEventCreate(&f->k1[1]) != cudaSuccess)
                {
                delete f;
                set_last_error("cudaEventCreate failed");
                cudaGetLastError();
                return -1;
                }
            g.frames.push_back(f);
            got = f;
            break;
            }
        auto t0 = std::chrono::steady_clock::now();
        g.cv_done.wait(lk);
        g_stats.commit_wait_s
            += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
    got->assembling = true;
    got->k1_timed = false;
    for (auto& bl : got->blocks)
        bl.used = 0;
    *out = got;
    return 0;
    }
    } // namespace

// File-stage / reader threads when the environment does not say: 8 on a host of our own; when several ranks share
// the host (LOCAL_WORLD_SIZE of torchrun, the local sizes of Open MPI / MVAPICH / Slurm) the cores are divided, at
// least 2 per rank -- 8 ranks x 8 writers + 8 stagers on 16 cores only took turns on the same page-cache locks.
static uint32_t default_io_threads()
    {
    long local = 1;
    for (const char* v : { "LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "MV2_COMM_WORLD_LOCAL_SIZE", "SLURM_NTASKS_PER_NODE" })
        if (const char* e = getenv(v))
            {
            const long k = atol(e);
            if (k >= 1)
                {
                local = k;
                break;
                }
            }
    const long cores = (long)std::thread::hardware_concurrency();
    long n = cores > 0 ? cores / local : 8;
    if (n > 8)
        n = 8;
    if (n < 2)
        n = 2;
    return (uint32_t)n;
    }

int dev_sm_count() { return g.sm_count; }

bool dev_cuda_available()
    {
    if (g.inited)
        return true;
    if (g.failed)
        return false;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        {
        cudaGetLastError();
        return false;
        }
    return true;
    }

int dev_init(int device)
    {
    if (g.inited)
        {
        if (device >= 0 && device != g.device)
            {
            set_last_error("device already initialised on another GPU");
            return -2;
            }
        // The current device is per host thread and defaults to GPU 0: a caller's helper thread
        // (e.g. the frame prefetcher of pgsd.hoomd) must land on this rank's GPU as well.
        static thread_local bool bound = false;
        if (!bound)
            {
            CUDA_TRY(cudaSetDevice(g.device), -1);
            bound = true;
            }
        return 0;
        }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        {
        cudaGetLastError();
        g.failed = true;
        set_last_error("no CUDA device available: libpgsd_b200's device path has no CPU fallback");
        return -1;
        }
    if (device < 0)
        {
        if (cudaGetDevice(&device) != cudaSuccess)
            device = 0;
        }
    CUDA_TRY(cudaSetDevice(device), -1);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device), -1);
    g.device = device;
    g.sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&g.copy[0], cudaStreamNonBlocking), -1);
    CUDA_TRY(cudaStreamCreateWithFlags(&g.copy[1], cudaStreamNonBlocking), -1);
    CUDA_TRY(cudaStreamCreateWithFlags(&g.aux, cudaStreamNonBlocking), -1);
    g.file_mode = file_mode_from_env();
    g.n_writers = default_io_threads();
    if (const char* w = getenv("PGSD_B200_WRITER_THREADS"))
        {
        int n = atoi(w);
        if (n >= 1 && n <= 64)
            g.n_writers = (uint32_t)n;
        }
    if (const char* w = getenv("PGSD_B200_PWRITE_THREADS"))
        {
        int n = atoi(w);
        if (n >= 1 && n <= 64)
            g.pwrite_threads = (uint32_t)n;
        }
    if (const char* t = getenv("PGSD_B200_TRACE"))
        if (*t)
            {
            // device intervals are measured from this event; its completion time is host time 0 of the trace
            g.trace_path = t;
            if (g.trace_path.find("%d") != std::string::npos)
                g.trace_path.replace(g.trace_path.find("%d"), 2, std::to_string(device));
            if (cudaEventCreate(&g.trace_base) == cudaSuccess && cudaEventRecord(g.trace_base, g.copy[0]) == cudaSuccess
                && cudaEventSynchronize(g.trace_base) == cudaSuccess)
                {
                g.trace_t0 = std::chrono::steady_clock::now();
                g.trace_on = true;
                }
            else
                cudaGetLastError();
            }
    g.inited = true;
    return 0;
    }

namespace
    {
void readers_release();
void cache_release_all();
    }

void dev_shutdown()
    {
    if (!g.inited)
        return;
    dev_drain();
    stop_threads();
    trace_dump();
    // an arena a writable handle is still assembling stays alive (its chunks are referenced by that handle)
    std::vector<ArenaFrame*> keep;
    for (ArenaFrame* f : g.frames)
        {
        if (f->assembling)
            {
            keep.push_back(f);
            continue;
            }
        for (auto& b : f->blocks)
            cudaFree(b.ptr);
        if (f->packed)
            cudaEventDestroy(f->packed);
        for (cudaEvent_t e : f->k1)
            if (e)
                cudaEventDestroy(e);
        delete f;
        }
    g.frames.swap(keep);
    readers_release();
    cache_release_all();
    if (g.scratch)
        cudaFree(g.scratch);
    g.scratch = nullptr;
    g.scratch_bytes = 0;
    sort_release_workspace();
    }

bool dev_is_device_pointer(const void* p)
    {
    if (p == nullptr || g.failed)
        return false;
    if (!g.inited && !dev_cuda_available())
        {
        g.failed = true;
        return false;
        }
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess)
        {
        cudaGetLastError();
        return false;
        }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
    }

void dev_set_user_stream(void* s) { g.user = (cudaStream_t)s; }
void* dev_user_stream() { return (void*)g.user; }

void dev_file_stage_config(int* writers, int* pwrite_threads, int* mode)
    {
    if (!g.inited)
        {
        // same defaults / environment as dev_init, without touching CUDA (host-only tools and tests)
        g.file_mode = file_mode_from_env();
        g.n_writers = default_io_threads();
        if (const char* w = getenv("PGSD_B200_WRITER_THREADS"))
            if (atoi(w) >= 1 && atoi(w) <= 64)
                g.n_writers = (uint32_t)atoi(w);
        if (const char* w = getenv("PGSD_B200_PWRITE_THREADS"))
            if (atoi(w) >= 1 && atoi(w) <= 64)
                g.pwrite_threads = (uint32_t)atoi(w);
        }
    *writers = (int)g.n_writers;
    *pwrite_threads = (int)g.pwrite_threads;
    *mode = (int)g.file_mode;
    }
uint64_t dev_slot_bytes() { return g.slot_bytes; }

int dev_configure_staging(uint32_t n_slots, uint64_t slot_bytes, uint32_t writer_threads)
    {
    if (g.threads_running)
        {
        int rc = dev_drain();
        if (rc != 0)
            return rc;
        stop_threads();
        }
    if (n_slots < 2 || slot_bytes < 4096 || writer_threads < 1 || writer_threads > 64)
        {
        set_last_error("configure_staging: need >= 2 slots of >= 4096 bytes and 1..64 writer threads");
        return -2;
        }
    g.n_slots = n_slots;
    g.slot_bytes = slot_bytes;
    g.n_writers = writer_threads;
    return 0;
    }

// ------------------------------------------------------------------------------ K1 entry points
static int fill_segment(PackSegment& s, void* dst, int dst_type, uint64_t N, uint32_t M, int src_type,
                        const Column* cols)
    {
    if (M == 0 || M > (uint32_t)PACK_MAX_COLS)
        {
        set_last_error("pack: M must be 1..8 for SoA packing");
        return -2;
        }
    memset(&s, 0, sizeof(s));
    s.dst = dst;
    s.N = N;
    s.M = M;
    s.src_type = src_type;
    s.dst_type = dst_type;
    for (uint32_t j = 0; j < M; j++)
        {
        s.base[j] = cols[j].base;
        s.stride[j] = cols[j].stride;
        }
    return 0;
    }

int dev_pack(void* dst_dev, int dst_type, uint64_t N, uint32_t M, int src_type, const Column* cols, void* stream)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    if (!cast_supported(src_type, dst_type) || cols == nullptr || (N > 0 && dst_dev == nullptr))
        {
        set_last_error("pack: invalid argument");
        return -2;
        }
    PackSegment s;
    rc = fill_segment(s, dst_dev, dst_type, N, M, src_type, cols);
    if (rc != 0)
        return rc;
    return pack_launch(&s, 1, (cudaStream_t)stream);
    }

// optional CUDA-event timing of the last K1 launch issued by a frame write (bench.py)
static bool g_pack_prof = false, g_pack_timed = false;
static cudaEvent_t g_pack_ev[2] = { nullptr, nullptr };
void dev_pack_profiling(bool on)
    {
    g_pack_prof = on;
    g_pack_timed = false;
    }
int dev_pack_last_ms(float* ms)
    {
    *ms = 0.f;
    if (!g_pack_timed)
        return -2;
    CUDA_TRY(cudaEventSynchronize(g_pack_ev[1]), -1);
    CUDA_TRY(cudaEventElapsedTime(ms, g_pack_ev[0], g_pack_ev[1]), -1);
    return 0;
    }

int dev_arena_pack(PackRequest* reqs, int n, void** frame_io)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    if (frame_io == nullptr)
        return -2;
    ArenaFrame* f = (ArenaFrame*)*frame_io;
    if (f == nullptr)
        {
        rc = acquire_frame(&f);
        if (rc != 0)
            return rc;
        *frame_io = f;
        }
    bool need_sync = false;
    int i0 = 0;
    while (i0 < n)
        {
        PackSegment segs[PACK_MAX_SEGS];
        int ns = 0;
        for (; i0 < n && ns < PACK_MAX_SEGS; i0++)
            {
            PackRequest& r = reqs[i0];
            r.arena_ptr = nullptr;
            if (!cast_supported(r.src_type, r.dst_type) || r.M == 0 || r.M > (uint32_t)PACK_MAX_COLS
                || (r.N > 0 && r.cols == nullptr))
                {
                set_last_error("write_chunk_soa: unsupported dtype cast or M (1..8)");
                return -2;
                }
            if (r.N == 0)
                continue;
            const size_t ds = type_size(r.dst_type), ss = type_size(r.src_type);
            rc = arena_alloc(f, r.N * r.M * ds, &r.arena_ptr);
            if (rc != 0)
                return rc;
            Column dc[PACK_MAX_COLS];
            for (uint32_t j = 0; j < r.M; j++)
                {
                dc[j] = r.cols[j];
                if (r.cols[j].base == nullptr)
                    {
                    set_last_error("write_chunk_soa: NULL column");
                    return -2;
                    }
                if (r.host_columns)
                    {
                    if (r.cols[j].stride < 1)
                        {
                        set_last_error("write_chunk_soa: host columns need a positive stride");
                        return -2;
                        }
                    // stage the host column (its whole strided span) into the arena
                    uint64_t span = ((r.N - 1) * (uint64_t)r.cols[j].stride + 1) * ss;
                    void* tmp = nullptr;
                    rc = arena_alloc(f, span, &tmp);
                    if (rc != 0)
                        return rc;
                    CUDA_TRY(cudaMemcpyAsync(tmp, r.cols[j].base, span, cudaMemcpyHostToDevice, g.user), -1);
                    g_stats.h2d_bytes += span;
                    dc[j].base = tmp;
                    need_sync = true;
                    }
                }
            rc = fill_segment(segs[ns], r.arena_ptr, r.dst_type, r.N, r.M, r.src_type, dc);
            if (rc != 0)
                return rc;
            ns++;
            }
        if (g_pack_prof)
            {
            if (!g_pack_ev[0])
                {
                cudaEventCreate(&g_pack_ev[0]);
                cudaEventCreate(&g_pack_ev[1]);
                }
            cudaEventRecord(g_pack_ev[0], g.user);
            }
        if (g.trace_on)
            cudaEventRecord(f->k1[0], g.user);
        nvtxRangePushA("pgsd K1 pack");
        rc = pack_launch(segs, ns, g.user);
        nvtxRangePop();
        if (rc != 0)
            return rc;
        if (g.trace_on)
            {
            cudaEventRecord(f->k1[1], g.user);
            f->k1_timed = true;
            }
        if (g_pack_prof)
            {
            cudaEventRecord(g_pack_ev[1], g.user);
            g_pack_timed = true;
            }
        }
    if (need_sync)
        {
        // host source buffers may be reused by the caller as soon as we return
        // (ref semantics: pgsd.c:521, :2229)
        CUDA_TRY(cudaStreamSynchronize(g.user), -1);
        }
    return 0;
    }

// ------------------------------------------------------------------------------ K3
int dev_frame_submit(int fd, const WriteJob* jobs, int njobs, void* frame)
    {
    ArenaFrame* f = (ArenaFrame*)frame;
    if (!g.inited || f == nullptr)
        {
        if (njobs == 0)
            return 0;
        set_last_error("frame_submit without a packed frame");
        return -2;
        }
    int rc = start_threads();
    if (rc != 0)
        return rc;
    CUDA_TRY(cudaEventRecord(f->packed, g.user), -1);
    if (g.trace_on && f->k1_timed)
        {
        double t0 = 0, t1 = 0;
        if (cudaEventSynchronize(f->k1[1]) == cudaSuccess && trace_device_interval(f->k1[0], f->k1[1], &t0, &t1))
            {
            std::lock_guard<std::mutex> lk(g.mu);
            g.trace.push_back(TraceEvent { 'K', g.frame_seq, t0, t1, 0, 0 });
            }
        }
        {
        std::lock_guard<std::mutex> lk(g.mu);
        f->seq = g.frame_seq++;
        uint64_t used = 0;
        for (const Block& bl : f->blocks)
            used += bl.used;
        g.last_frame_bytes = used;
        // small frame: all chunks inside one short span of the arena -> one bundled copy
        const char* lo = nullptr;
        const char* hi = nullptr;
        int nz = 0;
        for (int i = 0; i < njobs; i++)
            {
            if (jobs[i].bytes == 0)
                continue;
            const char* a = (const char*)jobs[i].dev_ptr;
            lo = (lo == nullptr || a < lo) ? a : lo;
            hi = (hi == nullptr || a + jobs[i].bytes > hi) ? a + jobs[i].bytes : hi;
            nz++;
            }
        const uint64_t bundle_max = g.slot_bytes < (4ull << 20) ? g.slot_bytes : (4ull << 20);
        if (nz >= 2 && (uint64_t)(hi - lo) <= bundle_max)
            {
            StageJob job { fd, lo, (uint64_t)(hi - lo), 0, f, {} };
            for (int i = 0; i < njobs; i++)
                if (jobs[i].bytes)
                    job.segs.push_back(Seg { (uint64_t)((const char*)jobs[i].dev_ptr - lo), jobs[i].bytes, jobs[i].file_off });
            f->outstanding++;
            g.jobs_outstanding++;
            g.stage_q.push_back(std::move(job));
            }
        else
            for (int i = 0; i < njobs; i++)
                {
                if (jobs[i].bytes == 0)
                    continue;
                f->outstanding++;
                g.jobs_outstanding++;
                g.stage_q.push_back(StageJob { fd, (const char*)jobs[i].dev_ptr, jobs[i].bytes, jobs[i].file_off, f, {} });
                }
        f->assembling = false;
        }
    g.cv_stage.notify_all();
    g.cv_done.notify_all(); // a frame without jobs is free again at once
    return 0;
    }

// Small writes of the file layer (index entries, write buffers, names, header) while the staging threads run: the
// bytes are copied and written by the serial writer thread, in submission order, behind the bundled frames queued
// before them.  false: the threads are not running -- the caller writes synchronously.  Errors surface at dev_drain().
bool dev_async_host_write(int fd, const void* buf, uint64_t n, uint64_t off)
    {
    if (!g.threads_running || n == 0 || n > (4u << 20))
        return false;
    WriteItem wi { -1, fd, off, n, nullptr, 0, {}, {} };
    wi.host_bytes.assign((const char*)buf, (const char*)buf + n);
        {
        std::lock_guard<std::mutex> lk(g.mu);
        g.jobs_outstanding++;
        g.serial_q.push_back(std::move(wi));
        }
    g.cv_serial.notify_one();
    return true;
    }

// a handle gives up the frame it was assembling (close / error paths): the arena may be recycled
void dev_frame_abandon(void* frame)
    {
    ArenaFrame* f = (ArenaFrame*)frame;
    if (f == nullptr)
        return;
        {
        std::lock_guard<std::mutex> lk(g.mu);
        f->assembling = false;
        }
    g.cv_done.notify_all();
    }

int dev_drain()
    {
    if (!g.inited)
        return 0;
    if (g.threads_running)
        {
        auto t0 = std::chrono::steady_clock::now();
        std::unique_lock<std::mutex> lk(g.mu);
        g.cv_done.wait(lk, [] { return g.jobs_outstanding == 0; });
        g_stats.commit_wait_s
            += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
    if (g.io_error.exchange(false))
        {
        set_last_error("a staged file write failed");
        return -1;
        }
    return 0;
    }

int dev_copy_to_host(void* host_dst, const void* dev_src, uint64_t bytes)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    CUDA_TRY(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, g.user), -1);
    CUDA_TRY(cudaStreamSynchronize(g.user), -1);
    g_stats.d2h_bytes += bytes;
    return 0;
    }

// file -> pinned -> device.  Large reads are split into 1 MiB pieces over reader threads (8 by default), each with
// its own pair of pinned buffers and copy stream: page-cache reads of one file scale with threads
// (no exclusive inode lock on the read side) and overlap the H2D DMA of the previous pieces.
namespace
    {
constexpr uint64_t READ_PIECE = 1ull << 20; // small pieces: the pipeline of a read fills in ~0.3 ms
// Reads of a few MiB (one chunk of a 1 Mi-row field; the reference's benchmark-read issues 1700 of 8 MiB) would give
// every reader thread ONE 1-MiB piece: all page-cache copies first, all H2D copies after them (8 MiB over PCIe is
// 0.16 ms -- 40 % on top of the copies).  Such reads are cut into 256-KiB pieces, 8 buffers per thread, so the DMA of a
// thread's first piece runs while it copies its second.
constexpr uint64_t READ_PIECE_SMALL = 256ull << 10, READ_SMALL_BELOW = 32ull << 20;
constexpr int READ_BUFS_MAX = (int)(2 * READ_PIECE / READ_PIECE_SMALL);
constexpr int READ_THREADS_MAX = 32;
int g_read_threads = 8; // PGSD_B200_READER_THREADS (file -> pinned saturates near 35 GB/s from 6 threads up)
struct Reader
    {
    char* buf = nullptr; // 2 * READ_PIECE bytes, page-locked: 2 large or 8 small piece buffers
    cudaEvent_t ev[READ_BUFS_MAX] = { nullptr };
    cudaStream_t st = nullptr;
    };
inline uint64_t read_piece_for(uint64_t bytes) { return bytes < READ_SMALL_BELOW ? READ_PIECE_SMALL : READ_PIECE; }
Reader g_readers[READ_THREADS_MAX];
bool g_readers_ready = false;

int readers_init()
    {
    if (g_readers_ready)
        return 0;
    if (const char* e = getenv("PGSD_B200_READER_THREADS"))
        {
        int v = atoi(e);
        if (v >= 1 && v <= READ_THREADS_MAX)
            g_read_threads = v;
        }
    for (int i = 0; i < g_read_threads; i++)
        {
        Reader& r = g_readers[i];
        CUDA_TRY(cudaStreamCreateWithFlags(&r.st, cudaStreamNonBlocking), -1);
        CUDA_TRY(cudaHostAlloc((void**)&r.buf, 2 * READ_PIECE, cudaHostAllocDefault), -6);
        for (int k = 0; k < READ_BUFS_MAX; k++)
            CUDA_TRY(cudaEventCreateWithFlags(&r.ev[k], cudaEventDisableTiming), -1);
        }
    g_readers_ready = true;
    return 0;
    }

void read_pool_stop();
void readers_release()
    {
    read_pool_stop();
    for (int i = 0; i < READ_THREADS_MAX; i++)
        {
        Reader& r = g_readers[i];
        if (r.buf)
            cudaFreeHost(r.buf);
        r.buf = nullptr;
        for (int k = 0; k < READ_BUFS_MAX; k++)
            {
            if (r.ev[k])
                cudaEventDestroy(r.ev[k]);
            r.ev[k] = nullptr;
            }
        if (r.st)
            cudaStreamDestroy(r.st);
        r.st = nullptr;
        }
    g_readers_ready = false;
    }

// pieces t, t+T, t+2T, ... of the read
bool reader_run(int t, int T, int fd, char* dev_dst, uint64_t bytes, uint64_t file_off)
    {
    cudaSetDevice(g.device);
    Reader& r = g_readers[t];
    const uint64_t piece = read_piece_for(bytes);
    const int nbuf = (int)(2 * READ_PIECE / piece);
    const uint64_t npieces = (bytes + piece - 1) / piece;
    int k = 0;
    bool ok = true;
    for (uint64_t i = (uint64_t)t; ok && i < npieces; i += (uint64_t)T)
        {
        const uint64_t off = i * piece;
        const uint64_t len = bytes - off < piece ? bytes - off : piece;
        char* const buf = r.buf + (uint64_t)k * piece;
        ok = cudaEventSynchronize(r.ev[k]) == cudaSuccess; // the copy that last used this buffer is done
        uint64_t got = 0;
        while (ok && got < len)
            {
            ssize_t n = pread(fd, buf + got, len - got, (off_t)(file_off + off + got));
            if (n < 0 && errno == EINTR)
                continue;
            if (n <= 0)
                ok = false;
            else
                got += (uint64_t)n;
            }
        ok = ok && cudaMemcpyAsync(dev_dst + off, buf, len, cudaMemcpyHostToDevice, r.st) == cudaSuccess
             && cudaEventRecord(r.ev[k], r.st) == cudaSuccess;
        k = k + 1 == nbuf ? 0 : k + 1;
        }
    return cudaStreamSynchronize(r.st) == cudaSuccess && ok;
    }
    } // namespace

// Persistent reader threads: a partitioned read is 1700 calls of 8 MiB in the reference's benchmark-read workload,
// and starting 7 threads per call cost as much as a sixth of the copy itself.
namespace
    {
struct ReadPool
    {
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::vector<std::thread> th;
    uint64_t gen = 0;
    int T = 0, pending = 0;
    bool stop = false, ok = true;
    int fd = -1;
    char* dst = nullptr;
    uint64_t bytes = 0, off = 0;
    };
ReadPool g_rp;

void read_pool_main(int t)
    {
    uint64_t seen = 0;
    for (;;)
        {
        int T, fd;
        char* dst;
        uint64_t bytes, off;
            {
            std::unique_lock<std::mutex> lk(g_rp.mu);
            g_rp.cv_go.wait(lk, [&] { return g_rp.stop || g_rp.gen != seen; });
            if (g_rp.stop)
                return;
            seen = g_rp.gen;
            T = g_rp.T;
            fd = g_rp.fd;
            dst = g_rp.dst;
            bytes = g_rp.bytes;
            off = g_rp.off;
            }
        bool ok = true;
        if (t < T)
            ok = reader_run(t, T, fd, dst, bytes, off);
            {
            std::lock_guard<std::mutex> lk(g_rp.mu);
            if (!ok)
                g_rp.ok = false;
            if (t < T && --g_rp.pending == 0)
                g_rp.cv_done.notify_all();
            }
        }
    }

void ahead_stop();
void read_pool_stop()
    {
    ahead_stop(); // its worker uses the pool: it goes first
    if (g_rp.th.empty())
        return;
        {
        std::lock_guard<std::mutex> lk(g_rp.mu);
        g_rp.stop = true;
        }
    g_rp.cv_go.notify_all();
    for (auto& x : g_rp.th)
        x.join();
    g_rp.th.clear();
    g_rp.stop = false;
    }
    } // namespace

static void read_pool_atexit()
    {
    static bool hook = false;
    if (!hook)
        {
        atexit(read_pool_stop);
        hook = true;
        }
    }

static int read_file_to_device_now(int fd, void* dev_dst, uint64_t bytes, uint64_t file_off)
    {
    int rc = 0;
    static std::mutex read_mu; // the reader contexts are shared: one read at a time (read-ahead and prefetch threads call us)
    std::lock_guard<std::mutex> read_lk(read_mu);
    rc = readers_init();
    if (rc != 0)
        return rc;
    const uint64_t piece = read_piece_for(bytes);
    const uint64_t npieces = (bytes + piece - 1) / piece;
    const int T = npieces < (uint64_t)g_read_threads ? (int)npieces : g_read_threads;
    bool ok = true;
    if (T == 1)
        ok = reader_run(0, 1, fd, (char*)dev_dst, bytes, file_off);
    else
        {
        if (g_rp.th.empty())
            {
            read_pool_atexit();
            for (int t = 1; t < g_read_threads; t++)
                g_rp.th.emplace_back(read_pool_main, t);
            }
            {
            std::lock_guard<std::mutex> lk(g_rp.mu);
            g_rp.T = T;
            g_rp.fd = fd;
            g_rp.dst = (char*)dev_dst;
            g_rp.bytes = bytes;
            g_rp.off = file_off;
            g_rp.pending = T - 1;
            g_rp.ok = true;
            g_rp.gen++;
            }
        g_rp.cv_go.notify_all();
        const bool ok0 = reader_run(0, T, fd, (char*)dev_dst, bytes, file_off);
        std::unique_lock<std::mutex> lk(g_rp.mu);
        g_rp.cv_done.wait(lk, [] { return g_rp.pending == 0; });
        ok = ok0 && g_rp.ok;
        }
    if (!ok)
        {
        set_last_error("reading the file into device memory failed (pread or H2D copy)");
        cudaGetLastError();
        return -1;
        }
    g_stats.h2d_bytes += bytes;
    g_stats.file_bytes_read += bytes;
    return 0;
    }

// ---- read-ahead ------------------------------------------------------------------------------------------------------
// A partitioned read of a trajectory is a long run of equally sized reads at a constant file stride (the reference's
// benchmark-read: 1700 reads of 8 MiB, one per key and frame, benchmark-read.cc:46-120; a rank's row slice of one
// field over the frames of a trajectory).  One such call costs the wake-up of the reader threads + the page-cache copy
// + the tail of the H2D, and nothing overlaps the caller's own work between calls.  After three reads with the same
// size and stride the next two are fetched ahead by a worker (same reader threads, same pinned pieces) into device
// staging buffers; a call that finds its range there only pays a device-to-device copy.  The state machine is
// read_ahead.cpp (no CUDA in it: the CPU suite drives it with host memory from several threads); here are its
// operations on device memory.  Only for read-only handles; every open / close of a handle drops what was fetched,
// and a file whose size or mtime changed is not served from staging (chunks of a GSD file are never rewritten in
// place, so a stale range needs an outside writer that replaces the file between two reads of one handle).
// OPT-IN: PGSD_B200_READ_AHEAD=1.  It is worth 5-8 % on back-to-back reads (the reader threads are bound by the
// host's page-cache copy either way); its point is to overlap the next read with what the caller does in between.
namespace
    {
bool ra_read_now(int fd, void* dst, uint64_t bytes, uint64_t off) { return read_file_to_device_now(fd, dst, bytes, off) == 0; }
bool ra_alloc(void** p, uint64_t bytes)
    {
    if (cudaMalloc(p, bytes) == cudaSuccess)
        return true;
    cudaGetLastError();
    *p = nullptr;
    return false;
    }
void ra_release(void* p) { cudaFree(p); }
bool ra_copy(void* dst, const void* src, uint64_t bytes) // callers are serialised (ReadAhead::read)
    {
    static cudaStream_t st = nullptr;
    if (st == nullptr && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess)
        {
        st = nullptr;
        cudaGetLastError();
        return false;
        }
    if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess)
        return true;
    cudaGetLastError();
    return false;
    }
void ra_thread_init() { cudaSetDevice(g.device); }
ReadAhead g_ra(ReadAheadOps { ra_read_now, ra_alloc, ra_release, ra_copy, ra_thread_init });
// its worker uses the reader pool: read_pool_stop() (atexit, dev_shutdown) stops it first
const bool g_ra_hooked = (g_ra.at_worker_start(read_pool_atexit), true);
void ahead_stop() { g_ra.stop(); }
    } // namespace

void dev_read_ahead_reset() { g_ra.reset(); }
void dev_read_ahead_stats(uint64_t* hits, uint64_t* issued, uint64_t* dropped) { g_ra.stats(hits, issued, dropped); }

int dev_read_file_to_device(int fd, void* dev_dst, uint64_t bytes, uint64_t file_off, bool read_only)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    if (bytes == 0)
        return 0;
    const char* ea = getenv("PGSD_B200_READ_AHEAD"); // opt-in, read per call
    if (!read_only || ea == nullptr || ea[0] != '1')
        return read_file_to_device_now(fd, dev_dst, bytes, file_off);
    return g_ra.read(fd, dev_dst, bytes, file_off) ? 0 : -1;
    }

// ------------------------------------------------------------------------------ K2
__global__ void k2_scan_sizes(const unsigned long long* __restrict__ sizes, int P, int C, int rank,
                              unsigned long long* __restrict__ out /* C x {excl,total,max,first} */)
    {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C)
        return;
    unsigned long long excl = 0, total = 0, mx = 0;
    for (int r = 0; r < P; r++)
        {
        unsigned long long v = sizes[(size_t)r * C + c];
        if (r < rank)
            excl += v;
        total += v;
        mx = v > mx ? v : mx;
        }
    out[4 * (size_t)c + 0] = excl;
    out[4 * (size_t)c + 1] = total;
    out[4 * (size_t)c + 2] = mx;
    out[4 * (size_t)c + 3] = sizes[c];
    }

static_assert(sizeof(SizeScan) == 32, "SizeScan must be 4 x u64");

static int scratch_reserve(size_t bytes)
    {
    if (g.scratch_bytes >= bytes)
        return 0;
    if (g.scratch)
        cudaFree(g.scratch);
    g.scratch = nullptr;
    g.scratch_bytes = 0;
    size_t cap = (bytes + (1u << 20) - 1) / (1u << 20) * (1u << 20);
    CUDA_TRY(cudaMalloc((void**)&g.scratch, cap), -6);
    g.scratch_bytes = cap;
    return 0;
    }

int dev_scan_sizes(const uint64_t* sizes, int P, int C, int rank, SizeScan* out)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    if (P <= 0 || C < 0 || rank < 0 || rank >= P || (C > 0 && (!sizes || !out)))
        return -2;
    if (C == 0)
        return 0;
    size_t in_b = (size_t)P * C * 8, out_b = (size_t)C * 32;
    unsigned long long* d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d, in_b + out_b), -6);
    cudaMemcpyAsync(d, sizes, in_b, cudaMemcpyHostToDevice, g.aux);
    k2_scan_sizes<<<(C + 127) / 128, 128, 0, g.aux>>>(d, P, C, rank, d + (size_t)P * C);
    g_stats.kernel_launches++;
    cudaMemcpyAsync(out, d + (size_t)P * C, out_b, cudaMemcpyDeviceToHost, g.aux);
    cudaError_t e = cudaStreamSynchronize(g.aux);
    cudaFree(d);
    if (e != cudaSuccess)
        {
        set_last_error(std::string("scan_sizes: ") + cudaGetErrorString(e));
        return -1;
        }
    return 0;
    }

// ------------------------------------------------------------------------------ NCCL transport
namespace
    {
struct NcclApi
    {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    };
NcclApi nccl;

bool nccl_load(std::string& err)
    {
    if (nccl.lib)
        return true;
    const char* env = getenv("PGSD_B200_NCCL_LIB");
    const char* names[] = { env, "libnccl.so.2", "libnccl.so" };
    for (const char* nm : names)
        {
        if (!nm || !*nm)
            continue;
        nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (nccl.lib)
            break;
        }
    if (!nccl.lib)
        {
        err = std::string("cannot load NCCL: ") + (dlerror() ? dlerror() : "?");
        return false;
        }
    nccl.GetUniqueId = (decltype(nccl.GetUniqueId))dlsym(nccl.lib, "ncclGetUniqueId");
    nccl.CommInitRank = (decltype(nccl.CommInitRank))dlsym(nccl.lib, "ncclCommInitRank");
    nccl.CommDestroy = (decltype(nccl.CommDestroy))dlsym(nccl.lib, "ncclCommDestroy");
    nccl.AllGather = (decltype(nccl.AllGather))dlsym(nccl.lib, "ncclAllGather");
    nccl.GetErrorString = (decltype(nccl.GetErrorString))dlsym(nccl.lib, "ncclGetErrorString");
    if (!nccl.GetUniqueId || !nccl.CommInitRank || !nccl.CommDestroy || !nccl.AllGather || !nccl.GetErrorString)
        {
        err = "NCCL library lacks a required symbol";
        nccl.lib = nullptr;
        return false;
        }
    return true;
    }

class NcclComm : public Comm
    {
    public:
    enum
        {
        MAX_WORDS = 32768 // u64 values per rank per collective (the bucket counts of a distributed reorder: <= 16385)
        };
    ncclComm_t comm = nullptr;
    cudaStream_t st = nullptr;
    unsigned long long* d_send = nullptr; // MAX_WORDS
    unsigned long long* d_recv = nullptr; // nprocs * MAX_WORDS
    unsigned long long* d_scan = nullptr; // MAX_WORDS * 4
    unsigned long long* h_pin = nullptr;  // pinned: max(nprocs, 4) * MAX_WORDS

    ~NcclComm() override
        {
        if (comm)
            nccl.CommDestroy(comm);
        if (d_send)
            cudaFree(d_send);
        if (h_pin)
            cudaFreeHost(h_pin);
        if (st)
            cudaStreamDestroy(st);
        }

    int gather_piece(const uint64_t* send, size_t n)
        {
        memcpy(h_pin, send, n * 8);
        if (cudaMemcpyAsync(d_send, h_pin, n * 8, cudaMemcpyHostToDevice, st) != cudaSuccess)
            return -1;
        ncclResult_t r = nccl.AllGather(d_send, d_recv, n, ncclUint64, comm, st);
        if (r != ncclSuccess)
            {
            set_last_error(std::string("ncclAllGather: ") + nccl.GetErrorString(r));
            return -1;
            }
        g_collectives++;
        return 0;
        }

    int allgather(const uint64_t* send, uint64_t* recv, size_t n) override
        {
        size_t done = 0;
        while (done < n)
            {
            size_t k = n - done < (size_t)MAX_WORDS ? n - done : (size_t)MAX_WORDS;
            if (gather_piece(send + done, k) != 0)
                return -1;
            if (cudaMemcpyAsync(h_pin, d_recv, (size_t)nprocs * k * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess
                || cudaStreamSynchronize(st) != cudaSuccess)
                {
                set_last_error("nccl allgather: device copy failed");
                cudaGetLastError();
                return -1;
                }
            for (int r = 0; r < nprocs; r++)
                memcpy(recv + (size_t)r * n + done, h_pin + (size_t)r * k, k * 8);
            done += k;
            }
        return 0;
        }

    // K2: the gathered [P][n] matrix never leaves the device; only the scan result comes back
    int allgather_scan(const uint64_t* send, SizeScan* out, size_t n) override
        {
        size_t done = 0;
        while (done < n)
            {
            size_t k = n - done < (size_t)MAX_WORDS ? n - done : (size_t)MAX_WORDS;
            if (gather_piece(send + done, k) != 0)
                return -1;
            k2_scan_sizes<<<(unsigned)((k + 127) / 128), 128, 0, st>>>(d_recv, nprocs, (int)k, rank, d_scan);
            g_stats.kernel_launches++;
            if (cudaMemcpyAsync(h_pin, d_scan, k * 32, cudaMemcpyDeviceToHost, st) != cudaSuccess
                || cudaStreamSynchronize(st) != cudaSuccess)
                {
                set_last_error("nccl allgather_scan: device copy failed");
                cudaGetLastError();
                return -1;
                }
            memcpy(out + done, h_pin, k * 32);
            done += k;
            }
        return 0;
        }
    };
    } // namespace

int nccl_unique_id(void* out128, std::string& err)
    {
    if (!nccl_load(err))
        return -1;
    ncclUniqueId id;
    ncclResult_t r = nccl.GetUniqueId(&id);
    if (r != ncclSuccess)
        {
        err = std::string("ncclGetUniqueId: ") + nccl.GetErrorString(r);
        return -1;
        }
    memcpy(out128, &id, 128);
    return 0;
    }

Comm* make_nccl_comm(int rank, int nprocs, const void* unique_id, int device, std::string& err)
    {
    if (!nccl_load(err))
        return nullptr;
    if (dev_init(device) != 0)
        {
        err = last_error();
        return nullptr;
        }
    NcclComm* c = new NcclComm;
    c->rank = rank;
    c->nprocs = nprocs;
    c->kind = CommKind::Nccl;
    ncclUniqueId id;
    memcpy(&id, unique_id, 128);
    ncclResult_t r = nccl.CommInitRank(&c->comm, nprocs, id, rank);
    if (r != ncclSuccess)
        {
        err = std::string("ncclCommInitRank: ") + nccl.GetErrorString(r);
        c->comm = nullptr;
        delete c;
        return nullptr;
        }
    size_t words = (size_t)NcclComm::MAX_WORDS;
    size_t pin_words = words * (size_t)(nprocs > 4 ? nprocs : 4);
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess
        || cudaMalloc((void**)&c->d_send, (words + words * nprocs + words * 4) * 8) != cudaSuccess
        || cudaHostAlloc((void**)&c->h_pin, pin_words * 8, cudaHostAllocDefault) != cudaSuccess)
        {
        err = "NCCL transport: CUDA allocation failed";
        cudaGetLastError();
        delete c;
        return nullptr;
        }
    c->d_recv = c->d_send + words;
    c->d_scan = c->d_recv + words * nprocs;
    return c;
    }

// ------------------------------------------------------------------------------ reorder (K4 + K5)
int dev_reorder(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm, int nfields,
                const ReorderField* fields, void* stream)
    {
    nvtxRangePushA("pgsd K4+K5 reorder");
    const int rc = dev_reorder_rows(n, keys, keys_sorted, perm, nfields, fields, stream);
    nvtxRangePop();
    return rc;
    }

// ------------------------------------------------------------------------------ reorder, host buffers
int dev_reorder_host(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm, int nfields,
                     const ReorderField* fields)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    if (n == 0)
        return 0;
    if (keys == nullptr || nfields < 0 || (nfields > 0 && fields == nullptr))
        return -2;
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    size_t need = 3 * up(n * 4);
    for (int i = 0; i < nfields; i++)
        {
        if (fields[i].row_bytes == 0 || fields[i].in == nullptr || fields[i].out == nullptr)
            return -2;
        need += 2 * up(n * (size_t)fields[i].row_bytes);
        }
    rc = scratch_reserve(need);
    if (rc != 0)
        return rc;
    cudaStream_t st = g.user;
    char* p = g.scratch;
    uint32_t* d_keys = (uint32_t*)p;
    p += up(n * 4);
    uint32_t* d_sorted = (uint32_t*)p;
    p += up(n * 4);
    uint32_t* d_perm = (uint32_t*)p;
    p += up(n * 4);
    std::vector<ReorderField> df((size_t)nfields);
    CUDA_TRY(cudaMemcpyAsync(d_keys, keys, n * 4, cudaMemcpyHostToDevice, st), -1);
    g_stats.h2d_bytes += n * 4;
    for (int i = 0; i < nfields; i++)
        {
        size_t b = n * (size_t)fields[i].row_bytes;
        df[i].in = p;
        p += up(b);
        df[i].out = p;
        p += up(b);
        df[i].row_bytes = fields[i].row_bytes;
        CUDA_TRY(cudaMemcpyAsync((void*)df[i].in, fields[i].in, b, cudaMemcpyHostToDevice, st), -1);
        g_stats.h2d_bytes += b;
        }
    rc = dev_reorder_rows(n, d_keys, d_sorted, perm ? d_perm : nullptr, nfields, df.data(), st);
    if (rc != 0)
        return rc;
    if (keys_sorted)
        {
        CUDA_TRY(cudaMemcpyAsync(keys_sorted, d_sorted, n * 4, cudaMemcpyDeviceToHost, st), -1);
        g_stats.d2h_bytes += n * 4;
        }
    if (perm)
        {
        CUDA_TRY(cudaMemcpyAsync(perm, d_perm, n * 4, cudaMemcpyDeviceToHost, st), -1);
        g_stats.d2h_bytes += n * 4;
        }
    for (int i = 0; i < nfields; i++)
        {
        size_t b = n * (size_t)fields[i].row_bytes;
        CUDA_TRY(cudaMemcpyAsync(fields[i].out, df[i].out, b, cudaMemcpyDeviceToHost, st), -1);
        g_stats.d2h_bytes += b;
        }
    CUDA_TRY(cudaStreamSynchronize(st), -1);
    return 0;
    }

// ------------------------------------------------------------------------------ helpers
// Caching allocator behind pgsd_b200_malloc/free: per-frame device arrays of the read path
// (6 chunks in, 6 out, ids) are recycled instead of paying cudaMalloc + the device-wide
// synchronisation of cudaFree every frame.  Sizes are rounded up to 2 MiB; at most 16 GiB is cached.
namespace
    {
std::mutex g_cache_mu;
std::map<uint64_t, std::vector<void*>> g_cache_free;
std::map<void*, uint64_t> g_cache_size;
uint64_t g_cache_bytes = 0;
constexpr uint64_t CACHE_MAX = 16ull << 30;
uint64_t cache_round(uint64_t b)
    {
    if (b == 0)
        b = 1;
    const uint64_t q = b < (1ull << 20) ? 512 : (2ull << 20);
    return (b + q - 1) / q * q;
    }
void cache_release_all()
    {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (auto& kv : g_cache_free)
        for (void* p : kv.second)
            {
            cudaFree(p);
            g_cache_size.erase(p);
            }
    g_cache_free.clear();
    g_cache_bytes = 0;
    }
    } // namespace

int dev_malloc(void** p, uint64_t bytes)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    const uint64_t sz = cache_round(bytes);
        {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto it = g_cache_free.find(sz);
        if (it != g_cache_free.end() && !it->second.empty())
            {
            *p = it->second.back();
            it->second.pop_back();
            g_cache_bytes -= sz;
            return 0;
            }
        }
    if (cudaMalloc(p, sz) != cudaSuccess)
        {
        cudaGetLastError();
        cache_release_all(); // give cached blocks back and retry once
        CUDA_TRY(cudaMalloc(p, sz), -6);
        }
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache_size[*p] = sz;
    return 0;
    }
int dev_free(void* p)
    {
    if (!p)
        return 0;
    if (g.inited)
        dev_init(-1); // binds this host thread to the rank's GPU
    // the block may be handed out again at once: everything queued on the caller's stream that
    // could still touch it must have finished (what cudaFree guarantees implicitly)
    CUDA_TRY(cudaStreamSynchronize(g.user), -1);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_cache_size.find(p);
    if (it == g_cache_size.end())
        {
        CUDA_TRY(cudaFree(p), -1);
        return 0;
        }
    if (g_cache_bytes + it->second > CACHE_MAX)
        {
        g_cache_size.erase(it);
        CUDA_TRY(cudaFree(p), -1);
        return 0;
        }
    g_cache_free[it->second].push_back(p);
    g_cache_bytes += it->second;
    return 0;
    }
int dev_host_alloc(void** p, uint64_t bytes)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    CUDA_TRY(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault), -6);
    return 0;
    }
int dev_host_free(void* p)
    {
    if (p)
        CUDA_TRY(cudaFreeHost(p), -1);
    return 0;
    }
int dev_memcpy(void* dst, const void* src, uint64_t bytes, int kind)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, k, g.user), -1);
    CUDA_TRY(cudaStreamSynchronize(g.user), -1);
    if (kind == 1)
        g_stats.h2d_bytes += bytes;
    if (kind == 2)
        g_stats.d2h_bytes += bytes;
    return 0;
    }
int dev_synchronize()
    {
    if (g.inited)
        dev_init(-1); // binds this host thread to the rank's GPU
    if (!g.inited)
        return 0;
    CUDA_TRY(cudaDeviceSynchronize(), -1);
    return 0;
    }

struct Timer
    {
    cudaEvent_t a, b;
    };
int dev_timer_create(void** t)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    Timer* x = new Timer;
    if (cudaEventCreate(&x->a) != cudaSuccess || cudaEventCreate(&x->b) != cudaSuccess)
        {
        delete x;
        return -1;
        }
    *t = x;
    return 0;
    }
int dev_timer_start(void* t)
    {
    CUDA_TRY(cudaEventRecord(((Timer*)t)->a, g.user), -1);
    return 0;
    }
int dev_timer_stop(void* t, float* ms)
    {
    Timer* x = (Timer*)t;
    CUDA_TRY(cudaEventRecord(x->b, g.user), -1);
    CUDA_TRY(cudaEventSynchronize(x->b), -1);
    CUDA_TRY(cudaEventElapsedTime(ms, x->a, x->b), -1);
    return 0;
    }
int dev_timer_destroy(void* t)
    {
    Timer* x = (Timer*)t;
    if (x)
        {
        cudaEventDestroy(x->a);
        cudaEventDestroy(x->b);
        delete x;
        }
    return 0;
    }

// write a buffer larger than L2 (126 MB) so the next timed launch starts cold
int dev_flush_l2()
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    static void* buf = nullptr;
    const size_t bytes = 256u << 20;
    if (!buf)
        CUDA_TRY(cudaMalloc(&buf, bytes), -6);
    static int v = 0;
    CUDA_TRY(cudaMemsetAsync(buf, ++v & 0xff, bytes, g.user), -1);
    return 0;
    }
} // namespace pgsdb
