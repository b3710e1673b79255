// read_ahead.h -- read-ahead of equally sized, constant-stride reads of one file into (device) memory.  The state
// machine knows nothing about CUDA: the memory it stages into and the way a range is read are given as operations, so
// the CPU test-suite can drive it with host memory from several threads (tests/test_read_ahead_host.py through
// pgsd_b200_read_ahead_host_read) while device.cu drives it with device staging buffers and the reader threads.
//
// Replaces nothing in the reference (its pgsd_read_chunk, pgsd.c:2436-2537, is one blocking MPI_File_read_at per call;
// the access pattern is the reference's benchmark-read.cc:46-120).  Opt-in: PGSD_B200_READ_AHEAD=1.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <sys/stat.h>
#include <thread>

namespace pgsdb
{
struct ReadAheadOps
    {
    bool (*read_now)(int fd, void* dst, uint64_t bytes, uint64_t off); // blocking: file range -> memory of this kind
    bool (*alloc)(void** p, uint64_t bytes);                           // staging buffer of this kind
    void (*release)(void* p);
    bool (*copy)(void* dst, const void* src, uint64_t bytes);          // staging -> destination, done on return
    void (*thread_init)();                                             // once in the worker thread (may be NULL)
    };

class ReadAhead
    {
    public:
    explicit ReadAhead(const ReadAheadOps& ops) : m_ops(ops) { }
    ~ReadAhead() { stop(); }
    static constexpr uint64_t MIN_BYTES = 256ull << 10, MAX_BYTES = 64ull << 20;

    // The read of [off, off + bytes) of fd into dst: from staging when the range was fetched ahead, else by
    // ops.read_now; keeps the pattern and queues what should be fetched next.  Callers are serialised.  false: the read failed.
    bool read(int fd, void* dst, uint64_t bytes, uint64_t off);
    void reset(); // forget the file: nothing queued, nothing running, descriptor closed (every open / close of a handle)
    void stop();  // reset + worker joined + staging released
    void stats(uint64_t* hits, uint64_t* issued, uint64_t* dropped);
    void at_worker_start(void (*hook)()) { m_on_start = hook; } // called once before the worker thread is created

    private:
    enum State { FREE = 0, QUEUED, RUNNING, READY, FAILED, COPYING };
    struct Slot
        {
        uint64_t off = 0, bytes = 0, cap = 0, seq = 0;
        void* mem = nullptr;
        int state = FREE;
        };
    static constexpr int SLOTS = 3;
    void worker();
    void forget(std::unique_lock<std::mutex>& lk);
    bool same_file(const struct stat& st) const;

    ReadAheadOps m_ops;
    void (*m_on_start)() = nullptr;
    std::mutex m_front; // serialises read() / reset(): taken before m_mu, never by the worker
    std::mutex m_mu;
    std::condition_variable m_work, m_done;
    std::thread m_th;
    bool m_running = false, m_stop = false;
    int m_fd = -1; // our own descriptor of the file being read ahead
    dev_t m_dev = 0;
    ino_t m_ino = 0;
    int64_t m_size = 0;
    struct timespec m_mtime = { 0, 0 };
    bool m_have_last = false;
    uint64_t m_last_off = 0, m_last_bytes = 0, m_seq = 0;
    int64_t m_stride = 0;
    int m_streak = 0;
    Slot m_slot[SLOTS];
    uint64_t m_hits = 0, m_issued = 0, m_dropped = 0;
    };
} // namespace pgsdb
