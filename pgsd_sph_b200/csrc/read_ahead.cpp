// read_ahead.cpp -- see read_ahead.h.  After three reads with the same size and file stride the next two ranges are
// fetched by a worker into staging buffers; a read that finds its range there pays a copy out of staging.  What was
// fetched is dropped when the handle's file changes identity (device, inode, size, mtime) and at every reset().
//
// Concurrency rules (the first version broke the second one: pgsd.hoomd's frame-prefetch thread and the main thread
// used the front end at the same time, one dropped the range the other was waiting for, and the waiter never woke
// -- an 8-GPU bench run stalled at the join of that thread, profiles/r5_bench_n8_stalled.err):
//   * read() and reset() are serialised by m_front; only they change what is queued or drop what is ready;
//   * the worker only moves a slot QUEUED -> RUNNING -> READY / FAILED;
//   * every wait is bounded: a caller that does not get its range in time reads it itself.
#include "read_ahead.h"

#include <chrono>
#include <unistd.h>

namespace pgsdb
{
void ReadAhead::worker()
    {
    if (m_ops.thread_init)
        m_ops.thread_init();
    for (;;)
        {
        Slot* s = nullptr;
        int fd = -1;
            {
            std::unique_lock<std::mutex> lk(m_mu);
            m_work.wait(lk, [&] {
                if (m_stop)
                    return true;
                for (Slot& x : m_slot)
                    if (x.state == QUEUED)
                        return true;
                return false;
            });
            if (m_stop)
                return;
            for (Slot& x : m_slot)
                if (x.state == QUEUED && (s == nullptr || x.seq < s->seq))
                    s = &x;
            s->state = RUNNING;
            fd = m_fd;
            }
        const bool ok = m_ops.read_now(fd, s->mem, s->bytes, s->off);
            {
            std::lock_guard<std::mutex> lk(m_mu);
            s->state = ok ? READY : FAILED;
            }
        m_done.notify_all();
        }
    }

void ReadAhead::forget(std::unique_lock<std::mutex>& lk)
    {
    for (Slot& x : m_slot)
        if (x.state == QUEUED)
            x.state = FREE;
    const bool idle = m_done.wait_for(lk, std::chrono::seconds(60), [&] {
        for (Slot& x : m_slot)
            if (x.state == RUNNING || x.state == COPYING)
                return false;
        return true;
    });
    for (Slot& x : m_slot)
        {
        if (x.state == READY)
            m_dropped++;
        if (x.state != RUNNING && x.state != COPYING)
            x.state = FREE;
        }
    if (m_fd >= 0 && idle) // a read that is still running keeps its descriptor (leaked rather than closed under it)
        close(m_fd);
    m_fd = -1;
    m_have_last = false;
    m_streak = 0;
    }

bool ReadAhead::same_file(const struct stat& st) const
    {
    return m_fd >= 0 && st.st_dev == m_dev && st.st_ino == m_ino && (int64_t)st.st_size == m_size
           && st.st_mtim.tv_sec == m_mtime.tv_sec && st.st_mtim.tv_nsec == m_mtime.tv_nsec;
    }

void ReadAhead::reset()
    {
    std::lock_guard<std::mutex> front(m_front);
    std::unique_lock<std::mutex> lk(m_mu);
    if (m_fd >= 0)
        forget(lk);
    }

void ReadAhead::stop()
    {
    std::lock_guard<std::mutex> front(m_front);
        {
        std::unique_lock<std::mutex> lk(m_mu);
        if (!m_running && m_fd < 0 && m_slot[0].mem == nullptr)
            return;
        forget(lk);
        m_stop = true;
        }
    m_work.notify_all();
    if (m_running)
        m_th.join();
    std::lock_guard<std::mutex> lk(m_mu);
    m_running = false;
    m_stop = false;
    for (Slot& x : m_slot)
        {
        if (x.mem && x.state != RUNNING && x.state != COPYING)
            m_ops.release(x.mem);
        x = Slot();
        }
    }

void ReadAhead::stats(uint64_t* hits, uint64_t* issued, uint64_t* dropped)
    {
    std::lock_guard<std::mutex> lk(m_mu);
    if (hits)
        *hits = m_hits;
    if (issued)
        *issued = m_issued;
    if (dropped)
        *dropped = m_dropped;
    }

bool ReadAhead::read(int fd, void* dst, uint64_t bytes, uint64_t file_off)
    {
    struct stat st;
    if (bytes < MIN_BYTES || bytes > MAX_BYTES || fstat(fd, &st) != 0)
        return m_ops.read_now(fd, dst, bytes, file_off);

    std::lock_guard<std::mutex> front(m_front);
    bool served = false;
        {
        std::unique_lock<std::mutex> lk(m_mu);
        if (!same_file(st))
            {
            forget(lk);
            m_fd = dup(fd);
            m_dev = st.st_dev;
            m_ino = st.st_ino;
            m_size = (int64_t)st.st_size;
            m_mtime = st.st_mtim;
            }
        // (1) fetched ahead?
        Slot* hit = nullptr;
        for (Slot& x : m_slot)
            if ((x.state == QUEUED || x.state == RUNNING || x.state == READY) && x.off == file_off && x.bytes == bytes)
                hit = &x;
        if (hit)
            {
            // bounded: whatever goes wrong with the worker, the caller falls back to reading the range itself
            m_done.wait_for(lk, std::chrono::seconds(5), [&] { return hit->state == READY || hit->state == FAILED; });
            if (hit->state != READY && hit->state != FAILED)
                hit = nullptr; // still queued or being read: left to the worker, dropped by a later call
            }
        if (hit)
            {
            if (hit->state == READY)
                {
                hit->state = COPYING; // neither matched nor recycled meanwhile
                lk.unlock();
                served = m_ops.copy(dst, hit->mem, bytes);
                lk.lock();
                if (served)
                    m_hits++;
                }
            hit->state = FREE;
            m_done.notify_all();
            }
        // (2) the pattern: same size, same stride
        if (m_have_last && bytes == m_last_bytes)
            {
            const int64_t d = (int64_t)file_off - (int64_t)m_last_off;
            if (d == m_stride && d != 0)
                m_streak++;
            else
                {
                m_stride = d;
                m_streak = d != 0 ? 1 : 0;
                }
            }
        else
            {
            m_stride = 0;
            m_streak = 0;
            }
        m_have_last = true;
        m_last_off = file_off;
        m_last_bytes = bytes;
        // (3) what should be in flight now: the next two ranges of the pattern; everything else is dropped
        uint64_t want[2] = { 0, 0 };
        int nwant = 0;
        if (m_streak >= 2 && m_fd >= 0)
            for (int k = 1; k <= 2; k++)
                {
                const int64_t t = (int64_t)file_off + k * m_stride;
                if (t < 0 || (uint64_t)t + bytes > (uint64_t)m_size)
                    break;
                want[nwant++] = (uint64_t)t;
                }
        for (Slot& x : m_slot)
            {
            if (x.state == FREE || x.state == RUNNING || x.state == COPYING)
                continue;
            bool wanted = false;
            for (int k = 0; k < nwant; k++)
                wanted = wanted || (x.off == want[k] && x.bytes == bytes);
            if (!wanted || x.state == FAILED)
                {
                if (x.state == READY)
                    m_dropped++;
                x.state = FREE;
                }
            }
        bool queued = false;
        for (int k = 0; k < nwant; k++)
            {
            bool have = false;
            for (Slot& x : m_slot)
                have = have || (x.state != FREE && x.off == want[k] && x.bytes == bytes);
            if (have)
                continue;
            Slot* f = nullptr;
            for (Slot& x : m_slot)
                if (x.state == FREE && f == nullptr)
                    f = &x;
            if (f == nullptr)
                break;
            if (f->cap < bytes)
                {
                if (f->mem)
                    m_ops.release(f->mem);
                f->mem = nullptr;
                f->cap = 0;
                if (!m_ops.alloc(&f->mem, bytes))
                    break;
                f->cap = bytes;
                }
            f->off = want[k];
            f->bytes = bytes;
            f->seq = ++m_seq;
            f->state = QUEUED;
            m_issued++;
            queued = true;
            }
        if (queued)
            {
            if (!m_running)
                {
                if (m_on_start)
                    m_on_start();
                m_th = std::thread(&ReadAhead::worker, this);
                m_running = true;
                }
            m_work.notify_one();
            }
        }
    if (served)
        return true;
    return m_ops.read_now(fd, dst, bytes, file_off);
    }
} // namespace pgsdb
