// file_stage.cpp -- host side of K3 (see file_stage.h).  Replaces the reference's blocking MPI_File_write_at of a
// rank's chunk bytes (/root/reference/pgsd/pgsd/pgsd.c:2229, :1154) for bytes that arrive from the device in pieces.
//
// Why two ways to write a piece (measured on the B200 box, 16 cores; profiles/r1_pwrite_sweep_*.txt,
// profiles/r3_pwrite_sweep2_*.txt):
//   * Buffered pwrite()s to ONE file serialise on the inode lock: 3.6-4.1 GB/s on tmpfs for 1..32 threads.  Copies
//     through per-piece shared mappings insert page-cache pages from all writer threads in parallel: 7.7 GB/s at
//     8 threads, 8.6 at 16.  Allocating a piece with one fallocate and / or MADV_POPULATE_WRITE first is slower
//     (6.3 / 4.1 GB/s), so the range is only allocated up front when the file system is short of space.
//   * On ext4 the order is the other way round: pwrite 5.5-6.7 GB/s (best with 1-2 threads), mappings 4.2 GB/s.
// Hence FileMode::Auto: mappings on tmpfs only, pwrite everywhere else (also the only safe choice on network
// file systems, where stores into a shared mapping of a file other hosts write to are not coherent).
//
// Page ownership: the stager cuts a job so that interior piece boundaries fall on page boundaries
// (file_first_piece_len), and file_write_piece maps only the whole pages of its range; the partial page at the
// start or end of a rank's chunk region is written with pwrite by both neighbours.  So no page-cache page is
// written through two mappings, or through a mapping and a pwrite.  Round 1 mapped arbitrary byte ranges, including
// pieces of a few bytes that shared their page with other writers' pieces and with rank 0's index writes; one run
// of a 2-rank random script produced a different file in that version and the cause was never isolated.  The
// present scheme removes the whole class (tests/test_file_stage.py: multi-process stress, 0 mismatches).
#include "file_stage.h"

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/statvfs.h>
#include <sys/vfs.h>
#include <thread>
#include <unistd.h>
#include <vector>

namespace pgsdb
{
FileMode file_mode_from_env()
    {
    const char* m = getenv("PGSD_B200_FILE_MODE");
    if (m == nullptr || *m == 0 || !strcmp(m, "auto"))
        return FileMode::Auto;
    if (!strcmp(m, "pwrite"))
        return FileMode::Pwrite;
    if (!strcmp(m, "mmap"))
        return FileMode::Mmap;
    return FileMode::Auto;
    }

uint64_t file_page_size()
    {
    static const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE);
    return page;
    }

bool file_is_tmpfs(int fd)
    {
    struct statfs sf;
    if (fstatfs(fd, &sf) != 0)
        return false;
    return (unsigned long)sf.f_type == 0x01021994ul; // TMPFS_MAGIC
    }

uint64_t file_first_piece_len(uint64_t off, uint64_t bytes, uint64_t piece)
    {
    const uint64_t page = file_page_size();
    uint64_t first = piece - (off & (page - 1)); // piece is a multiple of the page size (or smaller than a page)
    if (piece < page || (piece & (page - 1)) != 0)
        first = piece;
    return bytes < first ? bytes : first;
    }

static bool pwrite_range(int fd, const char* p, uint64_t off, uint64_t len, uint64_t* left)
    {
    while (len > 0)
        {
        ssize_t k = pwrite(fd, p, len, (off_t)off);
        if (k < 0)
            {
            if (errno == EINTR)
                continue;
            *left = len;
            return false;
            }
        p += k;
        off += (uint64_t)k;
        len -= (uint64_t)k;
        }
    *left = 0;
    return true;
    }

bool file_write_piece(int fd, const char* p, uint64_t off, uint64_t len, bool use_mmap, uint64_t* left)
    {
    *left = len;
    if (len == 0)
        {
        *left = 0;
        return true;
        }
    const uint64_t page = file_page_size();
    const uint64_t end = off + len;
    const uint64_t a = (off + page - 1) & ~(page - 1); // first whole page
    const uint64_t b = end & ~(page - 1);              // end of the last whole page
    // below 256 KiB of whole pages the mapping's own cost (mmap + munmap + TLB shootdown) eats the gain
    if (!use_mmap || b <= a || b - a < (256u << 10))
        return pwrite_range(fd, p, off, len, left);

    // the partial pages at either end: pwrite (their other part belongs to a neighbour who does the same)
    uint64_t l = 0;
    if (a > off && !pwrite_range(fd, p, off, a - off, &l))
        {
        *left = len - (a - off) + l;
        return false;
        }
    if (end > b && !pwrite_range(fd, p + (b - off), b, end - b, &l))
        {
        *left = (b - a) + l;
        return false;
        }
    // whole pages [a, b): make sure they exist, then copy through a mapping of exactly these pages
    struct stat st;
    bool mapped = false;
    if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode))
        {
        bool ok = true;
        // A store into a mapping cannot report ENOSPC (it raises SIGBUS).  With plenty of room only the file size
        // is advanced (1-byte allocation at the end of the range: never shrinks what another writer extended);
        // when the file system is getting full the whole range is allocated first, which returns the error here.
        struct statvfs vfs;
        const bool roomy = fstatvfs(fd, &vfs) == 0 && (uint64_t)vfs.f_bavail * vfs.f_frsize > 8 * len + (1ull << 30);
        if (!roomy)
            ok = fallocate(fd, 0, (off_t)a, (off_t)(b - a)) == 0;
        else if ((uint64_t)st.st_size < b)
            ok = fallocate(fd, 0, (off_t)(b - 1), 1) == 0;
        if (ok)
            {
            void* m = mmap(nullptr, (size_t)(b - a), PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)a);
            if (m != MAP_FAILED)
                {
                memcpy(m, p + (a - off), (size_t)(b - a));
                munmap(m, (size_t)(b - a));
                mapped = true;
                }
            }
        }
    if (!mapped && !pwrite_range(fd, p + (a - off), a, b - a, &l))
        {
        *left = l;
        return false;
        }
    *left = 0;
    return true;
    }

double file_stage_ceiling(int fd, uint64_t off, uint64_t bytes, uint64_t piece, int threads, bool use_mmap)
    {
    if (threads < 1 || piece == 0)
        return -1.0;
    std::vector<char*> src((size_t)threads, nullptr);
    for (int t = 0; t < threads; t++)
        {
        if (posix_memalign((void**)&src[(size_t)t], 4096, piece) != 0)
            return -1.0;
        memset(src[(size_t)t], 0x30 + t, piece);
        }
    // same cuts as the stager makes
    std::vector<std::pair<uint64_t, uint64_t>> pieces;
    uint64_t done = 0;
    while (done < bytes)
        {
        const uint64_t len = done == 0 ? file_first_piece_len(off, bytes, piece) : (bytes - done < piece ? bytes - done : piece);
        pieces.emplace_back(off + done, len);
        done += len;
        }
    std::atomic<size_t> next { 0 };
    std::atomic<bool> ok { true };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&, t]() {
            for (;;)
                {
                const size_t i = next.fetch_add(1);
                if (i >= pieces.size())
                    return;
                uint64_t left = 0;
                if (!file_write_piece(fd, src[(size_t)t], pieces[i].first, pieces[i].second, use_mmap, &left))
                    ok = false;
                }
        });
    for (auto& x : th)
        x.join();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (char* p : src)
        free(p);
    return ok.load() ? s : -1.0;
    }
} // namespace pgsdb
