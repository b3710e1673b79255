// pwrite_bench.cpp -- host-side ceiling of the K3 file stage: T threads pwrite()ing fixed-size
// pieces of one new file (page cache / tmpfs), the way libpgsd_b200's writer threads do.
//   pwrite_bench <path> <total_MiB> <piece_MiB> <threads> [mode]
//   mode: pwrite | mmap | falloc | mmap_piece[_pop] | falloc_piece[_pop] (whole piece allocated with one fallocate
//   before it is mapped; _pop: MADV_POPULATE_WRITE before the copy)
// Prints one line: threads piece mode GB/s.  Measurement tool only; not part of the library.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <thread>
#include <unistd.h>
#include <vector>

int main(int argc, char** argv)
    {
    if (argc < 5)
        return 2;
    const char* path = argv[1];
    size_t total = (size_t)atol(argv[2]) << 20, piece = (size_t)atol(argv[3]) << 20;
    int T = atoi(argv[4]);
    std::string mode = argc > 5 ? argv[5] : "pwrite";
    int fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (fd < 0)
        {
        perror(path);
        return 1;
        }
    std::vector<char*> src((size_t)T);
    for (int t = 0; t < T; t++)
        {
        src[(size_t)t] = (char*)aligned_alloc(4096, piece);
        memset(src[(size_t)t], t + 1, piece);
        }
    char* map = nullptr;
    auto t0 = std::chrono::steady_clock::now();
    if (mode == "falloc")
        {
        if (posix_fallocate(fd, 0, (off_t)total) != 0)
            perror("fallocate");
        }
    const bool populate = (mode == "mmap_pop");
    const bool bulk_falloc = (mode == "falloc_piece" || mode == "falloc_piece_pop");
    const bool piecewise = (mode == "mmap_piece" || mode == "mmap_piece_pop" || bulk_falloc);
    if (mode == "mmap" || mode == "mmap_pop")
        {
        if (ftruncate(fd, (off_t)total) != 0)
            perror("ftruncate");
        map = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        }
    std::atomic<size_t> next { 0 };
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t]() {
            for (;;)
                {
                size_t off = next.fetch_add(piece);
                if (off >= total)
                    return;
                size_t len = total - off < piece ? total - off : piece;
                if (piecewise)
                    {
                    // what a writer thread of the library would do: extend, map the piece, copy, unmap
                    if (bulk_falloc ? fallocate(fd, 0, (off_t)off, (off_t)len) != 0 : fallocate(fd, 0, (off_t)(off + len - 1), 1) != 0)
                        perror("fallocate");
                    char* m = (char*)mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)off);
                    if (mode == "mmap_piece_pop" || mode == "falloc_piece_pop")
                        madvise(m, len, 23 /* MADV_POPULATE_WRITE */);
                    memcpy(m, src[(size_t)t], len);
                    munmap(m, len);
                    }
                else if (map)
                    {
                    if (populate)
                        madvise(map + off, len, 23 /* MADV_POPULATE_WRITE */);
                    memcpy(map + off, src[(size_t)t], len);
                    }
                else
                    {
                    size_t done = 0;
                    while (done < len)
                        {
                        ssize_t k = pwrite(fd, src[(size_t)t] + done, len - done, (off_t)(off + done));
                        if (k <= 0)
                            {
                            perror("pwrite");
                            return;
                            }
                        done += (size_t)k;
                        }
                    }
                }
        });
    for (auto& x : th)
        x.join();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("threads=%d piece=%zuMiB mode=%s total=%zuMiB  %.2f GB/s\n", T, piece >> 20, mode.c_str(), total >> 20,
           total / s / 1e9);
    close(fd);
    unlink(path);
    return 0;
    }
