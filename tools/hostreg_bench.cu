// hostreg_bench.cu -- can the GPU DMA straight into page-cache pages?  Maps a new tmpfs/ext4 file,
// populates it, cudaHostRegister()s the mapping and copies device memory into it; prints the time
// of every step.  Measurement tool only.
//   nvcc -O2 -o hostreg_bench hostreg_bench.cu -lpthread ; ./hostreg_bench /dev/shm/x.bin 2048 8 [chunk_MiB]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <thread>
#include <unistd.h>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv)
    {
    const char* path = argv[1];
    size_t total = (size_t)atol(argv[2]) << 20;
    int T = atoi(argv[3]);
    size_t chunk = argc > 4 ? (size_t)atol(argv[4]) << 20 : total;
    void* dev;
    cudaMalloc(&dev, total);
    cudaMemset(dev, 0x5a, total);
    cudaDeviceSynchronize();
    for (int rep = 0; rep < 2; rep++)
        {
        int fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
        double t0 = now();
        if (ftruncate(fd, (off_t)total) != 0)
            perror("ftruncate");
        char* m = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        double t1 = now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t]() {
                size_t per = (total / T + 4095) & ~(size_t)4095;
                size_t a = per * t, b = a + per > total ? total : a + per;
                if (a < b && madvise(m + a, b - a, 23) != 0)
                    for (size_t i = a; i < b; i += 4096)
                        m[i] = 0;
            });
        for (auto& x : th)
            x.join();
        double t2 = now(), treg = 0, tcpy = 0, tunreg = 0;
        for (size_t off = 0; off < total; off += chunk)
            {
            size_t len = total - off < chunk ? total - off : chunk;
            double a = now();
            cudaError_t e = cudaHostRegister(m + off, len, cudaHostRegisterDefault);
            if (e != cudaSuccess)
                {
                printf("cudaHostRegister failed: %s\n", cudaGetErrorString(e));
                return 1;
                }
            double b = now();
            cudaMemcpy(m + off, (char*)dev + off, len, cudaMemcpyDeviceToHost);
            double c = now();
            cudaHostUnregister(m + off);
            double d = now();
            treg += b - a;
            tcpy += c - b;
            tunreg += d - c;
            }
        double t3 = now();
        bool ok = m[0] == 0x5a && m[total - 1] == 0x5a;
        munmap(m, total);
        close(fd);
        double t4 = now();
        printf("rep %d total=%zuMiB T=%d chunk=%zuMiB: map %.3f populate %.3f (%.1f GB/s) register %.3f copy %.3f (%.1f GB/s) unregister %.3f unmap %.3f => %.2f GB/s end to end, data %s\n",
               rep, total >> 20, T, chunk >> 20, t1 - t0, t2 - t1, total / (t2 - t1) / 1e9, treg, tcpy, total / tcpy / 1e9, tunreg,
               t4 - t3, total / (t4 - t0) / 1e9, ok ? "ok" : "BAD");
        unlink(path);
        }
    return 0;
    }
