// atomic_overlap_bench.cu -- do the cursor atomics of k6_slot_scatter (pgsd_sph_b200/csrc/kernels_slot.cu) and the
// streaming traffic of the same pass overlap in the memory system, or do they queue on one resource?
//   A: n atomicAdd(+1, result used) on `nb` cursors, one per 128-byte line, bucket = hash(row) -- what the scatter issues
//   C: a copy of `bytes` bytes (16-byte loads and stores) -- its 40 + 40 B/row of payload
//   F: both in ONE kernel, every thread interleaving its share of A with its share of C
// and A and C launched together on two streams.  Each alone, then together; "sum" and "max" are what serial and
// perfectly overlapped execution would give.  Development tool, not part of the library.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/atomic_overlap_bench tools/atomic_overlap_bench.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t hash(uint32_t x)
    {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
    }

__global__ void __launch_bounds__(512) k_atomics(uint32_t* cursor, uint32_t nb, uint32_t cstride, uint64_t n, uint32_t* sink)
    {
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        acc += atomicAdd(cursor + (size_t)(hash((uint32_t)i) % nb) * cstride, 1u);
    if (acc == 0xffffffffu)
        *sink = acc;
    }

__global__ void __launch_bounds__(512) k_copy(const uint4* __restrict__ in, uint4* __restrict__ out, uint64_t n16)
    {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = __ldg(in + i);
    }

// one atomic per `per` 16-byte copies (40-byte rows: 5 copies of 16 B move 40 B in + 40 B out -> per = 2.5; use 5 : 2)
__global__ void __launch_bounds__(512) k_fused(const uint4* __restrict__ in, uint4* __restrict__ out, uint64_t n16, uint32_t* cursor,
                                              uint32_t nb, uint32_t cstride, uint32_t* sink)
    {
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i * 5 < n16; i += (uint64_t)gridDim.x * blockDim.x)
        {
        acc += atomicAdd(cursor + (size_t)(hash((uint32_t)(2 * i)) % nb) * cstride, 1u);
        acc += atomicAdd(cursor + (size_t)(hash((uint32_t)(2 * i + 1)) % nb) * cstride, 1u);
        const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
        const uint64_t total5 = (n16 + 4) / 5;
#pragma unroll
        for (int k = 0; k < 5; k++)
            {
            const uint64_t j = i + (uint64_t)k * total5;
            if (j < n16)
                out[j] = __ldg(in + j);
            }
        (void)stride;
        }
    if (acc == 0xffffffffu)
        *sink = acc;
    }

#define CK(x)                                                                                      \
    do                                                                                             \
        {                                                                                          \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess)                                                                     \
            {                                                                                      \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                               \
            return 1;                                                                              \
            }                                                                                      \
        } while (0)

int main(int argc, char** argv)
    {
    const uint64_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : (16ull << 20);
    const uint32_t nb = 16384, cstride = 32;
    const uint64_t bytes = n * 40, n16 = bytes / 16;
    uint32_t *cursor, *sink;
    uint4 *in, *out;
    CK(cudaMalloc(&cursor, (size_t)nb * cstride * 4));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMalloc(&in, bytes));
    CK(cudaMalloc(&out, bytes));
    CK(cudaMemset(cursor, 0, (size_t)nb * cstride * 4));
    CK(cudaMemset(in, 1, bytes));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, f0, f1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventCreate(&f0));
    CK(cudaEventCreate(&f1));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto best = [&](auto launch, int reps)
        {
        float b = 1e9f;
        for (int r = 0; r < reps; r++)
            {
            cudaDeviceSynchronize();
            cudaEventRecord(e0, s1);
            launch();
            cudaEventRecord(e1, s1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            b = std::min(b, ms);
            }
        return b;
        };
    for (int ctas_per_sm : { 2, 4 })
        {
        const int grid = sms * ctas_per_sm;
        const float ta = best([&] { k_atomics<<<grid, 512, 0, s1>>>(cursor, nb, cstride, n, sink); }, 5);
        const float tc = best([&] { k_copy<<<grid, 512, 0, s1>>>(in, out, n16); }, 5);
        const float tf = best([&] { k_fused<<<grid, 512, 0, s1>>>(in, out, n16, cursor, nb, cstride, sink); }, 5);
        // two streams: half of the CTAs each, started together; the time is until both are done
        const float t2 = best(
            [&]
                {
                cudaEventRecord(f0, s1);
                cudaStreamWaitEvent(s2, f0, 0);
                k_atomics<<<grid / 2, 512, 0, s1>>>(cursor, nb, cstride, n, sink);
                k_copy<<<grid / 2, 512, 0, s2>>>(in, out, n16);
                cudaEventRecord(f1, s2);
                cudaStreamWaitEvent(s1, f1, 0);
                },
            5);
        const float ta_h = best([&] { k_atomics<<<grid / 2, 512, 0, s1>>>(cursor, nb, cstride, n, sink); }, 5);
        const float tc_h = best([&] { k_copy<<<grid / 2, 512, 0, s1>>>(in, out, n16); }, 5);
        printf("n=%llu rows, %d CTAs/SM x 512 threads: atomics alone %.3f ms (%.0f G/s) | copy of %.2f GB alone %.3f ms (%.0f GB/s) | "
               "fused in one kernel %.3f ms (sum %.3f, max %.3f) | two streams, half the CTAs each: %.3f ms (alone at half: atomics %.3f, "
               "copy %.3f)\n",
               (unsigned long long)n, ctas_per_sm, ta, n / ta / 1e6, 2.0 * bytes / 1e9, tc, 2.0 * bytes / tc / 1e6, tf, ta + tc,
               std::max(ta, tc), t2, ta_h, tc_h);
        }
    return 0;
    }
