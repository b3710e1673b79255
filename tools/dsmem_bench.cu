// dsmem_bench.cu -- what distributed shared memory delivers for the record traffic of k7_cluster_place
// (pgsd_sph_b200/csrc/kernels_cluster.cu): clusters of 8 CTAs, every CTA stores `iters` rounds of records into the
// shared memory of pseudo-random CTAs of its cluster.  Development tool, not part of the library.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/dsmem_bench tools/dsmem_bench.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cl_map(uint32_t a, uint32_t r)
    {
    uint32_t o;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r));
    return o;
    }
__device__ __forceinline__ void cl_sync()
    {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
__device__ __forceinline__ uint32_t ctarank()
    {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
    }
__device__ __forceinline__ uint32_t hash(uint32_t x)
    {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
    }

// mode: lane bytes LB in {4, 8, 16}; GROUP lanes write one contiguous "record" of GROUP * LB bytes at a random
// slot (slot pitch = pitch bytes) of a random (remote=1), own (remote=0) or fixed-neighbour (remote=2) CTA.
template <int LB>
__global__ void __launch_bounds__(1024) k_store(int group, uint32_t pitch, uint32_t nslots, int iters, int remote,
                                                unsigned long long* cycles)
    {
    extern __shared__ __align__(128) unsigned char buf[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t me = ctarank();
    const uint32_t base = smem_u32(buf);
    const int G = 32 / group;
    const int g = lane / group, c = lane - g * group;
    const bool act = g < G;
    cl_sync();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++)
        {
        const uint32_t h = hash((uint32_t)(it * 1024 * 64 + (blockIdx.x * 32 + w) * 8 + g) * 2654435761u + 12345u);
        const uint32_t slot = h % nslots;
        uint32_t owner = (h >> 20) & 7u;
        if (remote == 0)
            owner = me;
        else if (remote == 2)
            owner = (me + 1) & 7u;
        const uint32_t a = cl_map(base + slot * pitch + (uint32_t)c * LB, owner);
        if (act)
            {
            if (LB == 4)
                asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(a), "r"(h) : "memory");
            else if (LB == 8)
                asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(h), "r"(h) : "memory");
            else
                asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(h), "r"(h), "r"(h), "r"(h) : "memory");
            }
        }
    cl_sync();
    const long long t1 = clock64();
    if (tid == 0)
        cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    }

// bulk copies shared::cta -> shared::cluster (TMA engine): every CTA sends `chunks` pieces of `bytes` to each peer
__global__ void __launch_bounds__(1024) k_bulk(uint32_t bytes, int rounds, unsigned long long* cycles)
    {
    extern __shared__ __align__(128) unsigned char buf[];
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x;
    const uint32_t me = ctarank();
    const uint32_t half = 96 * 1024; // [0, half): send area, [half, 2 half): receive area
    const uint32_t base = smem_u32(buf), b = smem_u32(&bar);
    if (tid == 0)
        {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    cl_sync();
    const long long t0 = clock64();
    uint32_t parity = 0;
    for (int r = 0; r < rounds; r++)
        {
        if (tid == 0)
            {
            unsigned long long st;
            asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 %0, [%1], %2;" : "=l"(st) : "r"(b), "r"(bytes * 8u) : "memory");
            }
        __syncthreads();
        if (tid < 8)
            {
            const uint32_t peer = (uint32_t)tid;
            const uint32_t dst = cl_map(base + half + me * bytes, peer);
            const uint32_t rbar = cl_map(b, peer);
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "r"(base + peer * bytes), "r"(bytes), "r"(rbar)
                         : "memory");
            }
        // wait for my 8 incoming pieces
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(b), "r"(parity)
                         : "memory");
        parity ^= 1;
        cl_sync(); // nobody overwrites a receive area that is still being read (nothing reads here, keeps rounds apart)
        }
    const long long t1 = clock64();
    if (tid == 0)
        cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    }

template <typename K, typename... A>
static float launch(K k, int nclusters, size_t smem, A... args)
    {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nclusters * 8);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 8;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaLaunchKernelEx(&cfg, k, args...); // warm-up
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, args...);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess)
        printf("  launch failed: %s\n", cudaGetErrorString(e));
    return ms;
    }

int main()
    {
    unsigned long long* cyc;
    cudaMalloc(&cyc, 8 * 4096);
    const size_t smem = 192 * 1024;
    int ncl = 0;
        {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(8 * 64);
        cfg.blockDim = dim3(1024);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 8;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaFuncSetAttribute(k_store<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveClusters(&ncl, k_store<8>, &cfg);
        printf("max active clusters of 8 CTAs x 1024 threads x 192 KB: %d\n", ncl);
        }
    if (ncl <= 0)
        ncl = 16;
    const int iters = 2000;
    struct Case
        {
        const char* name;
        int lb, group;
        uint32_t pitch, nslots;
        int remote;
        };
    const Case cases[] = {
        { "4 B lanes, 10 per 40 B record, random CTA", 4, 10, 40, 4096, 1 },
        { "8 B lanes, 5 per 40 B record, random CTA", 8, 5, 40, 4096, 1 },
        { "8 B lanes, 5 per 40 B record, own CTA", 8, 5, 40, 4096, 0 },
        { "8 B lanes, 5 per 40 B record, next CTA", 8, 5, 40, 4096, 2 },
        { "16 B lanes, 3 per 48 B record, random CTA", 16, 3, 48, 4000, 1 },
        { "16 B lanes, 2 per 32 B record, random CTA", 16, 2, 32, 4096, 1 },
        { "16 B lanes, 4 per 64 B record, random CTA", 16, 4, 64, 2048, 1 },
        { "16 B lanes, 8 per 128 B record, random CTA", 16, 8, 128, 1024, 1 },
        { "8 B lanes, 16 per 128 B record, random CTA", 8, 16, 128, 1024, 1 },
        { "4 B lanes, 32 per 128 B record, random CTA", 4, 32, 128, 1024, 1 },
        { "16 B lanes, 32 per 512 B record, random CTA", 16, 32, 512, 256, 1 },
        { "16 B lanes, 32 per 512 B record, own CTA", 16, 32, 512, 256, 0 },
        { "4 B lanes, 1 per 4 B record, random CTA", 4, 1, 4, 32768, 1 },
        { "8 B lanes, 1 per 8 B record, random CTA", 8, 1, 8, 16384, 1 },
        { "16 B lanes, 1 per 16 B record, random CTA", 16, 1, 16, 8192, 1 },
    };
    for (const Case& c : cases)
        {
        float ms;
        if (c.lb == 4)
            ms = launch(k_store<4>, ncl, smem, c.group, c.pitch, c.nslots, iters, c.remote, cyc);
        else if (c.lb == 8)
            ms = launch(k_store<8>, ncl, smem, c.group, c.pitch, c.nslots, iters, c.remote, cyc);
        else
            ms = launch(k_store<16>, ncl, smem, c.group, c.pitch, c.nslots, iters, c.remote, cyc);
        unsigned long long h[8];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const int G = 32 / c.group;
        const double recs = (double)iters * 32 * G;               // per CTA
        const double bytes = recs * c.group * c.lb;
        printf("%-48s: %8.0f cycles/CTA  %6.2f B/cyc/SM  %6.2f cyc/record  %6.3f cyc/warp-store  (%.3f ms, %d clusters)\n", c.name,
               (double)h[0], bytes / (double)h[0], (double)h[0] / recs, (double)h[0] / ((double)iters * 32), ms, ncl);
        }
    for (uint32_t bytes : { 2048u, 8192u, 11264u })
        {
        const int rounds = 200;
        float ms = launch(k_bulk, ncl, smem, bytes, rounds, cyc);
        unsigned long long h[8];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("bulk smem->peer smem, 8 x %5u B per CTA per round      : %8.0f cycles/CTA  %6.2f B/cyc/SM sent  (%.3f ms)\n", bytes,
               (double)h[0], (double)rounds * 8 * bytes / (double)h[0], ms);
        }
    return 0;
    }
