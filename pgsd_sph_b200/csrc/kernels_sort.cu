// kernels_sort.cu -- K4 (stable LSD radix sort of particle ids) and K5 (permutation gather).
//
// Oracle definition of the operation (BASELINE.json north_star):
//     o = numpy.argsort(ids, kind='stable');  out_f = in_f[o]  for every field f
// applied to the arrays pgsd.hoomd's frame decode returns in file (rank) order
// (/root/reference/pgsd/pgsd/hoomd.py:724-902 never sorts; README.md:29).
//
// K4: 8-bit digits, least significant first.  One pre-pass builds the four global digit
//     histograms so passes whose digit is constant are skipped (dense ids < 2^24 -> 3 passes).
//     Each pass = tile histogram (shared-memory atomics) -> per-digit row scan -> scatter.
//     The scatter ranks keys with warp-level __match_any_sync/popc against per-warp
//     shared-memory digit counters, re-orders the tile in shared memory and writes
//     digit runs out coalesced.  Everything is order preserving => the sort is stable.
// K5: out[i] = in[perm[i]] for all fields in one launch; each warp owns 32 consecutive
//     output rows and moves them word by word so that stores are fully coalesced and the
//     words of one source row are fetched by adjacent lanes.
//
// sm_100a only.  No CPU fallback: callers fail when CUDA is unavailable.
#include "device_internal.h"

namespace pgsdb
{
namespace
    {
constexpr int RADIX = 256;
constexpr int SORT_THREADS = 512;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS; // 8192 keys per CTA

__device__ __forceinline__ unsigned lanemask_lt()
    {
    unsigned m;
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
    }

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p)
    {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
    }

// ---- pre-pass: four global 256-bin histograms (one per key byte) ------------------------------
__global__ void __launch_bounds__(512) k4_digit_census(const uint32_t* __restrict__ keys, uint64_t n,
                                                      unsigned long long* __restrict__ census)
    {
    __shared__ unsigned int h[4][RADIX];
    for (int i = threadIdx.x; i < 4 * RADIX; i += blockDim.x)
        (&h[0][0])[i] = 0;
    __syncthreads();
    const uint64_t n4 = n / 4;
    const uint4* k4 = reinterpret_cast<const uint4*>(keys);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (uint64_t)gridDim.x * blockDim.x)
        {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(k4 + i));
        uint32_t a[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int j = 0; j < 4; j++)
            {
            atomicAdd(&h[0][a[j] & 255u], 1u);
            atomicAdd(&h[1][(a[j] >> 8) & 255u], 1u);
            atomicAdd(&h[2][(a[j] >> 16) & 255u], 1u);
            atomicAdd(&h[3][a[j] >> 24], 1u);
            }
        }
    if (blockIdx.x == 0)
        for (uint64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x)
            {
            uint32_t a = keys[i];
            atomicAdd(&h[0][a & 255u], 1u);
            atomicAdd(&h[1][(a >> 8) & 255u], 1u);
            atomicAdd(&h[2][(a >> 16) & 255u], 1u);
            atomicAdd(&h[3][a >> 24], 1u);
            }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * RADIX; i += blockDim.x)
        {
        unsigned int c = (&h[0][0])[i];
        if (c)
            atomicAdd(census + i, (unsigned long long)c);
        }
    }

// ---- upsweep: digit histogram of every tile; counts[d * ntiles + tile] ------------------------
__global__ void __launch_bounds__(SORT_THREADS) k4_tile_histogram(const uint32_t* __restrict__ keys,
                                                                  uint64_t n, int shift, uint32_t ntiles,
                                                                  uint32_t* __restrict__ counts)
    {
    __shared__ unsigned int h[SORT_WARPS / 4][RADIX]; // 4 sub-histograms to thin out contention
    for (int i = threadIdx.x; i < (SORT_WARPS / 4) * RADIX; i += SORT_THREADS)
        (&h[0][0])[i] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * SORT_TILE;
    unsigned int* mine = h[(threadIdx.x >> 5) & (SORT_WARPS / 4 - 1)];
    if (base + SORT_TILE <= n)
        {
        const uint4* k4 = reinterpret_cast<const uint4*>(keys + base);
#pragma unroll
        for (int k = 0; k < SORT_ITEMS / 4; k++)
            {
            uint4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "l"(k4 + threadIdx.x + k * SORT_THREADS));
            atomicAdd(&mine[(v.x >> shift) & 255u], 1u);
            atomicAdd(&mine[(v.y >> shift) & 255u], 1u);
            atomicAdd(&mine[(v.z >> shift) & 255u], 1u);
            atomicAdd(&mine[(v.w >> shift) & 255u], 1u);
            }
        }
    else
        {
        for (uint64_t i = base + threadIdx.x; i < n; i += SORT_THREADS)
            atomicAdd(&mine[(keys[i] >> shift) & 255u], 1u);
        }
    __syncthreads();
    for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS)
        {
        unsigned int c = 0;
#pragma unroll
        for (int s = 0; s < SORT_WARPS / 4; s++)
            c += h[s][d];
        counts[(size_t)d * ntiles + blockIdx.x] = c;
        }
    }

// ---- scan: one CTA per digit row, exclusive scan over tiles in place; row total out -----------
__global__ void __launch_bounds__(256) k4_row_scan(uint32_t* __restrict__ counts, uint32_t ntiles,
                                                   unsigned long long* __restrict__ row_total)
    {
    __shared__ uint32_t warp_sum[8];
    __shared__ uint32_t carry_s;
    uint32_t* row = counts + (size_t)blockIdx.x * ntiles;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0)
        carry_s = 0;
    __syncthreads();
    for (uint32_t start = 0; start < ntiles; start += 256)
        {
        uint32_t i = start + threadIdx.x;
        uint32_t v = i < ntiles ? row[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
            {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o)
                x += y;
            }
        if (lane == 31)
            warp_sum[w] = x;
        __syncthreads();
        uint32_t wbase = 0;
        for (int j = 0; j < w; j++)
            wbase += warp_sum[j];
        uint32_t carry = carry_s;
        if (i < ntiles)
            row[i] = carry + wbase + x - v;
        __syncthreads();
        if (threadIdx.x == 255)
            carry_s = carry + wbase + x;
        __syncthreads();
        }
    if (threadIdx.x == 0)
        row_total[blockIdx.x] = carry_s;
    }

// digit_base[d] = sum of row_total[d' < d]
__global__ void __launch_bounds__(256) k4_digit_base(const unsigned long long* __restrict__ row_total,
                                                     unsigned long long* __restrict__ digit_base)
    {
    __shared__ unsigned long long s[RADIX];
    s[threadIdx.x] = row_total[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0)
        {
        unsigned long long acc = 0;
        for (int d = 0; d < RADIX; d++)
            {
            unsigned long long c = s[d];
            s[d] = acc;
            acc += c;
            }
        }
    __syncthreads();
    digit_base[threadIdx.x] = s[threadIdx.x];
    }

// ---- downsweep: stable scatter of one tile ---------------------------------------------------
// FIRST: the index payload is implicit (idx = global position), saving its read.
template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS)
    k4_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ idx_in,
               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ idx_out, uint64_t n, int shift,
               uint32_t ntiles, const uint32_t* __restrict__ tile_offset,
               const unsigned long long* __restrict__ digit_base)
    {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* skeys = reinterpret_cast<uint32_t*>(smem_raw);              // SORT_TILE
    uint32_t* sidx = skeys + SORT_TILE;                                    // SORT_TILE
    uint32_t* whist = sidx + SORT_TILE;                                    // SORT_WARPS * RADIX
    uint32_t* dstart = whist + SORT_WARPS * RADIX;                         // RADIX  (tile-local start of digit run)
    unsigned long long* gdelta
        = reinterpret_cast<unsigned long long*>(dstart + RADIX);           // RADIX  (global - local)
    __shared__ uint32_t wtot[8];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint64_t tile_base = (uint64_t)blockIdx.x * SORT_TILE;
    const uint32_t tile_n = (uint32_t)((n - tile_base) < (uint64_t)SORT_TILE ? (n - tile_base) : SORT_TILE);

    for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS)
        whist[i] = 0;
    __syncthreads();

    // element e = w * (32*ITEMS) + k * 32 + lane : tile order == global order
    uint32_t key[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
    const uint32_t wbase = (uint32_t)w * (32 * SORT_ITEMS);
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++)
        {
        uint32_t e = wbase + k * 32 + lane;
        key[k] = e < tile_n ? ld_stream_u32(keys_in + tile_base + e) : 0xffffffffu;
        }
    uint32_t* myhist = whist + w * RADIX;
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++)
        {
        uint32_t e = wbase + k * 32 + lane;
        bool valid = e < tile_n;
        unsigned vm = __ballot_sync(0xffffffffu, valid);
        rank[k] = 0;
        if (valid)
            {
            uint32_t d = (key[k] >> shift) & 255u;
            unsigned peers = __match_any_sync(vm, d);
            int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader)
                {
                old = myhist[d];
                myhist[d] = old + __popc(peers);
                }
            old = __shfl_sync(peers, old, leader);
            rank[k] = old + __popc(peers & lt);
            }
        __syncwarp();
        }
    __syncthreads();

    // per digit: exclusive scan over the warps, digit total, then exclusive scan over digits
    uint32_t total = 0;
    if (tid < RADIX)
        {
#pragma unroll
        for (int j = 0; j < SORT_WARPS; j++)
            {
            uint32_t c = whist[j * RADIX + tid];
            whist[j * RADIX + tid] = total;
            total += c;
            }
        uint32_t x = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
            {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o)
                x += y;
            }
        if (lane == 31)
            wtot[w] = x;
        total = x - total; // exclusive within the warp, completed below
        }
    __syncthreads();
    if (tid < RADIX)
        {
        uint32_t b = 0;
        for (int j = 0; j < w; j++)
            b += wtot[j];
        uint32_t start = total + b;
        dstart[tid] = start;
        gdelta[tid] = digit_base[tid] + (unsigned long long)tile_offset[(size_t)tid * ntiles + blockIdx.x]
                      - (unsigned long long)start;
        }
    __syncthreads();

    // re-order the tile in shared memory
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++)
        {
        uint32_t e = wbase + k * 32 + lane;
        if (e < tile_n)
            {
            uint32_t d = (key[k] >> shift) & 255u;
            uint32_t pos = dstart[d] + myhist[d] + rank[k];
            skeys[pos] = key[k];
            uint32_t src;
            if (FIRST)
                src = (uint32_t)(tile_base + e);
            else
                src = ld_stream_u32(idx_in + tile_base + e);
            sidx[pos] = src;
            }
        }
    __syncthreads();

    // write digit runs out: consecutive j of one digit -> consecutive global addresses
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++)
        {
        uint32_t j = tid + k * SORT_THREADS;
        if (j < tile_n)
            {
            uint32_t kv = skeys[j];
            uint32_t d = (kv >> shift) & 255u;
            unsigned long long g = gdelta[d] + j;
            keys_out[g] = kv;
            idx_out[g] = sidx[j];
            }
        }
    }

constexpr size_t SCATTER_SMEM
    = (size_t)(2 * SORT_TILE + SORT_WARPS * RADIX + RADIX) * sizeof(uint32_t) + RADIX * sizeof(unsigned long long);

__global__ void k4_iota(uint32_t* __restrict__ perm, uint64_t n)
    {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        perm[i] = (uint32_t)i;
    }

// ---- K5 ---------------------------------------------------------------------------------------
constexpr int MAX_GATHER_FIELDS = 16;
struct GatherField
    {
    const unsigned char* in;
    unsigned char* out;
    uint32_t row_bytes;
    uint32_t words; // row_bytes / 4 when the word path may be used, else 0
    };
struct GatherArgs
    {
    GatherField f[MAX_GATHER_FIELDS];
    int nfields;
    };

constexpr int GATHER_THREADS = 256;

__global__ void __launch_bounds__(GATHER_THREADS)
    k5_gather(const uint32_t* __restrict__ perm, uint64_t n, const __grid_constant__ GatherArgs args)
    {
    const int lane = threadIdx.x & 31;
    const uint64_t warps_total = (uint64_t)gridDim.x * (GATHER_THREADS / 32);
    const uint64_t nchunks = (n + 31) / 32;
    for (uint64_t c = (uint64_t)blockIdx.x * (GATHER_THREADS / 32) + (threadIdx.x >> 5); c < nchunks;
         c += warps_total)
        {
        const uint64_t row0 = c * 32;
        const uint32_t rows = (uint32_t)((n - row0) < 32 ? (n - row0) : 32);
        const uint32_t p = lane < rows ? ld_stream_u32(perm + row0 + lane) : 0u;
        for (int fi = 0; fi < args.nfields; fi++)
            {
            const GatherField f = args.f[fi];
            if (f.words == 1)
                {
                if (lane < rows)
                    reinterpret_cast<uint32_t*>(f.out)[row0 + lane]
                        = __ldg(reinterpret_cast<const uint32_t*>(f.in) + p);
                }
            else if (f.words != 0)
                {
                const uint32_t W = f.words;
                const uint32_t total = rows * W;
                const uint32_t* in = reinterpret_cast<const uint32_t*>(f.in);
                uint32_t* out = reinterpret_cast<uint32_t*>(f.out) + row0 * W;
                for (uint32_t q0 = 0; q0 < total; q0 += 32)
                    {
                    uint32_t q = q0 + lane;
                    uint32_t r = q / W;
                    uint32_t comp = q - r * W;
                    uint32_t src = __shfl_sync(0xffffffffu, p, r & 31);
                    if (q < total)
                        out[q] = __ldg(in + (uint64_t)src * W + comp);
                    }
                }
            else
                {
                const uint32_t W = f.row_bytes;
                const uint32_t total = rows * W;
                unsigned char* out = f.out + row0 * W;
                for (uint32_t q0 = 0; q0 < total; q0 += 32)
                    {
                    uint32_t q = q0 + lane;
                    uint32_t r = q / W;
                    uint32_t comp = q - r * W;
                    uint32_t src = __shfl_sync(0xffffffffu, p, r & 31);
                    if (q < total)
                        out[q] = f.in[(uint64_t)src * W + comp];
                    }
                }
            }
        }
    }
    } // namespace

// ---- host side ----------------------------------------------------------------------------------
struct SortWorkspace
    {
    void* base = nullptr;
    size_t bytes = 0;
    };
static SortWorkspace g_sort_ws;
static unsigned long long* g_census_host = nullptr; // pinned, 4*256

static int sort_workspace(size_t need, void** out)
    {
    if (g_sort_ws.bytes < need)
        {
        if (g_sort_ws.base)
            cudaFree(g_sort_ws.base);
        g_sort_ws.base = nullptr;
        g_sort_ws.bytes = 0;
        if (cudaMalloc(&g_sort_ws.base, need) != cudaSuccess)
            {
            set_last_error("cudaMalloc of the sort workspace failed");
            cudaGetLastError();
            return -6;
            }
        g_sort_ws.bytes = need;
        }
    *out = g_sort_ws.base;
    return 0;
    }

void sort_release_workspace()
    {
    if (g_sort_ws.base)
        cudaFree(g_sort_ws.base);
    g_sort_ws = SortWorkspace();
    if (g_census_host)
        cudaFreeHost(g_census_host);
    g_census_host = nullptr;
    }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int dev_sort_ids(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm, void* stream_v)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    if (n >= 0xffffffffull)
        {
        set_last_error("sort_ids: n must be < 2^32 - 1 (uint32 permutation)");
        return -2;
        }
    if (n == 0)
        return 0;
    if (keys == nullptr)
        return -2;
    cudaStream_t st = (cudaStream_t)stream_v;
    static bool attr_done = false;
    if (!attr_done)
        {
        cudaFuncSetAttribute(k4_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM);
        cudaFuncSetAttribute(k4_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM);
        attr_done = true;
        }
    if (!g_census_host && cudaHostAlloc((void**)&g_census_host, 4 * RADIX * 8, cudaHostAllocDefault) != cudaSuccess)
        {
        set_last_error("cudaHostAlloc failed");
        return -6;
        }
    const uint32_t ntiles = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE);
    // workspace: keys A/B, idx A/B, counts[256*ntiles], row_total[256], digit_base[256], census[1024]
    const size_t arr = align_up((size_t)n * 4, 256);
    const size_t counts_b = align_up((size_t)RADIX * ntiles * 4, 256);
    const size_t need = 4 * arr + counts_b + 2 * RADIX * 8 + 4 * RADIX * 8;
    void* ws = nullptr;
    rc = sort_workspace(need, &ws);
    if (rc != 0)
        return rc;
    unsigned char* p = (unsigned char*)ws;
    uint32_t* kbuf[2] = { (uint32_t*)p, (uint32_t*)(p + arr) };
    uint32_t* ibuf[2] = { (uint32_t*)(p + 2 * arr), (uint32_t*)(p + 3 * arr) };
    uint32_t* counts = (uint32_t*)(p + 4 * arr);
    unsigned long long* row_total = (unsigned long long*)(p + 4 * arr + counts_b);
    unsigned long long* digit_base = row_total + RADIX;
    unsigned long long* census = digit_base + RADIX;

    DevStats& stats = dev_stats();
    cudaMemsetAsync(census, 0, 4 * RADIX * 8, st);
    int census_grid = dev_sm_count() * 4;
    uint64_t want = (n / 4 + 511) / 512;
    if ((uint64_t)census_grid > want)
        census_grid = want ? (int)want : 1;
    k4_digit_census<<<census_grid, 512, 0, st>>>(keys, n, census);
    stats.kernel_launches++;
    cudaMemcpyAsync(g_census_host, census, 4 * RADIX * 8, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess)
        {
        set_last_error(std::string("sort_ids census: ") + cudaGetErrorString(cudaGetLastError()));
        return -1;
        }
    int passes[4], npass = 0;
    for (int b = 0; b < 4; b++)
        {
        bool constant = false;
        for (int d = 0; d < RADIX; d++)
            if (g_census_host[b * RADIX + d] == n)
                constant = true;
        if (!constant)
            passes[npass++] = b;
        }

    const uint32_t* kin = keys;
    const uint32_t* iin = nullptr;
    int cur = 0;
    for (int pi = 0; pi < npass; pi++)
        {
        const int shift = passes[pi] * 8;
        const bool last = (pi == npass - 1);
        uint32_t* kout = (last && keys_sorted) ? keys_sorted : kbuf[cur];
        uint32_t* iout = (last && perm) ? perm : ibuf[cur];
        k4_tile_histogram<<<ntiles, SORT_THREADS, 0, st>>>(kin, n, shift, ntiles, counts);
        k4_row_scan<<<RADIX, 256, 0, st>>>(counts, ntiles, row_total);
        k4_digit_base<<<1, RADIX, 0, st>>>(row_total, digit_base);
        if (pi == 0)
            k4_scatter<true><<<ntiles, SORT_THREADS, SCATTER_SMEM, st>>>(kin, nullptr, kout, iout, n, shift,
                                                                         ntiles, counts, digit_base);
        else
            k4_scatter<false><<<ntiles, SORT_THREADS, SCATTER_SMEM, st>>>(kin, iin, kout, iout, n, shift,
                                                                          ntiles, counts, digit_base);
        stats.kernel_launches += 4;
        kin = kout;
        iin = iout;
        cur ^= 1;
        }
    if (npass == 0)
        {
        // all keys equal: the stable order is the identity
        if (perm)
            {
            k4_iota<<<dev_sm_count() * 8, 256, 0, st>>>(perm, n);
            stats.kernel_launches++;
            }
        if (keys_sorted && keys_sorted != keys)
            cudaMemcpyAsync(keys_sorted, keys, n * 4, cudaMemcpyDeviceToDevice, st);
        }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        {
        set_last_error(std::string("sort_ids: ") + cudaGetErrorString(e));
        return -1;
        }
    return 0;
    }

int dev_gather(uint64_t n, const uint32_t* perm, int nfields, const ReorderField* fields, void* stream_v)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    if (nfields < 0 || (nfields > 0 && fields == nullptr))
        return -2;
    if (n == 0 || nfields == 0)
        return 0;
    if (perm == nullptr)
        return -2;
    cudaStream_t st = (cudaStream_t)stream_v;
    for (int f0 = 0; f0 < nfields; f0 += MAX_GATHER_FIELDS)
        {
        GatherArgs args;
        memset(&args, 0, sizeof(args));
        args.nfields = nfields - f0 < MAX_GATHER_FIELDS ? nfields - f0 : MAX_GATHER_FIELDS;
        for (int i = 0; i < args.nfields; i++)
            {
            const ReorderField& f = fields[f0 + i];
            if (f.row_bytes == 0 || f.row_bytes > 1024 || f.in == nullptr || f.out == nullptr || f.in == f.out)
                {
                set_last_error("gather: bad field (row_bytes 1..1024, in/out non-null and distinct)");
                return -2;
                }
            args.f[i].in = (const unsigned char*)f.in;
            args.f[i].out = (unsigned char*)f.out;
            args.f[i].row_bytes = f.row_bytes;
            bool word_ok = (f.row_bytes % 4 == 0) && (((uintptr_t)f.in | (uintptr_t)f.out) % 4 == 0);
            args.f[i].words = word_ok ? f.row_bytes / 4 : 0;
            }
        uint64_t warps = (n + 31) / 32;
        uint64_t blocks = (warps + GATHER_THREADS / 32 - 1) / (GATHER_THREADS / 32);
        uint64_t cap = (uint64_t)dev_sm_count() * 32;
        if (blocks > cap)
            blocks = cap;
        k5_gather<<<(unsigned)blocks, GATHER_THREADS, 0, st>>>(perm, n, args);
        dev_stats().kernel_launches++;
        }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        {
        set_last_error(std::string("gather: ") + cudaGetErrorString(e));
        return -1;
        }
    return 0;
    }
} // namespace pgsdb
