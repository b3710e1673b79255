// kernels_sort.cu -- K4 (stable LSD radix sort of particle ids) and K5 (permutation gather).
//
// Oracle definition of the operation (BASELINE.json north_star):
//     o = numpy.argsort(ids, kind='stable');  out_f = in_f[o]  for every field f
// applied to the arrays pgsd.hoomd's frame decode returns in file (rank) order
// (/root/reference/pgsd/pgsd/hoomd.py:724-902 never sorts; README.md:29).
//
// Pipeline of pgsd_b200_reorder_device (dev_reorder_rows below):
//   census     k4_digit_census: OR and AND of all keys in one streaming read -> which key bytes vary
//              (constant bytes are skipped) and the highest varying bit, which selects the bucket digit.
//              Unique ids then take the slot path of kernels_slot.cu; what follows is the general path.
//   bucket     k4_bucket_aos: stable partition of WHOLE ROWS by the top 8 significant key bits into
//              one interleaved copy -- afterwards the output rows of a bucket and their source rows
//              occupy the same index range, so the final gather is local to a few MB (L2) instead of
//              random over the frame.  (k4_bucket_rows: per-field copy for rows wider than 40 words;
//              with a single varying key byte the bucket pass writes the final order directly.)
//   pairs      LSD passes over (key, position in the bucketed copy), 8-bit digits: k4_tile_histogram
//              (shared-memory atomics) -> k4_row_scan -> k4_digit_base -> k4_scatter.  Keys are ranked
//              with warp ballots against per-warp shared-memory digit counters, the tile is re-ordered
//              in (padded) shared memory and digit runs are written coalesced.  Order preserving =>
//              every pass, and the whole sort, is stable.
//   gather     k5_gather_aos: rows of the bucketed copy -> shared memory (cp.async, after a sequential
//              L2 prefetch of the bucket slice) -> fields written back SoA, fully coalesced.
// pgsd_b200_sort_ids = census + pairs on the caller's keys; pgsd_b200_gather = k5_gather, the plain
// out[i] = in[perm[i]] for an arbitrary caller permutation.
//
// sm_100a only.  No CPU fallback: callers fail when CUDA is unavailable.
#include "device_internal.h"

#include <cstdlib>
#include <vector>

namespace pgsdb
{
namespace
    {
constexpr int RADIX = 256;
constexpr int SORT_THREADS = 512;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS; // 8192 keys per CTA (pair passes)
constexpr int PADDED_TILE = SORT_TILE + SORT_TILE / 32;
constexpr int ROWS_ITEMS = 8;
constexpr int ROWS_TILE = SORT_THREADS * ROWS_ITEMS; // 4096 rows per CTA (bucket pass with payload)

// how a warp finds, for each key, the lanes holding the same digit
enum RankMode : int
    {
    RANK_BALLOT = 0, // 8 x __ballot_sync, one per digit bit
    RANK_MATCH = 1   // __match_any_sync
    };

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

__device__ __forceinline__ uint4 ldg_stream_v4(const uint4* p)
    {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
    }

// ---- pre-pass: which key bits vary (OR and AND of all keys; bits 1 in OR and 0 in AND differ somewhere) ----
// out[0] |= OR, out[1] &= AND.  Pure streaming read; replaces four 256-bin byte histograms (the plan only
// ever asked which bytes vary and where the highest varying bit is).
__global__ void __launch_bounds__(512) k4_digit_census(const uint32_t* __restrict__ keys, uint64_t n,
                                                      uint32_t* __restrict__ out)
    {
    uint32_t o = 0u, a = 0xffffffffu;
    const uint64_t n4 = n / 4;
    const uint4* k4 = reinterpret_cast<const uint4*>(keys);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (uint64_t)gridDim.x * blockDim.x)
        {
        const uint4 v = ldg_stream_v4(k4 + i);
        o |= v.x | v.y | v.z | v.w;
        a &= v.x & v.y & v.z & v.w;
        }
    if (blockIdx.x == 0)
        for (uint64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x)
            {
            const uint32_t k = keys[i];
            o |= k;
            a &= k;
            }
    o = __reduce_or_sync(0xffffffffu, o);
    a = __reduce_and_sync(0xffffffffu, a);
    if ((threadIdx.x & 31) == 0)
        {
        if (o != 0u)
            atomicOr(out, o);
        if (a != 0xffffffffu)
            atomicAnd(out + 1, a);
        }
    }

// ---- upsweep: digit histogram of every tile; counts[d * ntiles + tile] ------------------------
template <int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS) k4_tile_histogram(const uint32_t* __restrict__ keys,
                                                                  uint64_t n, int shift, uint32_t ntiles,
                                                                  uint32_t* __restrict__ counts)
    {
    constexpr int SORT_ITEMS = ITEMS;             // shadows the pair-pass constants on purpose:
    constexpr int SORT_TILE = SORT_THREADS * ITEMS; // the tile must match the scatter kernel's
    __shared__ unsigned int h[SORT_WARPS][RADIX]; // one sub-histogram per warp: atomics contend only inside a warp
    for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS)
        (&h[0][0])[i] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * SORT_TILE;
    unsigned int* mine = h[threadIdx.x >> 5];
    if (ITEMS >= 4 && base + SORT_TILE <= n)
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
        const uint64_t end = base + SORT_TILE < n ? base + SORT_TILE : n;
        for (uint64_t i = base + threadIdx.x; i < end; i += SORT_THREADS)
            atomicAdd(&mine[(keys[i] >> shift) & 255u], 1u);
        }
    __syncthreads();
    for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS)
        {
        unsigned int c = 0;
#pragma unroll
        for (int s = 0; s < SORT_WARPS; s++)
            c += h[s][d];
        counts[(size_t)d * ntiles + blockIdx.x] = c;
        }
    }

// ---- scan: one CTA per digit row, exclusive scan over tiles in place; row total out -----------
__global__ void __launch_bounds__(256) k4_row_scan(uint32_t* __restrict__ counts, uint32_t ntiles,
                                                   unsigned long long* __restrict__ row_total)
    {
    constexpr int PER = 8; // tiles per thread: 2048 tiles per block-wide step
    __shared__ uint32_t warp_sum[8];
    __shared__ uint32_t carry_s;
    uint32_t* row = counts + (size_t)blockIdx.x * ntiles;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0)
        carry_s = 0;
    __syncthreads();
    for (uint32_t start = 0; start < ntiles; start += 256 * PER)
        {
        const uint32_t i0 = start + threadIdx.x * PER;
        uint32_t v[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; k++)
            {
            v[k] = i0 + k < ntiles ? row[i0 + k] : 0u;
            sum += v[k];
            }
        uint32_t x = sum;
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
        const uint32_t carry = carry_s;
        uint32_t acc = carry + wbase + x - sum;
#pragma unroll
        for (int k = 0; k < PER; k++)
            {
            if (i0 + k < ntiles)
                row[i0 + k] = acc;
            acc += v[k];
            }
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

// ---- downsweep: stable partition of one tile by one digit --------------------------------------
// Shared by the pair passes (payload = original index) and the bucket pass (payload = whole rows).
//
// Element e = w*(32*ITEMS) + k*32 + lane of the tile is held by lane `lane` of warp `w` in register
// slot k: tile order == global order.  Ranking is order preserving, so every pass is stable:
//   1. per warp, per item: peers = lanes with the same digit (ballots or match.any);
//      rank = running per-warp digit count + number of peers in lower lanes;
//   2. per digit: exclusive scan over the warps, then over the digits -> tile-local start;
//   3. pos = start[d] + warp offset[d] + rank : position in the tile sorted by digit.
struct TileRankSmem
    {
    uint32_t* whist;            // SORT_WARPS * RADIX
    uint32_t* dstart;           // RADIX
    unsigned long long* gdelta; // RADIX : global position - tile-local position
    uint32_t* wtot;             // 8
    };

template <int ITEMS, int RM>
__device__ __forceinline__ void tile_rank(const uint32_t (&key)[ITEMS], uint32_t (&pos)[ITEMS], uint32_t tile_n,
                                          int shift, const TileRankSmem& sm, unsigned long long my_gbase)
    {
    // my_gbase (threads < RADIX): global output position of this tile's first key with digit `tid`
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t wbase = (uint32_t)w * (32 * ITEMS);
    for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS)
        sm.whist[i] = 0;
    __syncthreads();
    uint32_t* myhist = sm.whist + w * RADIX;
    const unsigned lt = lanemask_lt();
    // (a) peers of every item: register-only work, all ballots of the tile pipeline freely
    unsigned peers[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++)
        {
        const uint32_t e = wbase + k * 32 + lane;
        const bool valid = e < tile_n;
        const uint32_t d = (key[k] >> shift) & 255u;
        if (RM == RANK_MATCH)
            {
            // invalid lanes (tile tail) match among themselves on an out-of-range value
            peers[k] = __match_any_sync(0xffffffffu, valid ? d : 0x100u);
            }
        else
            {
            unsigned m = __ballot_sync(0xffffffffu, valid);
#pragma unroll
            for (int bit = 0; bit < 8; bit++)
                {
                const bool one = (d >> bit) & 1u;
                const unsigned bal = __ballot_sync(0xffffffffu, one);
                m &= one ? bal : ~bal;
                }
            peers[k] = m;
            }
        }
    // (b) running per-warp digit counts: every peer reads the count (one broadcast read), then the
    //     group's first lane adds the group size -- a short read -> write chain per item
#pragma unroll
    for (int k = 0; k < ITEMS; k++)
        {
        const uint32_t e = wbase + k * 32 + lane;
        const bool valid = e < tile_n;
        const uint32_t d = (key[k] >> shift) & 255u;
        uint32_t old = 0;
        if (valid)
            old = myhist[d];
        __syncwarp();
        if (valid && (peers[k] & lt) == 0)
            myhist[d] = old + __popc(peers[k]);
        __syncwarp();
        pos[k] = old + __popc(peers[k] & lt);
        }
    __syncthreads();

    uint32_t total = 0;
    if (tid < RADIX)
        {
#pragma unroll
        for (int j = 0; j < SORT_WARPS; j++)
            {
            uint32_t c = sm.whist[j * RADIX + tid];
            sm.whist[j * RADIX + tid] = total;
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
            sm.wtot[w] = x;
        total = x - total; // exclusive within the warp, completed below
        }
    __syncthreads();
    if (tid < RADIX)
        {
        uint32_t b = 0;
        for (int j = 0; j < w; j++)
            b += sm.wtot[j];
        const uint32_t start = total + b;
        sm.dstart[tid] = start;
        sm.gdelta[tid] = my_gbase - (unsigned long long)start;
        }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ITEMS; k++)
        {
        const uint32_t d = (key[k] >> shift) & 255u;
        pos[k] += sm.dstart[d] + myhist[d];
        }
    }

// Tile table of the segmented pair passes: after the bucket pass every bucket is sorted on its own
// (low key bits only), so tiles never straddle buckets and no final pass over the bucket digit is needed.
struct SegTables
    {
    const uint32_t* tile_begin; // first key of tile t
    const uint32_t* tile_cnt;   // keys in tile t (<= SORT_TILE)
    const uint32_t* tile_bkt;   // bucket of tile t
    const uint32_t* ntiles;     // [0] = number of tiles actually used
    const uint32_t* bstart;     // bucket start (257 entries)
    const uint32_t* bbase;      // [bucket * 256 + digit]: keys of the bucket with a smaller digit
    uint32_t stride;            // row stride of counts[] (upper bound of the tile count)
    };

// pair pass.  FIRST: the index payload is implicit (idx = global position), saving its read.
// SEG: tiles and output positions come from the per-bucket tables above.
template <bool FIRST, int RM, bool SEG>
__global__ void __launch_bounds__(SORT_THREADS, 2)
    k4_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ idx_in,
               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ idx_out, uint64_t n, int shift,
               uint32_t ntiles, const uint32_t* __restrict__ tile_offset,
               const unsigned long long* __restrict__ digit_base, const SegTables seg)
    {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // one pad word per 32: with exactly 32 keys per digit in a tile (dense ids in the later passes)
    // the 32 lanes of a warp would otherwise all store to positions 32*d + c, i.e. one bank
    uint32_t* skeys = reinterpret_cast<uint32_t*>(smem_raw); // PADDED_TILE
    uint32_t* sidx = skeys + PADDED_TILE;                     // PADDED_TILE
    TileRankSmem sm;
    sm.whist = sidx + PADDED_TILE;
    sm.dstart = sm.whist + SORT_WARPS * RADIX;
    sm.gdelta = reinterpret_cast<unsigned long long*>(sm.dstart + RADIX);
    __shared__ uint32_t wtot[8];
    sm.wtot = wtot;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    uint64_t tile_base;
    uint32_t tile_n;
    unsigned long long my_gbase = 0;
    if (SEG)
        {
        if (blockIdx.x >= seg.ntiles[0])
            return;
        tile_base = seg.tile_begin[blockIdx.x];
        tile_n = seg.tile_cnt[blockIdx.x];
        if (tid < RADIX)
            {
            const uint32_t b = seg.tile_bkt[blockIdx.x];
            my_gbase = (unsigned long long)seg.bstart[b] + seg.bbase[b * RADIX + tid]
                       + tile_offset[(size_t)tid * seg.stride + blockIdx.x];
            }
        }
    else
        {
        tile_base = (uint64_t)blockIdx.x * SORT_TILE;
        tile_n = (uint32_t)((n - tile_base) < (uint64_t)SORT_TILE ? (n - tile_base) : SORT_TILE);
        if (tid < RADIX)
            my_gbase = digit_base[tid] + (unsigned long long)tile_offset[(size_t)tid * ntiles + blockIdx.x];
        }
    const uint32_t wbase = (uint32_t)w * (32 * SORT_ITEMS);

    uint32_t key[SORT_ITEMS], pos[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++)
        {
        const uint32_t e = wbase + k * 32 + lane;
        key[k] = e < tile_n ? ld_stream_u32(keys_in + tile_base + e) : 0xffffffffu;
        }
    tile_rank<SORT_ITEMS, RM>(key, pos, tile_n, shift, sm, my_gbase);
    // the index payload is fetched only now (keeps the ranking loop's register footprint small);
    // all loads of the batch are issued before the first shared-memory store consumes one
    uint32_t src[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++)
        {
        const uint32_t e = wbase + k * 32 + lane;
        if (FIRST)
            src[k] = (uint32_t)(tile_base + e);
        else
            src[k] = e < tile_n ? ld_stream_u32(idx_in + tile_base + e) : 0u;
        }
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++)
        {
        const uint32_t e = wbase + k * 32 + lane;
        if (e < tile_n)
            {
            const uint32_t pp = pos[k] + (pos[k] >> 5);
            skeys[pp] = key[k];
            sidx[pp] = src[k];
            }
        }
    __syncthreads();
    // write digit runs out: consecutive j of one digit -> consecutive global addresses
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++)
        {
        const uint32_t j = tid + k * SORT_THREADS;
        if (j < tile_n)
            {
            const uint32_t jp = j + (j >> 5);
            const uint32_t kv = skeys[jp];
            const unsigned long long g = sm.gdelta[(kv >> shift) & 255u] + j;
            keys_out[g] = kv;
            idx_out[g] = sidx[jp];
            }
        }
    }

// ---- tables of the segmented passes --------------------------------------------------------------
// one CTA of 256 threads, thread b = bucket b: tile counts, their exclusive scan, tile descriptors
__global__ void __launch_bounds__(RADIX)
    k4s_setup(const unsigned long long* __restrict__ bucket_base, uint64_t n, uint32_t* __restrict__ bstart,
              uint32_t* __restrict__ tile_first, uint32_t* __restrict__ tile_begin, uint32_t* __restrict__ tile_cnt,
              uint32_t* __restrict__ tile_bkt, uint32_t* __restrict__ ntiles_out)
    {
    __shared__ uint32_t s[RADIX];
    const int b = threadIdx.x;
    const uint32_t begin = (uint32_t)bucket_base[b];
    const uint32_t end = b + 1 < RADIX ? (uint32_t)bucket_base[b + 1] : (uint32_t)n;
    const uint32_t size = end - begin;
    const uint32_t nt = (size + SORT_TILE - 1) / SORT_TILE;
    s[b] = nt;
    __syncthreads();
    if (b == 0)
        {
        uint32_t acc = 0;
        for (int i = 0; i < RADIX; i++)
            {
            const uint32_t c = s[i];
            s[i] = acc;
            acc += c;
            }
        ntiles_out[0] = acc;
        tile_first[RADIX] = acc;
        bstart[RADIX] = (uint32_t)n;
        }
    __syncthreads();
    const uint32_t first = s[b];
    bstart[b] = begin;
    tile_first[b] = first;
    for (uint32_t i = 0; i < nt; i++)
        {
        tile_begin[first + i] = begin + i * SORT_TILE;
        tile_cnt[first + i] = size - i * SORT_TILE < (uint32_t)SORT_TILE ? size - i * SORT_TILE : SORT_TILE;
        tile_bkt[first + i] = (uint32_t)b;
        }
    }

__global__ void __launch_bounds__(SORT_THREADS)
    k4s_histogram(const uint32_t* __restrict__ keys, int shift, const SegTables seg, uint32_t* __restrict__ counts)
    {
    if (blockIdx.x >= seg.ntiles[0])
        return;
    __shared__ unsigned int h[SORT_WARPS][RADIX];
    for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS)
        (&h[0][0])[i] = 0;
    __syncthreads();
    const uint32_t base = seg.tile_begin[blockIdx.x], cnt = seg.tile_cnt[blockIdx.x];
    unsigned int* mine = h[threadIdx.x >> 5];
    uint32_t v[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++)
        {
        const uint32_t e = threadIdx.x + k * SORT_THREADS;
        v[k] = e < cnt ? ld_stream_u32(keys + base + e) : 0u;
        }
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++)
        if (threadIdx.x + k * SORT_THREADS < cnt)
            atomicAdd(&mine[(v[k] >> shift) & 255u], 1u);
    __syncthreads();
    for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS)
        {
        unsigned int c = 0;
#pragma unroll
        for (int q = 0; q < SORT_WARPS; q++)
            c += h[q][d];
        counts[(size_t)d * seg.stride + blockIdx.x] = c;
        }
    }

// CTA d = digit d: exclusive scan of the digit's tile counts inside every bucket (in place) and the
// bucket totals.  Buckets of <= 32 tiles (the normal case: 8 tiles per bucket for 16 Mi dense ids) are
// walked by one thread each, all 256 at once; larger ones by a whole warp, 32 tiles per step.
__global__ void __launch_bounds__(256)
    k4s_scan(uint32_t* __restrict__ counts, const uint32_t* __restrict__ tile_first, uint32_t stride,
             uint32_t* __restrict__ btot)
    {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t* row = counts + (size_t)blockIdx.x * stride;
        {
        const int b = threadIdx.x;
        const uint32_t t0 = tile_first[b], t1 = tile_first[b + 1];
        if (t1 - t0 <= 32)
            {
            uint32_t acc = 0;
            for (uint32_t i = t0; i < t1; i++)
                {
                const uint32_t v = row[i];
                row[i] = acc;
                acc += v;
                }
            btot[b * RADIX + blockIdx.x] = acc;
            }
        }
    for (int b = w; b < RADIX; b += 8)
        {
        const uint32_t t0 = tile_first[b], t1 = tile_first[b + 1];
        if (t1 - t0 <= 32)
            continue;
        uint32_t carry = 0;
        for (uint32_t s = t0; s < t1; s += 32)
            {
            const uint32_t i = s + lane;
            const uint32_t v = i < t1 ? row[i] : 0u;
            uint32_t x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
                {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o)
                    x += y;
                }
            if (i < t1)
                row[i] = carry + x - v;
            carry += __shfl_sync(0xffffffffu, x, 31);
            }
        if (lane == 0)
            btot[b * RADIX + blockIdx.x] = carry;
        }
    }

// CTA b = bucket b: exclusive scan over the digits of the bucket's totals
__global__ void __launch_bounds__(RADIX) k4s_base(const uint32_t* __restrict__ btot, uint32_t* __restrict__ bbase)
    {
    __shared__ uint32_t ws[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t v = btot[blockIdx.x * RADIX + threadIdx.x];
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
        {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o)
            x += y;
        }
    if (lane == 31)
        ws[w] = x;
    __syncthreads();
    uint32_t add = 0;
    for (int j = 0; j < w; j++)
        add += ws[j];
    bbase[blockIdx.x * RADIX + threadIdx.x] = add + x - v;
    }

constexpr size_t SCATTER_SMEM
    = (size_t)(2 * PADDED_TILE + SORT_WARPS * RADIX + RADIX) * sizeof(uint32_t) + RADIX * sizeof(unsigned long long);

// ---- bucket pass: the same stable partition, carrying whole rows ---------------------------------
// Groups the rows of every field by the top 8 significant key bits so that the permutation gather
// that follows the pair passes reads from a window of a few buckets (L2 resident) instead of from
// random rows of the whole frame (measured without it: 9.2 GB of DRAM reads for 0.67 GB of payload
// at 16 Mi particles, profiles/r1a).  When only one key byte varies this pass IS the sort.
constexpr int MAX_ROW_FIELDS = 16;
constexpr int ROWS_STAGE_WORDS = 3 * ROWS_TILE; // staging buffer: rows of <= 3 words in one go
struct RowField
    {
    const uint32_t* in; // n rows of `words` 32-bit words; NULL: the row's original index (iota)
    uint32_t* out;
    uint32_t words;
    };
struct RowArgs
    {
    RowField f[MAX_ROW_FIELDS];
    int nfields;
    };

template <int RM>
__global__ void __launch_bounds__(SORT_THREADS)
    k4_bucket_rows(const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ keys_out, uint64_t n, int shift,
                   uint32_t ntiles, const uint32_t* __restrict__ tile_offset,
                   const unsigned long long* __restrict__ digit_base, const __grid_constant__ RowArgs args)
    {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* skeys = reinterpret_cast<uint32_t*>(smem_raw); // ROWS_TILE : keys in tile-sorted order
    uint32_t* stage = skeys + ROWS_TILE;                      // ROWS_STAGE_WORDS
    uint16_t* spos = reinterpret_cast<uint16_t*>(stage + ROWS_STAGE_WORDS); // ROWS_TILE : row -> sorted position
    TileRankSmem sm;
    sm.whist = reinterpret_cast<uint32_t*>(spos + ROWS_TILE);
    sm.dstart = sm.whist + SORT_WARPS * RADIX;
    sm.gdelta = reinterpret_cast<unsigned long long*>(sm.dstart + RADIX);
    __shared__ uint32_t wtot[8];
    sm.wtot = wtot;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint64_t tile_base = (uint64_t)blockIdx.x * ROWS_TILE;
    const uint32_t tile_n = (uint32_t)((n - tile_base) < (uint64_t)ROWS_TILE ? (n - tile_base) : ROWS_TILE);
    const uint32_t wbase = (uint32_t)w * (32 * ROWS_ITEMS);

    uint32_t key[ROWS_ITEMS], pos[ROWS_ITEMS];
#pragma unroll
    for (int k = 0; k < ROWS_ITEMS; k++)
        {
        const uint32_t e = wbase + k * 32 + lane;
        key[k] = e < tile_n ? ld_stream_u32(keys_in + tile_base + e) : 0xffffffffu;
        }
    const unsigned long long my_gbase
        = tid < RADIX ? digit_base[tid] + (unsigned long long)tile_offset[(size_t)tid * ntiles + blockIdx.x] : 0ull;
    tile_rank<ROWS_ITEMS, RM>(key, pos, tile_n, shift, sm, my_gbase);
#pragma unroll
    for (int k = 0; k < ROWS_ITEMS; k++)
        {
        const uint32_t e = wbase + k * 32 + lane;
        if (e < tile_n)
            {
            skeys[pos[k]] = key[k];
            spos[e] = (uint16_t)pos[k];
            }
        }
    __syncthreads();
    // keys: digit runs out, coalesced
    if (keys_out)
        {
#pragma unroll
        for (int k = 0; k < ROWS_ITEMS; k++)
            {
            const uint32_t j = tid + k * SORT_THREADS;
            if (j < tile_n)
                {
                const uint32_t kv = skeys[j];
                keys_out[sm.gdelta[(kv >> shift) & 255u] + j] = kv;
                }
            }
        }
    // fields: tile rows -> staging buffer in sorted order -> digit runs out; both sides coalesced
    for (int fi = 0; fi < args.nfields; fi++)
        {
        const RowField f = args.f[fi];
        const uint32_t W = f.words;
        const uint32_t rows_per_round = ROWS_STAGE_WORDS / W; // >= ROWS_TILE when W <= 3
        for (uint32_t r0 = 0; r0 < tile_n; r0 += rows_per_round)
            {
            // sorted positions [r0, r1) are staged in this round
            const uint32_t r1 = r0 + rows_per_round < tile_n ? r0 + rows_per_round : tile_n;
            __syncthreads(); // staging buffer free
            if (f.in == nullptr)
                {
                for (uint32_t e = tid; e < tile_n; e += SORT_THREADS)
                    {
                    const uint32_t p = spos[e];
                    if (p >= r0 && p < r1)
                        stage[p - r0] = (uint32_t)(tile_base + e);
                    }
                }
            else if (W == 1)
                {
                // 8 independent loads per thread in flight, then the shared-memory stores
                const uint32_t* in = f.in + tile_base;
                uint32_t v[ROWS_ITEMS];
#pragma unroll
                for (int k = 0; k < ROWS_ITEMS; k++)
                    {
                    const uint32_t e = tid + k * SORT_THREADS;
                    v[k] = e < tile_n ? ld_stream_u32(in + e) : 0u;
                    }
#pragma unroll
                for (int k = 0; k < ROWS_ITEMS; k++)
                    {
                    const uint32_t e = tid + k * SORT_THREADS;
                    if (e < tile_n)
                        {
                        const uint32_t p = spos[e];
                        if (p >= r0 && p < r1)
                            stage[p - r0] = v[k];
                        }
                    }
                }
            else if (W <= 3 && tile_n == ROWS_TILE && (reinterpret_cast<uintptr_t>(f.in) & 15u) == 0)
                {
                // full tile of 2- or 3-word rows: the tile is a flat run of 1024*W 16-byte vectors;
                // every thread issues its 2*W vector loads first
                const uint4* in4 = reinterpret_cast<const uint4*>(f.in + tile_base * W);
                uint4 v[2 * 3];
#pragma unroll
                for (int k = 0; k < 2 * 3; k++)
                    if (k < 2 * (int)W)
                        v[k] = ldg_stream_v4(in4 + tid + k * SORT_THREADS);
#pragma unroll
                for (int k = 0; k < 2 * 3; k++)
                    if (k < 2 * (int)W)
                        {
                        const uint32_t q = (tid + k * SORT_THREADS) * 4;
                        const uint32_t words[4] = { v[k].x, v[k].y, v[k].z, v[k].w };
                        uint32_t e = q / W, c = q - e * W;
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            {
                            const uint32_t p = spos[e];
                            if (p >= r0 && p < r1)
                                stage[(p - r0) * W + c] = words[u];
                            if (++c == W)
                                {
                                c = 0;
                                e++;
                                }
                            }
                        }
                }
            else
                {
                const uint32_t* in = f.in + tile_base * W;
                const uint32_t total = tile_n * W;
                for (uint32_t q0 = 0; q0 < total; q0 += 8 * SORT_THREADS)
                    {
                    uint32_t v[8];
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        {
                        const uint32_t q = q0 + tid + k * SORT_THREADS;
                        v[k] = q < total ? ld_stream_u32(in + q) : 0u;
                        }
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        {
                        const uint32_t q = q0 + tid + k * SORT_THREADS;
                        if (q < total)
                            {
                            const uint32_t e = q / W, c = q - e * W;
                            const uint32_t p = spos[e];
                            if (p >= r0 && p < r1)
                                stage[(p - r0) * W + c] = v[k];
                            }
                        }
                    }
                }
            __syncthreads();
            const uint32_t total_out = (r1 - r0) * W;
            if (W == 1)
                {
                for (uint32_t j = r0 + tid; j < r1; j += SORT_THREADS)
                    f.out[sm.gdelta[(skeys[j] >> shift) & 255u] + j] = stage[j - r0];
                }
            else
                {
                uint32_t j = tid / W, c = tid - j * W;
                const uint32_t dj = SORT_THREADS / W, dc = SORT_THREADS - dj * W;
                for (uint32_t q = tid; q < total_out; q += SORT_THREADS)
                    {
                    const unsigned long long g = sm.gdelta[(skeys[r0 + j] >> shift) & 255u] + r0 + j;
                    f.out[g * W + c] = stage[q];
                    j += dj;
                    c += dc;
                    if (c >= W)
                        {
                        c -= W;
                        j++;
                        }
                    }
                }
            }
        }
    }

constexpr size_t ROWS_SMEM = (size_t)(ROWS_TILE + ROWS_STAGE_WORDS + SORT_WARPS * RADIX + RADIX) * sizeof(uint32_t)
                             + ROWS_TILE * sizeof(uint16_t) + RADIX * sizeof(unsigned long long);

// ---- bucket pass, AoS flavour (the fast path) ----------------------------------------------------
// Same stable partition, but the bucketed copy is ONE array of interleaved rows (all fields of a
// particle side by side, RW words).  The whole tile is staged in shared memory in sorted order and
// leaves as one contiguous run per bucket (hundreds of bytes instead of one short run per field);
// the gather that follows reads 2 sectors per particle instead of one per field.
constexpr int AOS_MAX_ROW_WORDS = 40;
struct AosField
    {
    const uint32_t* in; // n rows of `words` words; NULL: the row's original index
    uint32_t* out;      // final destination (used by k5_gather_aos only)
    uint32_t words;
    uint32_t off;       // word offset inside the interleaved row
    };
struct AosArgs
    {
    AosField f[MAX_ROW_FIELDS];
    int nfields;
    uint32_t row_words; // RW
    };

template <int ITEMS, int RM>
__global__ void __launch_bounds__(SORT_THREADS, 2)
    k4_bucket_aos(const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ aos_out,
                  uint64_t n, int shift, uint32_t ntiles, const uint32_t* __restrict__ tile_offset,
                  const unsigned long long* __restrict__ digit_base, const __grid_constant__ AosArgs args)
    {
    constexpr int TILE = SORT_THREADS * ITEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* skeys = reinterpret_cast<uint32_t*>(smem_raw);          // TILE : keys in tile-sorted order
    uint16_t* sinv = reinterpret_cast<uint16_t*>(skeys + TILE);       // TILE : sorted position -> tile row
    TileRankSmem sm;
    sm.whist = reinterpret_cast<uint32_t*>(sinv + TILE);
    sm.dstart = sm.whist + SORT_WARPS * RADIX;
    sm.gdelta = reinterpret_cast<unsigned long long*>(sm.dstart + RADIX);
    uint32_t* raw = reinterpret_cast<uint32_t*>(sm.gdelta + RADIX);   // TILE * (words of the real fields), field-major
    __shared__ uint32_t wtot[8];
    __shared__ uint32_t col[AOS_MAX_ROW_WORDS]; // column c of the interleaved row: (raw base << 8) | W; W = 0: original index
    sm.wtot = wtot;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t RW = args.row_words;
    const uint64_t tile_base = (uint64_t)blockIdx.x * TILE;
    const uint32_t tile_n = (uint32_t)((n - tile_base) < (uint64_t)TILE ? (n - tile_base) : TILE);
    const uint32_t wbase = (uint32_t)w * (32 * ITEMS);

    // (1) the tile of every field starts its way global -> shared now (cp.async, no registers held);
    //     ranking the keys below overlaps with the DRAM latency of the payload
    uint32_t fbase = 0; // word offset of the field's tile copy in raw[]
    for (int fi = 0; fi < args.nfields; fi++)
        {
        const AosField f = args.f[fi];
        const uint32_t W = f.words;
        if (tid < (int)W)
            col[f.off + tid] = f.in ? (((fbase + (uint32_t)tid) << 8) | W) : 0u;
        if (f.in == nullptr)
            continue;
        const uint32_t* in = f.in + tile_base * W;
        const uint32_t total = tile_n * W;
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(raw + fbase);
        if (tile_n == TILE && (reinterpret_cast<uintptr_t>(in) & 15u) == 0)
            {
            const uint32_t nvec = total / 4;
            for (uint32_t vq = tid; vq < nvec; vq += SORT_THREADS)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + vq * 16), "l"(in + vq * 4) : "memory");
            }
        else
            {
            for (uint32_t q = tid; q < total; q += SORT_THREADS)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sbase + q * 4), "l"(in + q) : "memory");
            }
        fbase += (uint32_t)TILE * W;
        }
    asm volatile("cp.async.commit_group;" ::: "memory");

    // (2) rank the keys
    uint32_t key[ITEMS], pos[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++)
        {
        const uint32_t e = wbase + k * 32 + lane;
        key[k] = e < tile_n ? ld_stream_u32(keys_in + tile_base + e) : 0xffffffffu;
        }
    const unsigned long long my_gbase
        = tid < RADIX ? digit_base[tid] + (unsigned long long)tile_offset[(size_t)tid * ntiles + blockIdx.x] : 0ull;
    tile_rank<ITEMS, RM>(key, pos, tile_n, shift, sm, my_gbase);
#pragma unroll
    for (int k = 0; k < ITEMS; k++)
        {
        const uint32_t e = wbase + k * 32 + lane;
        if (e < tile_n)
            {
            skeys[pos[k]] = key[k];
            sinv[pos[k]] = (uint16_t)e;
            }
        }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // (3) out: keys as digit runs; rows as one contiguous run of RW-word records per digit, read from
    //     the raw field tiles through the sorted-position -> row map
    if (keys_out)
        {
#pragma unroll
        for (int k = 0; k < ITEMS; k++)
            {
            const uint32_t j = tid + k * SORT_THREADS;
            if (j < tile_n)
                {
                const uint32_t kv = skeys[j];
                keys_out[sm.gdelta[(kv >> shift) & 255u] + j] = kv;
                }
            }
        }
    if ((RW & 1u) == 0)
        {
        // rows are 8-byte multiples: one 64-bit store per two columns.  Thread -> (row slot, column
        // pair) is fixed for the whole loop, so the column descriptors live in registers and
        // consecutive threads still write consecutive 8-byte pieces of consecutive rows.
        const uint32_t R2 = RW / 2;
        const uint32_t rows_per_step = SORT_THREADS / R2;
        const uint32_t slot = tid / R2, c = tid - slot * R2;
        if (slot < rows_per_step)
            {
            uint2* out2 = reinterpret_cast<uint2*>(aos_out);
            const uint32_t c0 = col[2 * c], c1 = col[2 * c + 1];
            const uint32_t w0 = c0 & 255u, w1 = c1 & 255u, b0 = c0 >> 8, b1 = c1 >> 8;
            for (uint32_t j = slot; j < tile_n; j += rows_per_step)
                {
                const unsigned long long g = sm.gdelta[(skeys[j] >> shift) & 255u] + j;
                const uint32_t e = sinv[j];
                uint2 v;
                v.x = w0 ? raw[b0 + e * w0] : (uint32_t)(tile_base + e);
                v.y = w1 ? raw[b1 + e * w1] : (uint32_t)(tile_base + e);
                out2[g * R2 + c] = v;
                }
            }
        }
    else
        {
        const uint32_t total_out = tile_n * RW;
        uint32_t j = tid / RW, c = tid - j * RW;
        const uint32_t dj = SORT_THREADS / RW, dc = SORT_THREADS - dj * RW;
        for (uint32_t q = tid; q < total_out; q += SORT_THREADS)
            {
            const unsigned long long g = sm.gdelta[(skeys[j] >> shift) & 255u] + j;
            const uint32_t e = sinv[j];
            const uint32_t cc = col[c];
            aos_out[g * RW + c] = (cc & 255u) ? raw[(cc >> 8) + e * (cc & 255u)] : (uint32_t)(tile_base + e);
            j += dj;
            c += dc;
            if (c >= RW)
                {
                c -= RW;
                j++;
                }
            }
        }
    }

template <int ITEMS> constexpr size_t aos_smem_fixed()
    {
    return (size_t)SORT_THREADS * ITEMS * (sizeof(uint32_t) + sizeof(uint16_t))
           + (size_t)(SORT_WARPS * RADIX + RADIX) * sizeof(uint32_t) + RADIX * sizeof(unsigned long long);
    }

// ---- K5, AoS flavour: out_f[i] = row perm[i] of the bucketed copy, split back into the fields -----
// A warp owns 4 chunks of 32 consecutive output rows.  The source rows go global -> shared with
// cp.async (no registers held while ~40 random 4/8-byte reads per lane are in flight), then every
// field is written out fully coalesced from the staged rows.
constexpr int GA_THREADS = 128;
constexpr int GA_CHUNKS = 4;
constexpr int GA_ROWS_PER_CTA = (GA_THREADS / 32) * 32 * GA_CHUNKS;

template <int VEC, bool PREFETCH> // VEC: words per cp.async, 2 when RW is even (8-byte aligned copies), else 1
__global__ void __launch_bounds__(GA_THREADS)
    k5_gather_aos(const uint32_t* __restrict__ perm, uint64_t n, const uint32_t* __restrict__ aos,
                  const __grid_constant__ AosArgs args)
    {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t RW = args.row_words;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t* stage = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)w * (GA_CHUNKS * 32) * RW;
    const uint64_t warp_row0 = (uint64_t)blockIdx.x * GA_ROWS_PER_CTA + (uint64_t)w * (32 * GA_CHUNKS);
    if (warp_row0 >= n)
        return;
    uint32_t p[GA_CHUNKS], rows[GA_CHUNKS];
#pragma unroll
    for (int g = 0; g < GA_CHUNKS; g++)
        {
        const uint64_t row0 = warp_row0 + 32 * g;
        rows[g] = row0 >= n ? 0u : (uint32_t)((n - row0) < 32 ? (n - row0) : 32);
        p[g] = lane < rows[g] ? ld_stream_u32(perm + row0 + lane) : 0u;
        }
    // Output rows [r0, r0+128) of this warp and source rows [r0, r0+128) of the bucketed copy lie in
    // the same bucket(s): the warps that are resident together pull their bucket into L2 with
    // sequential 128-byte prefetches, so that DRAM serves streams while the random reads below
    // mostly hit L2 (random 32-byte sector reads alone hold DRAM at about 1/3 of its bandwidth).
    if (PREFETCH)
        {
        const uint64_t b0 = warp_row0 * RW * 4;
        const uint64_t last = warp_row0 + 32 * GA_CHUNKS < n ? warp_row0 + 32 * GA_CHUNKS : n;
        const uint64_t b1 = last * RW * 4;
        const char* base = reinterpret_cast<const char*>(aos);
        for (uint64_t b = (b0 & ~127ull) + (uint64_t)lane * 128; b < b1; b += 32 * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + b));
        }
    const uint32_t RV = RW / VEC; // copies per row
#pragma unroll
    for (int g = 0; g < GA_CHUNKS; g++)
        {
        const uint32_t total = rows[g] * RV;
        uint32_t r = lane / RV, c = lane - r * RV;
        const uint32_t dr = 32 / RV, dc = 32 - dr * RV;
        for (uint32_t q = lane; q < 32 * RV; q += 32)
            {
            const uint32_t src = __shfl_sync(0xffffffffu, p[g], r & 31);
            if (q < total)
                {
                const uint32_t* gp = aos + (uint64_t)src * RW + c * VEC;
                const uint32_t sp = (uint32_t)__cvta_generic_to_shared(stage + (size_t)g * 32 * RW + q * VEC);
                if (VEC == 2)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sp), "l"(gp) : "memory");
                else
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sp), "l"(gp) : "memory");
                }
            r += dr;
            c += dc;
            if (c >= RV)
                {
                c -= RV;
                r++;
                }
            }
        }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    for (int fi = 0; fi < args.nfields; fi++)
        {
        const AosField f = args.f[fi];
        const uint32_t W = f.words;
#pragma unroll
        for (int g = 0; g < GA_CHUNKS; g++)
            {
            const uint32_t* srow = stage + (size_t)g * 32 * RW + f.off;
            uint32_t* out = f.out + (warp_row0 + 32 * g) * W;
            const uint32_t total = rows[g] * W;
            if (W == 1)
                {
                if (lane < total)
                    out[lane] = srow[lane * RW];
                }
            else if (W == 3)
                {
#pragma unroll
                for (int i = 0; i < 3; i++)
                    {
                    const uint32_t q = i * 32 + lane;
                    const uint32_t r = q / 3, c = q - r * 3;
                    if (q < total)
                        out[q] = srow[r * RW + c];
                    }
                }
            else
                {
                for (uint32_t q = lane; q < total; q += 32)
                    {
                    const uint32_t r = q / W, c = q - r * W;
                    out[q] = srow[r * RW + c];
                    }
                }
            }
        }
    }

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

constexpr int GATHER_THREADS = 128;
constexpr int GATHER_CHUNKS = 4;                                            // 32-row chunks per warp
constexpr int GATHER_ROWS_PER_CTA = (GATHER_THREADS / 32) * 32 * GATHER_CHUNKS; // 512

// CTA b owns output rows [512 b, 512 b + 512): the CTAs resident at any time read a window of a few
// hundred thousand consecutive perm entries -- after the bucket pass that is a few buckets of source
// rows, which stay in L2.  A warp moves 4 chunks of 32 rows at once: all loads of a field are issued
// before its first store (the source rows are random, so latency, not bandwidth, is what has to be
// hidden), stores are fully coalesced.
__global__ void __launch_bounds__(GATHER_THREADS)
    k5_gather(const uint32_t* __restrict__ perm, uint64_t n, const __grid_constant__ GatherArgs args)
    {
    const int lane = threadIdx.x & 31;
    const uint64_t warp_row0 = (uint64_t)blockIdx.x * GATHER_ROWS_PER_CTA + (uint64_t)(threadIdx.x >> 5) * (32 * GATHER_CHUNKS);
    if (warp_row0 >= n)
        return;
    uint32_t p[GATHER_CHUNKS], rows[GATHER_CHUNKS];
#pragma unroll
    for (int g = 0; g < GATHER_CHUNKS; g++)
        {
        const uint64_t row0 = warp_row0 + 32 * g;
        rows[g] = row0 >= n ? 0u : (uint32_t)((n - row0) < 32 ? (n - row0) : 32);
        p[g] = lane < rows[g] ? ld_stream_u32(perm + row0 + lane) : 0u;
        }
    for (int fi = 0; fi < args.nfields; fi++)
        {
        const GatherField f = args.f[fi];
        if (f.words == 1)
            {
            const uint32_t* in = reinterpret_cast<const uint32_t*>(f.in);
            uint32_t* out = reinterpret_cast<uint32_t*>(f.out) + warp_row0;
            uint32_t v[GATHER_CHUNKS];
#pragma unroll
            for (int g = 0; g < GATHER_CHUNKS; g++)
                v[g] = lane < rows[g] ? __ldg(in + p[g]) : 0u;
#pragma unroll
            for (int g = 0; g < GATHER_CHUNKS; g++)
                if (lane < rows[g])
                    out[32 * g + lane] = v[g];
            }
        else if (f.words == 3)
            {
            // 96 words per chunk; word q of the chunk = component q % 3 of row q / 3
            const uint32_t* in = reinterpret_cast<const uint32_t*>(f.in);
            uint32_t* out = reinterpret_cast<uint32_t*>(f.out) + warp_row0 * 3;
            uint32_t v[GATHER_CHUNKS][3];
#pragma unroll
            for (int g = 0; g < GATHER_CHUNKS; g++)
#pragma unroll
                for (int i = 0; i < 3; i++)
                    {
                    const uint32_t q = i * 32 + lane;
                    const uint32_t r = q / 3, comp = q - r * 3;
                    const uint32_t src = __shfl_sync(0xffffffffu, p[g], r & 31);
                    v[g][i] = q < rows[g] * 3 ? __ldg(in + (uint64_t)src * 3 + comp) : 0u;
                    }
#pragma unroll
            for (int g = 0; g < GATHER_CHUNKS; g++)
#pragma unroll
                for (int i = 0; i < 3; i++)
                    {
                    const uint32_t q = i * 32 + lane;
                    if (q < rows[g] * 3)
                        out[96 * g + q] = v[g][i];
                    }
            }
        else if (f.words != 0)
            {
            const uint32_t W = f.words;
            const uint32_t* in = reinterpret_cast<const uint32_t*>(f.in);
#pragma unroll
            for (int g = 0; g < GATHER_CHUNKS; g++)
                {
                const uint32_t total = rows[g] * W;
                uint32_t* out = reinterpret_cast<uint32_t*>(f.out) + (warp_row0 + 32 * g) * W;
                for (uint32_t q0 = 0; q0 < total; q0 += 32)
                    {
                    uint32_t q = q0 + lane;
                    uint32_t r = q / W;
                    uint32_t comp = q - r * W;
                    uint32_t src = __shfl_sync(0xffffffffu, p[g], r & 31);
                    if (q < total)
                        out[q] = __ldg(in + (uint64_t)src * W + comp);
                    }
                }
            }
        else
            {
            const uint32_t W = f.row_bytes;
#pragma unroll
            for (int g = 0; g < GATHER_CHUNKS; g++)
                {
                const uint32_t total = rows[g] * W;
                unsigned char* out = f.out + (warp_row0 + 32 * g) * W;
                for (uint32_t q0 = 0; q0 < total; q0 += 32)
                    {
                    uint32_t q = q0 + lane;
                    uint32_t r = q / W;
                    uint32_t comp = q - r * W;
                    uint32_t src = __shfl_sync(0xffffffffu, p[g], r & 31);
                    if (q < total)
                        out[q] = f.in[(uint64_t)src * W + comp];
                    }
                }
            }
        }
    }
    } // namespace

// ---- host side ----------------------------------------------------------------------------------
struct Workspace
    {
    void* base = nullptr;
    size_t bytes = 0;
    };
static Workspace g_sort_ws; // pair buffers + pass tables
static Workspace g_rows_ws; // bucketed copies of keys and fields
static unsigned long long* g_census_host = nullptr; // pinned, 4*256

static int ws_reserve(Workspace& w, size_t need, void** out)
    {
    if (w.bytes < need)
        {
        if (w.base)
            cudaFree(w.base);
        w.base = nullptr;
        w.bytes = 0;
        if (cudaMalloc(&w.base, need) != cudaSuccess)
            {
            set_last_error("cudaMalloc of the sort workspace failed");
            cudaGetLastError();
            return -6;
            }
        w.bytes = need;
        }
    *out = w.base;
    return 0;
    }

void sort_release_workspace()
    {
    if (g_sort_ws.base)
        cudaFree(g_sort_ws.base);
    g_sort_ws = Workspace();
    if (g_rows_ws.base)
        cudaFree(g_rows_ws.base);
    g_rows_ws = Workspace();
    if (g_census_host)
        cudaFreeHost(g_census_host);
    g_census_host = nullptr;
    slot_release_workspace();
    }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// optional phase timing of the last reorder (CUDA events on the caller's stream)
static bool g_phase_prof = false;
static cudaEvent_t g_phase_ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
static int g_phase_n = 0;
static void phase_mark(int i, cudaStream_t st)
    {
    if (!g_phase_prof)
        return;
    if (!g_phase_ev[i])
        cudaEventCreate(&g_phase_ev[i]);
    cudaEventRecord(g_phase_ev[i], st);
    g_phase_n = i + 1;
    }
void dev_reorder_profiling(bool on)
    {
    g_phase_prof = on;
    g_phase_n = 0;
    }
// out[0..3] = census, bucket pass, pair passes, gather (ms); phases that did not run are 0
int dev_reorder_phase_ms(float* out)
    {
    for (int i = 0; i < 4; i++)
        out[i] = 0.f;
    if (g_phase_n < 2)
        return -2;
    if (cudaEventSynchronize(g_phase_ev[g_phase_n - 1]) != cudaSuccess)
        return -1;
    for (int i = 0; i + 1 < g_phase_n; i++)
        cudaEventElapsedTime(&out[i], g_phase_ev[i], g_phase_ev[i + 1]);
    return 0;
    }

// phase marks of the slot path (kernels_slot.cu): scatter = "bucket pass", no pair passes, placement = "gather"
static void slot_mark(int i, cudaStream_t st)
    {
    if (i == 0)
        {
        phase_mark(2, st);
        phase_mark(3, st);
        }
    else
        phase_mark(4, st);
    }

static int rank_mode()
    {
    // read on every call (cheap) so that tests can switch it
    const char* e = getenv("PGSD_B200_RANK_MODE");
    return (e && !strcmp(e, "match")) ? RANK_MATCH : RANK_BALLOT;
    }

static int sort_setup()
    {
    static bool attr_done = false;
    if (!attr_done)
        {
        cudaFuncSetAttribute(k4_scatter<true, RANK_BALLOT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM);
        cudaFuncSetAttribute(k4_scatter<false, RANK_BALLOT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM);
        cudaFuncSetAttribute(k4_scatter<true, RANK_MATCH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM);
        cudaFuncSetAttribute(k4_scatter<false, RANK_MATCH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM);
        cudaFuncSetAttribute(k4_scatter<true, RANK_BALLOT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM);
        cudaFuncSetAttribute(k4_scatter<false, RANK_BALLOT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM);
        cudaFuncSetAttribute(k4_scatter<true, RANK_MATCH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM);
        cudaFuncSetAttribute(k4_scatter<false, RANK_MATCH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM);
        cudaFuncSetAttribute(k4_bucket_rows<RANK_BALLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROWS_SMEM);
        cudaFuncSetAttribute(k4_bucket_rows<RANK_MATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROWS_SMEM);
        const int big = 200 * 1024;
        cudaFuncSetAttribute(k4_bucket_aos<4, RANK_BALLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k4_bucket_aos<2, RANK_BALLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k4_bucket_aos<1, RANK_BALLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k4_bucket_aos<4, RANK_MATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k4_bucket_aos<2, RANK_MATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k4_bucket_aos<1, RANK_MATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k5_gather_aos<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k5_gather_aos<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k5_gather_aos<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k5_gather_aos<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        attr_done = true;
        }
    if (!g_census_host && cudaHostAlloc((void**)&g_census_host, 4 * RADIX * 8, cudaHostAllocDefault) != cudaSuccess)
        {
        set_last_error("cudaHostAlloc failed");
        cudaGetLastError();
        return -6;
        }
    return 0;
    }

// pass tables shared by the pair passes and the bucket pass (sized for the finest tiling)
constexpr int MIN_TILE = SORT_THREADS;
struct PassTables
    {
    uint32_t* counts;
    unsigned long long* row_total;
    unsigned long long* digit_base;
    unsigned long long* census;
    };
static size_t pass_tables_bytes(uint64_t n)
    {
    const size_t ntiles = (size_t)((n + MIN_TILE - 1) / MIN_TILE);
    return align_up((size_t)RADIX * ntiles * 4, 256) + 2 * RADIX * 8 + 4 * RADIX * 8;
    }
static PassTables pass_tables_at(unsigned char* p, uint64_t n)
    {
    const size_t ntiles = (size_t)((n + MIN_TILE - 1) / MIN_TILE);
    PassTables t;
    t.counts = (uint32_t*)p;
    t.row_total = (unsigned long long*)(p + align_up((size_t)RADIX * ntiles * 4, 256));
    t.digit_base = t.row_total + RADIX;
    t.census = t.digit_base + RADIX;
    return t;
    }

// Census of the key bits (one host round trip): which bytes vary, and the bit length of the
// varying part.  passes[] lists the varying bytes, least significant first.
struct KeyPlan
    {
    int passes[4];
    int npass;
    int topbit; // keys differ only in bits [0, topbit)
    uint32_t key_const; // the bits all keys share at and above `topbit`
    };
static int key_census(uint64_t n, const uint32_t* keys, const PassTables& t, cudaStream_t st, KeyPlan* plan)
    {
    uint32_t* bits = reinterpret_cast<uint32_t*>(t.census);
    cudaMemsetAsync(bits, 0, 4, st);
    cudaMemsetAsync(bits + 1, 0xff, 4, st);
    int census_grid = dev_sm_count() * 4;
    uint64_t want = (n / 4 + 511) / 512;
    if ((uint64_t)census_grid > want)
        census_grid = want ? (int)want : 1;
    k4_digit_census<<<census_grid, 512, 0, st>>>(keys, n, bits);
    dev_stats().kernel_launches++;
    uint32_t* host = reinterpret_cast<uint32_t*>(g_census_host);
    cudaMemcpyAsync(host, bits, 8, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess)
        {
        set_last_error(std::string("sort_ids census: ") + cudaGetErrorString(cudaGetLastError()));
        return -1;
        }
    const uint32_t varying = host[0] & ~host[1];
    plan->npass = 0;
    plan->topbit = 0;
    for (int b = 0; b < 4; b++)
        {
        uint32_t x = (varying >> (8 * b)) & 255u;
        if (x)
            {
            plan->passes[plan->npass++] = b;
            int bits_b = 0; // the keys agree on this byte above bit `bits_b`
            while (x)
                {
                bits_b++;
                x >>= 1;
                }
            plan->topbit = 8 * b + bits_b;
            }
        }
    plan->key_const = plan->topbit >= 32 ? 0u : (host[1] & ~((1u << plan->topbit) - 1u));
    return 0;
    }

template <bool SEG>
static void launch_scatter(bool first, const uint32_t* kin, const uint32_t* iin, uint32_t* kout, uint32_t* iout,
                           uint64_t n, int shift, uint32_t ntiles, const uint32_t* counts,
                           const unsigned long long* digit_base, const SegTables& seg, cudaStream_t st)
    {
    const bool m = rank_mode() == RANK_MATCH;
    if (first)
        {
        if (m)
            k4_scatter<true, RANK_MATCH, SEG><<<ntiles, SORT_THREADS, SCATTER_SMEM, st>>>(kin, iin, kout, iout, n, shift, ntiles, counts, digit_base, seg);
        else
            k4_scatter<true, RANK_BALLOT, SEG><<<ntiles, SORT_THREADS, SCATTER_SMEM, st>>>(kin, iin, kout, iout, n, shift, ntiles, counts, digit_base, seg);
        }
    else
        {
        if (m)
            k4_scatter<false, RANK_MATCH, SEG><<<ntiles, SORT_THREADS, SCATTER_SMEM, st>>>(kin, iin, kout, iout, n, shift, ntiles, counts, digit_base, seg);
        else
            k4_scatter<false, RANK_BALLOT, SEG><<<ntiles, SORT_THREADS, SCATTER_SMEM, st>>>(kin, iin, kout, iout, n, shift, ntiles, counts, digit_base, seg);
        }
    }

// LSD pair passes over the planned bytes.  kbuf/ibuf: two ping-pong buffers each.
static void pair_passes(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm, const KeyPlan& plan,
                        uint32_t* const kbuf[2], uint32_t* const ibuf[2], const PassTables& t, cudaStream_t st)
    {
    const uint32_t ntiles = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE);
    const uint32_t* kin = keys;
    const uint32_t* iin = nullptr;
    int cur = 0;
    for (int pi = 0; pi < plan.npass; pi++)
        {
        const int shift = plan.passes[pi] * 8;
        const bool last = (pi == plan.npass - 1);
        uint32_t* kout = (last && keys_sorted) ? keys_sorted : kbuf[cur];
        uint32_t* iout = (last && perm) ? perm : ibuf[cur];
        k4_tile_histogram<SORT_ITEMS><<<ntiles, SORT_THREADS, 0, st>>>(kin, n, shift, ntiles, t.counts);
        k4_row_scan<<<RADIX, 256, 0, st>>>(t.counts, ntiles, t.row_total);
        k4_digit_base<<<1, RADIX, 0, st>>>(t.row_total, t.digit_base);
        launch_scatter<false>(pi == 0, kin, iin, kout, iout, n, shift, ntiles, t.counts, t.digit_base, SegTables(), st);
        dev_stats().kernel_launches += 4;
        kin = kout;
        iin = iout;
        cur ^= 1;
        }
    }

// Segmented LSD passes on keys already grouped by the bucket digit (bits >= bshift): only the key
// bytes below bshift are sorted, inside each bucket; the pass over the bucket digit itself is saved.
// bucket_base: start of every bucket (the bucket pass's digit_base).  ws: workspace for the tables.
static size_t seg_tables_bytes(uint64_t n)
    {
    const size_t tmax = (size_t)(n / SORT_TILE) + RADIX + 1;
    return align_up((size_t)RADIX * tmax * 4, 256) + align_up(3 * tmax * 4, 256) + align_up(2 * (RADIX + 1) * 4, 256)
           + 2 * (size_t)RADIX * RADIX * 4 + 256;
    }

static int pair_passes_segmented(uint64_t n, const uint32_t* keys_b, uint32_t* keys_sorted, uint32_t* perm_b,
                                 const KeyPlan& plan, int bshift, const unsigned long long* bucket_base,
                                 uint32_t* const kbuf[2], uint32_t* const ibuf[2], unsigned char* ws, cudaStream_t st)
    {
    const uint32_t tmax = (uint32_t)(n / SORT_TILE) + RADIX + 1;
    unsigned char* p = ws;
    uint32_t* counts = (uint32_t*)p;
    p += align_up((size_t)RADIX * tmax * 4, 256);
    uint32_t* tile_begin = (uint32_t*)p;
    uint32_t* tile_cnt = tile_begin + tmax;
    uint32_t* tile_bkt = tile_cnt + tmax;
    p += align_up(3 * (size_t)tmax * 4, 256);
    uint32_t* bstart = (uint32_t*)p;
    uint32_t* tile_first = bstart + RADIX + 1;
    p += align_up(2 * (RADIX + 1) * 4, 256);
    uint32_t* btot = (uint32_t*)p;
    uint32_t* bbase = btot + RADIX * RADIX;
    uint32_t* ntiles_dev = bbase + RADIX * RADIX;

    k4s_setup<<<1, RADIX, 0, st>>>(bucket_base, n, bstart, tile_first, tile_begin, tile_cnt, tile_bkt, ntiles_dev);
    dev_stats().kernel_launches++;
    SegTables seg;
    seg.tile_begin = tile_begin;
    seg.tile_cnt = tile_cnt;
    seg.tile_bkt = tile_bkt;
    seg.ntiles = ntiles_dev;
    seg.bstart = bstart;
    seg.bbase = bbase;
    seg.stride = tmax;

    int low[4], nlow = 0;
    for (int pi = 0; pi < plan.npass; pi++)
        if (plan.passes[pi] * 8 < bshift)
            low[nlow++] = plan.passes[pi];
    const uint32_t* kin = keys_b;
    const uint32_t* iin = nullptr;
    int cur = 0;
    for (int pi = 0; pi < nlow; pi++)
        {
        const int shift = low[pi] * 8;
        const bool last = (pi == nlow - 1);
        uint32_t* kout = (last && keys_sorted) ? keys_sorted : kbuf[cur];
        uint32_t* iout = last ? perm_b : ibuf[cur];
        k4s_histogram<<<tmax, SORT_THREADS, 0, st>>>(kin, shift, seg, counts);
        k4s_scan<<<RADIX, 256, 0, st>>>(counts, tile_first, tmax, btot);
        k4s_base<<<RADIX, RADIX, 0, st>>>(btot, bbase);
        launch_scatter<true>(pi == 0, kin, iin, kout, iout, n, shift, tmax, counts, nullptr, seg, st);
        dev_stats().kernel_launches += 4;
        kin = kout;
        iin = iout;
        cur ^= 1;
        }
    if (nlow == 0)
        {
        // nothing varies below the bucket digit: the bucketed order is the sorted order
        k4_iota<<<dev_sm_count() * 8, 256, 0, st>>>(perm_b, n);
        dev_stats().kernel_launches++;
        if (keys_sorted)
            cudaMemcpyAsync(keys_sorted, keys_b, n * 4, cudaMemcpyDeviceToDevice, st);
        }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }

static int check_n(uint64_t n)
    {
    if (n >= 0xffffffffull)
        {
        set_last_error("sort_ids: n must be < 2^32 - 1 (uint32 permutation)");
        return -2;
        }
    return 0;
    }

int dev_sort_ids(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm, void* stream_v)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    if ((rc = check_n(n)) != 0)
        return rc;
    if (n == 0)
        return 0;
    if (keys == nullptr)
        return -2;
    cudaStream_t st = (cudaStream_t)stream_v;
    if ((rc = sort_setup()) != 0)
        return rc;
    // workspace: keys A/B, idx A/B, pass tables
    const size_t arr = align_up((size_t)n * 4, 256);
    void* ws = nullptr;
    rc = ws_reserve(g_sort_ws, 4 * arr + pass_tables_bytes(n), &ws);
    if (rc != 0)
        return rc;
    unsigned char* p = (unsigned char*)ws;
    uint32_t* kbuf[2] = { (uint32_t*)p, (uint32_t*)(p + arr) };
    uint32_t* ibuf[2] = { (uint32_t*)(p + 2 * arr), (uint32_t*)(p + 3 * arr) };
    const PassTables t = pass_tables_at(p + 4 * arr, n);
    KeyPlan plan;
    if ((rc = key_census(n, keys, t, st, &plan)) != 0)
        return rc;
    pair_passes(n, keys, keys_sorted, perm, plan, kbuf, ibuf, t, st);
    if (plan.npass == 0)
        {
        // all keys equal: the stable order is the identity
        if (perm)
            {
            k4_iota<<<dev_sm_count() * 8, 256, 0, st>>>(perm, n);
            dev_stats().kernel_launches++;
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

// K4 + K5 as one operation.  Frames large enough for the gather to fall out of L2 first get the
// bucket pass (rows grouped by the top 8 significant key bits, payload moved once, coalesced);
// the pair passes then sort (key, position in the bucketed copy) and the gather reads bucket-local.
//   npass == 0 : identity.   npass == 1 : the bucket pass writes the final order directly.
static uint64_t bucket_min_rows()
    {
    const char* e = getenv("PGSD_B200_BUCKET_MIN_ROWS");
    return e ? (uint64_t)atoll(e) : (1ull << 20);
    }

// Slot path attempts.  First with the key range GUESSED from n (dense ids 0..n-1: no census, no host round
// trip before the kernels; k6_slot_hist verifies the guess).  If only the guess was wrong (ids with an offset
// or gaps), once more with the range the census measures.  Duplicates: *done stays 0, the caller sorts.
static int reorder_try_slot(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm, int nfields,
                            const ReorderField* fields, void* stream_v, bool timed, int* done)
    {
    *done = 0;
    cudaStream_t st = (cudaStream_t)stream_v;
    int rc, miss = 0;
    const char* eg = getenv("PGSD_B200_SLOT_GUESS");
    if (n >= 2 && !(eg && eg[0] == '0'))
        {
        int tg = 0;
        while (tg < 32 && ((n - 1) >> tg) != 0)
            tg++;
        if (timed)
            {
            phase_mark(0, st);
            phase_mark(1, st);
            }
        if ((rc = dev_reorder_slot(n, keys, keys_sorted, perm, nfields, fields, tg, 1, 0u, stream_v, done, &miss, timed ? slot_mark : nullptr)) != 0)
            return rc;
        if (*done || !miss)
            return 0; // finished, or not applicable / duplicates: a measured range would not change that
        }
    if ((rc = sort_setup()) != 0)
        return rc;
    void* ws = nullptr;
    const size_t arr = align_up((size_t)n * 4, 256);
    if ((rc = ws_reserve(g_sort_ws, 4 * arr + pass_tables_bytes(n), &ws)) != 0)
        return rc;
    const PassTables t = pass_tables_at((unsigned char*)ws + 4 * arr, n);
    KeyPlan plan;
    if (timed)
        phase_mark(0, st);
    if ((rc = key_census(n, keys, t, st, &plan)) != 0)
        return rc;
    if (timed)
        phase_mark(1, st);
    if (plan.npass == 0)
        return 0; // all keys equal: the caller's identity path
    return dev_reorder_slot(n, keys, keys_sorted, perm, nfields, fields, plan.topbit, 0, plan.key_const, stream_v, done, &miss, timed ? slot_mark : nullptr);
    }

static uint64_t slot_min_rows()
    {
    const char* e = getenv("PGSD_B200_SLOT_MIN_ROWS");
    return e ? (uint64_t)atoll(e) : 2048ull;
    }

int dev_reorder_rows(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm, int nfields,
                     const ReorderField* fields, void* stream_v)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    if ((rc = check_n(n)) != 0)
        return rc;
    if (n == 0)
        return 0;
    if (keys == nullptr || nfields < 0 || (nfields > 0 && fields == nullptr))
        return -2;
    cudaStream_t st = (cudaStream_t)stream_v;
    bool rows_ok = nfields + 1 <= MAX_ROW_FIELDS;
    size_t payload = 0;
    for (int i = 0; i < nfields; i++)
        {
        const ReorderField& f = fields[i];
        if (f.row_bytes == 0 || f.row_bytes > 1024 || f.in == nullptr || f.out == nullptr || f.in == f.out)
            {
            set_last_error("reorder: bad field (row_bytes 1..1024, in/out non-null and distinct)");
            return -2;
            }
        if (f.row_bytes % 4 != 0 || (((uintptr_t)f.in | (uintptr_t)f.out) % 4) != 0 || f.row_bytes / 4 > ROWS_STAGE_WORDS)
            rows_ok = false;
        payload += align_up((size_t)n * f.row_bytes, 256);
        }
    if (!rows_ok || n < bucket_min_rows())
        {
        // small frames with unique ids: the slot path (kernels_slot.cu) is 4-5 launches instead of ~10
        if (rows_ok && n >= slot_min_rows())
            {
            int done = 0;
            if ((rc = reorder_try_slot(n, keys, keys_sorted, perm, nfields, fields, stream_v, false, &done)) != 0)
                return rc;
            if (done)
                return 0;
            }
        // small or oddly shaped frames: pair sort + gather from the caller's arrays
        uint32_t* p = perm;
        if (p == nullptr)
            {
            void* ws = nullptr;
            if ((rc = ws_reserve(g_rows_ws, align_up((size_t)n * 4, 256), &ws)) != 0)
                return rc;
            p = (uint32_t*)ws;
            }
        if ((rc = dev_sort_ids(n, keys, keys_sorted, p, stream_v)) != 0)
            return rc;
        return dev_gather(n, p, nfields, fields, stream_v);
        }
    if ((rc = sort_setup()) != 0)
        return rc;
    const size_t arr = align_up((size_t)n * 4, 256);
    void* ws = nullptr;
    if ((rc = ws_reserve(g_sort_ws, 4 * arr + pass_tables_bytes(n), &ws)) != 0)
        return rc;
    unsigned char* p = (unsigned char*)ws;
    uint32_t* kbuf[2] = { (uint32_t*)p, (uint32_t*)(p + arr) };
    uint32_t* ibuf[2] = { (uint32_t*)(p + 2 * arr), (uint32_t*)(p + 3 * arr) };
    const PassTables t = pass_tables_at(p + 4 * arr, n);
    // unique ids (the normal case): two passes over the rows, no ranking (kernels_slot.cu); duplicates are
    // detected on the device and fall through to the stable general path below
        {
        int done = 0;
        if ((rc = reorder_try_slot(n, keys, keys_sorted, perm, nfields, fields, stream_v, true, &done)) != 0)
            return rc;
        if (done)
            return 0;
        }
    KeyPlan plan;
    phase_mark(0, st);
    if ((rc = key_census(n, keys, t, st, &plan)) != 0)
        return rc;
    phase_mark(1, st);
    if (plan.npass == 0)
        {
        if (perm)
            {
            k4_iota<<<dev_sm_count() * 8, 256, 0, st>>>(perm, n);
            dev_stats().kernel_launches++;
            }
        if (keys_sorted && keys_sorted != keys)
            cudaMemcpyAsync(keys_sorted, keys, n * 4, cudaMemcpyDeviceToDevice, st);
        for (int i = 0; i < nfields; i++)
            cudaMemcpyAsync(fields[i].out, fields[i].in, n * (size_t)fields[i].row_bytes, cudaMemcpyDeviceToDevice, st);
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
        }
    const bool direct = plan.npass == 1; // one varying byte: the bucket pass is the whole sort
    uint32_t row_words = perm ? 1u : 0u;
    for (int i = 0; i < nfields; i++)
        row_words += fields[i].row_bytes / 4;
    const char* lay = getenv("PGSD_B200_BUCKET_LAYOUT");
    if (!direct && row_words >= 1 && row_words <= AOS_MAX_ROW_WORDS && !(lay && !strcmp(lay, "soa")))
        {
        // ---- fast path: interleaved bucketed copy
        const size_t aos_bytes = align_up((size_t)n * row_words * 4, 256);
        void* rws = nullptr;
        if ((rc = ws_reserve(g_rows_ws, 2 * arr + aos_bytes + seg_tables_bytes(n), &rws)) != 0)
            return rc;
        unsigned char* rp = (unsigned char*)rws;
        uint32_t* keys_b = (uint32_t*)rp;
        uint32_t* perm_b = (uint32_t*)(rp + arr);
        uint32_t* aos = (uint32_t*)(rp + 2 * arr);
        AosArgs aa;
        memset(&aa, 0, sizeof(aa));
        uint32_t off = 0;
        int nf = 0;
        for (int i = 0; i < nfields; i++)
            {
            aa.f[nf].in = (const uint32_t*)fields[i].in;
            aa.f[nf].out = (uint32_t*)fields[i].out;
            aa.f[nf].words = fields[i].row_bytes / 4;
            aa.f[nf].off = off;
            off += aa.f[nf].words;
            nf++;
            }
        if (perm)
            {
            aa.f[nf].in = nullptr; // iota
            aa.f[nf].out = perm;
            aa.f[nf].words = 1;
            aa.f[nf].off = off;
            off += 1;
            nf++;
            }
        aa.nfields = nf;
        aa.row_words = row_words;
        const int bshift = plan.topbit > 8 ? plan.topbit - 8 : 0;
        const int items = row_words <= 10 ? 4 : (row_words <= 22 ? 2 : 1);
        const uint32_t tile = (uint32_t)SORT_THREADS * items;
        const uint32_t rtiles = (uint32_t)((n + tile - 1) / tile);
        const bool m = rank_mode() == RANK_MATCH;
        const size_t stage_b = (size_t)tile * row_words * 4;
#define PGSD_LAUNCH_AOS(IT)                                                                                      \
    {                                                                                                              \
    k4_tile_histogram<IT><<<rtiles, SORT_THREADS, 0, st>>>(keys, n, bshift, rtiles, t.counts);                    \
    k4_row_scan<<<RADIX, 256, 0, st>>>(t.counts, rtiles, t.row_total);                                            \
    k4_digit_base<<<1, RADIX, 0, st>>>(t.row_total, t.digit_base);                                                \
    if (m)                                                                                                         \
        k4_bucket_aos<IT, RANK_MATCH><<<rtiles, SORT_THREADS, aos_smem_fixed<IT>() + stage_b, st>>>(              \
            keys, keys_b, aos, n, bshift, rtiles, t.counts, t.digit_base, aa);                                     \
    else                                                                                                           \
        k4_bucket_aos<IT, RANK_BALLOT><<<rtiles, SORT_THREADS, aos_smem_fixed<IT>() + stage_b, st>>>(             \
            keys, keys_b, aos, n, bshift, rtiles, t.counts, t.digit_base, aa);                                     \
    }
        if (items == 4)
            PGSD_LAUNCH_AOS(4)
        else if (items == 2)
            PGSD_LAUNCH_AOS(2)
        else
            PGSD_LAUNCH_AOS(1)
#undef PGSD_LAUNCH_AOS
        dev_stats().kernel_launches += 4;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            {
            set_last_error(std::string("reorder bucket pass: ") + cudaGetErrorString(e));
            return -1;
            }
        phase_mark(2, st);
        const char* sg = getenv("PGSD_B200_SEGMENTED");
        if (!(sg && sg[0] == '0'))
            {
            // the bucket starts are in t.digit_base (left there by the bucket pass)
            if (pair_passes_segmented(n, keys_b, keys_sorted, perm_b, plan, bshift, t.digit_base, kbuf, ibuf,
                                      rp + 2 * arr + aos_bytes, st) != 0)
                {
                set_last_error(std::string("reorder segmented passes: ") + cudaGetErrorString(cudaGetLastError()));
                return -1;
                }
            }
        else
            pair_passes(n, keys_b, keys_sorted, perm_b, plan, kbuf, ibuf, t, st);
        phase_mark(3, st);
        const uint64_t blocks = (n + GA_ROWS_PER_CTA - 1) / GA_ROWS_PER_CTA;
        const size_t gsm = (size_t)GA_ROWS_PER_CTA * row_words * 4;
        const char* pf = getenv("PGSD_B200_GATHER_PREFETCH");
        const bool prefetch = !(pf && pf[0] == '0');
        if (row_words % 2 == 0)
            {
            if (prefetch)
                k5_gather_aos<2, true><<<(unsigned)blocks, GA_THREADS, gsm, st>>>(perm_b, n, aos, aa);
            else
                k5_gather_aos<2, false><<<(unsigned)blocks, GA_THREADS, gsm, st>>>(perm_b, n, aos, aa);
            }
        else
            {
            if (prefetch)
                k5_gather_aos<1, true><<<(unsigned)blocks, GA_THREADS, gsm, st>>>(perm_b, n, aos, aa);
            else
                k5_gather_aos<1, false><<<(unsigned)blocks, GA_THREADS, gsm, st>>>(perm_b, n, aos, aa);
            }
        dev_stats().kernel_launches++;
        phase_mark(4, st);
        e = cudaGetLastError();
        if (e != cudaSuccess)
            {
            set_last_error(std::string("reorder gather: ") + cudaGetErrorString(e));
            return -1;
            }
        return 0;
        }
    // bucketed copies: keys, original index (only if the caller wants perm), fields
    unsigned char* rp = nullptr;
    if (!direct)
        {
        void* rws = nullptr;
        if ((rc = ws_reserve(g_rows_ws, 3 * arr + payload, &rws)) != 0)
            return rc;
        rp = (unsigned char*)rws;
        }
    uint32_t* keys_b = direct ? keys_sorted : (uint32_t*)rp;
    uint32_t* orig_b = direct ? perm : (perm ? (uint32_t*)(rp + arr) : nullptr);
    uint32_t* perm_b = direct ? nullptr : (uint32_t*)(rp + 2 * arr);
    RowArgs ra;
    memset(&ra, 0, sizeof(ra));
    std::vector<ReorderField> gf((size_t)nfields + 1);
    size_t off = 3 * arr;
    int nr = 0;
    for (int i = 0; i < nfields; i++)
        {
        ra.f[nr].in = (const uint32_t*)fields[i].in;
        ra.f[nr].words = fields[i].row_bytes / 4;
        if (direct)
            ra.f[nr].out = (uint32_t*)fields[i].out;
        else
            {
            ra.f[nr].out = (uint32_t*)(rp + off);
            gf[(size_t)i].in = rp + off;
            gf[(size_t)i].out = fields[i].out;
            gf[(size_t)i].row_bytes = fields[i].row_bytes;
            off += align_up((size_t)n * fields[i].row_bytes, 256);
            }
        nr++;
        }
    int ngf = nfields;
    if (orig_b)
        {
        ra.f[nr].in = nullptr; // iota
        ra.f[nr].words = 1;
        ra.f[nr].out = orig_b;
        nr++;
        if (!direct)
            {
            gf[(size_t)ngf].in = orig_b;
            gf[(size_t)ngf].out = perm;
            gf[(size_t)ngf].row_bytes = 4;
            ngf++;
            }
        }
    ra.nfields = nr;
    const int bshift = direct ? plan.passes[0] * 8 : (plan.topbit > 8 ? plan.topbit - 8 : 0);
    const uint32_t rtiles = (uint32_t)((n + ROWS_TILE - 1) / ROWS_TILE);
    k4_tile_histogram<ROWS_ITEMS><<<rtiles, SORT_THREADS, 0, st>>>(keys, n, bshift, rtiles, t.counts);
    k4_row_scan<<<RADIX, 256, 0, st>>>(t.counts, rtiles, t.row_total);
    k4_digit_base<<<1, RADIX, 0, st>>>(t.row_total, t.digit_base);
    if (rank_mode() == RANK_MATCH)
        k4_bucket_rows<RANK_MATCH><<<rtiles, SORT_THREADS, ROWS_SMEM, st>>>(keys, keys_b, n, bshift, rtiles, t.counts, t.digit_base, ra);
    else
        k4_bucket_rows<RANK_BALLOT><<<rtiles, SORT_THREADS, ROWS_SMEM, st>>>(keys, keys_b, n, bshift, rtiles, t.counts, t.digit_base, ra);
    dev_stats().kernel_launches += 4;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        {
        set_last_error(std::string("reorder bucket pass: ") + cudaGetErrorString(e));
        return -1;
        }
    phase_mark(2, st);
    if (direct)
        return 0;
    pair_passes(n, keys_b, keys_sorted, perm_b, plan, kbuf, ibuf, t, st);
    e = cudaGetLastError();
    if (e != cudaSuccess)
        {
        set_last_error(std::string("reorder pair passes: ") + cudaGetErrorString(e));
        return -1;
        }
    phase_mark(3, st);
    rc = dev_gather(n, perm_b, ngf, gf.data(), stream_v);
    phase_mark(4, st);
    return rc;
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
        uint64_t blocks = (n + GATHER_ROWS_PER_CTA - 1) / GATHER_ROWS_PER_CTA;
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
