// kernels_cluster.cu -- K4/K5 "cluster path": the reorder of a decoded frame with UNIQUE particle ids in exactly two
// streaming passes over the rows, no per-row global atomics and no histogram pass.
//
// Same oracle as the other reorder paths:  o = numpy.argsort(ids, kind='stable');  out_f = in_f[o]  (SURVEY.md
// section 8 a19; ids come from log/particles/id, /root/reference/pgsd/pgsd/hoomd.py:885-893).  With unique keys the
// stable order is the key order and a row's rank inside an aligned key range is its low key bits.
//
//   k7_coarse_scatter  software write-combining partition into COARSE buckets of 2^LC consecutive ids (LC = 15: 512
//                      buckets at 16 Mi particles).  A tile of T rows (all fields) is staged SoA -> shared memory by
//                      one TMA bulk copy per field; the tile's rows are counted per bucket in shared memory (warp-
//                      aggregated), ONE global atomic per (tile, non-empty bucket) reserves a run in the bucket's
//                      region of the interleaved copy, and the rows leave sorted by bucket as 40-byte records, so a
//                      tile writes runs of T / buckets records (320 B at T = 4096) instead of single records.
//                      Bucket regions have the fixed capacity 2^LC records: the cursors double as the counts, which
//                      is why no histogram has to run first.  A run that does not fit proves duplicate ids -> flag.
//   k7_cluster_place   one thread-block CLUSTER of 8 CTAs per coarse bucket.  CTA r owns the slots
//                      [r * 2^LF, (r + 1) * 2^LF) of the bucket (LF = 12: 4096 records = 160 KB of shared memory).
//                      Every CTA streams one eighth of the bucket's records from global memory into registers and
//                      stores each record at slot = key & (2^LF - 1) of the owner CTA's shared memory (distributed
//                      shared memory: mapa + st.shared::cluster).  After a cluster barrier every CTA holds its 4096
//                      rows in id order and writes the fields back SoA, fully coalesced.  Missing ids (gaps) are
//                      compacted with an occupancy bitmap; a bucket whose occupied slots are fewer than its records
//                      held two records with one id -> flag (the caller falls back to the stable general path).
//
// Bytes moved: (40 + 40) + (40 + 40) = 160 B/particle for the 40-byte SPH row -- two passes, the minimum for a
// bucketed reorder whose frame does not fit on chip -- against 164 B + one atomic per row of the slot path
// (kernels_slot.cu) whose second pass can only take buckets of <= 4096 ids and therefore needs 4096..16384
// concurrently filling buckets in the first one.
//
// sm_100a only (thread-block clusters, distributed shared memory, cp.async.bulk + mbarrier).  No CPU fallback.
#include "slot_common.cuh"

#include <cstdlib>

namespace pgsdb
{
using namespace slotk;
namespace
    {
constexpr int CL_SIZE = 8;           // CTAs per cluster (portable maximum)
constexpr int CL_MAX_BUCKETS = 2048; // coarse buckets (shared-memory tables of k7_coarse_scatter)

__device__ __forceinline__ unsigned cl_lanemask_lt()
    {
    unsigned m;
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
    }
__device__ __forceinline__ uint32_t cl_ctarank()
    {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
    }
__device__ __forceinline__ void cl_sync()
    {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
__device__ __forceinline__ uint32_t cl_map(uint32_t smem_addr, uint32_t rank)
    {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
    }
__device__ __forceinline__ void cl_st_v2(uint32_t addr, uint32_t x, uint32_t y)
    {
    asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
    }
__device__ __forceinline__ void cl_st_u32(uint32_t addr, uint32_t x)
    {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(x) : "memory");
    }
__device__ __forceinline__ uint2 cl_ldg_v2(const uint2* p)
    {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
    }
__device__ __forceinline__ uint32_t cl_ldg_u32(const uint32_t* p)
    {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
    }

// flag words (device): [0] pass 1: a run did not fit its bucket region (more than 2^LC ids in one bucket: duplicates)
//                      [1] pass 2: 1 = two records of a CTA's slot range share a slot (duplicate ids), 3 = a bulk copy never arrived
//                      [2] pass 1: a key lies outside the range guessed from n or measured by the census
//                      [3] pass 2: more records than slots in one CTA's range (duplicate ids)

// ---- pass 1: rows -> records, partitioned into coarse buckets, one contiguous run per (tile, bucket) ----------------
// Shared memory: [field tiles: T * W words each + skew][sdst u32 T][order u16 T][scnt nb][sofs nb][sbase nb]
template <int T, int NT>
__global__ void __launch_bounds__(NT) k7_coarse_scatter(uint64_t n, int LC, uint32_t nb, uint32_t key_const,
                                                       uint32_t* __restrict__ cursor, uint32_t cstride, uint32_t* __restrict__ rec,
                                                       uint32_t* __restrict__ flag, int aggregate,
                                                       const __grid_constant__ SlotArgs args)
    {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar_mem;
    __shared__ uint32_t col[SLOT_MAX_ROW_WORDS];
    __shared__ uint32_t wsum[32];
    uint32_t* raw = reinterpret_cast<uint32_t*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NW = NT / 32;
    constexpr int PER = T / NT;
    const uint64_t tile0 = (uint64_t)blockIdx.x * T;
    const uint32_t tile_n = (uint32_t)((n - tile0) < (uint64_t)T ? (n - tile0) : T);
    const uint32_t RW = args.row_words;
    const uint32_t capc = 1u << LC;

    uint32_t key[PER];
#pragma unroll
    for (int k = 0; k < PER; k++)
        {
        const uint32_t r = (uint32_t)tid + (uint32_t)k * NT;
        key[k] = r < tile_n ? __ldg(args.f[0].in + tile0 + r) : 0u;
        }
    uint32_t fbase = 0, tx_bytes = 0;
    for (int fi = 0; fi < args.nfields; fi++)
        {
        const uint32_t W = args.f[fi].words;
        const bool real = args.f[fi].in != nullptr;
        if (tid < (int)W)
            col[args.f[fi].off + tid] = real ? (((fbase + (uint32_t)tid) << 8) | W) : 0u;
        if (real)
            {
            fbase += (uint32_t)T * W + SLOT_SKEW;
            if (fi > 0)
                tx_bytes += (uint32_t)T * W * 4u;
            }
        }
    uint32_t* sdst = raw + fbase;                            // T : row -> record index in the copy (~0: dropped)
    uint16_t* order = reinterpret_cast<uint16_t*>(sdst + T); // T : position sorted by bucket -> row
    uint32_t* scnt = reinterpret_cast<uint32_t*>(order + T); // nb: rows of this tile per bucket
    uint32_t* sofs = scnt + nb;                              // nb: first sorted position of the bucket
    uint32_t* sbase = sofs + nb;                             // nb: first position of the tile's run in the bucket region
    for (uint32_t i = tid; i < nb; i += NT)
        scnt[i] = 0;
    const bool bulk = args.bulk && tile_n == (uint32_t)T && tx_bytes != 0;
    const uint32_t bar = smem_u32(&bar_mem);
    if (bulk && tid == 0)
        mbar_init(bar, 1);
    __syncthreads();
    // (1) the payload tiles start their way global -> shared
    if (bulk)
        {
        if (tid == 0)
            {
            mbar_expect_tx(bar, tx_bytes);
            uint32_t fb = (uint32_t)T + SLOT_SKEW;
            for (int fi = 1; fi < args.nfields; fi++)
                {
                const uint32_t W = args.f[fi].words;
                if (args.f[fi].in == nullptr)
                    continue;
                bulk_g2s(smem_u32(raw + fb), args.f[fi].in + tile0 * W, (uint32_t)T * W * 4u, bar);
                fb += (uint32_t)T * W + SLOT_SKEW;
                }
            }
        }
    else
        {
        uint32_t fb = (uint32_t)T + SLOT_SKEW;
        for (int fi = 1; fi < args.nfields; fi++)
            {
            const uint32_t W = args.f[fi].words;
            if (args.f[fi].in == nullptr)
                continue;
            const uint32_t* in = args.f[fi].in + tile0 * W;
            const uint32_t total = tile_n * W;
            for (uint32_t q = tid; q < total; q += NT)
                raw[fb + q] = __ldg(in + q);
            fb += (uint32_t)T * W + SLOT_SKEW;
            }
        }
    // (2) rank of every row among the tile's rows of the same bucket.  Lanes with equal buckets share one
    // shared-memory atomic (sorted or clustered ids put a whole warp into one bucket).
    uint32_t bk[PER], rk[PER];
    bool oob = false;
#pragma unroll
    for (int k = 0; k < PER; k++)
        {
        const uint32_t r = (uint32_t)tid + (uint32_t)k * NT;
        uint32_t b = 0xffffffffu;
        if (r < tile_n)
            {
            b = (key[k] ^ key_const) >> LC; // key_const: the bits all keys share above the varying ones (census), 0 when guessed
            if (b >= nb)
                {
                oob = true;
                b = 0xffffffffu;
                }
            }
        bk[k] = b;
        rk[k] = 0;
        if (aggregate)
            {
            const unsigned m = __match_any_sync(0xffffffffu, b);
            const int leader = __ffs(m) - 1;
            uint32_t base = 0;
            if (lane == leader && b != 0xffffffffu)
                base = atomicAdd(&scnt[b], (uint32_t)__popc(m));
            rk[k] = __shfl_sync(0xffffffffu, base, leader) + (uint32_t)__popc(m & cl_lanemask_lt());
            }
        else if (b != 0xffffffffu)
            rk[k] = atomicAdd(&scnt[b], 1u);
        }
    if (oob)
        flag[2] = 1;
    __syncthreads();
    // (3) one global atomic per non-empty bucket reserves the tile's run; exclusive scan of the counts
    const uint32_t per = (nb + NT - 1) / NT;
    const uint32_t b0 = (uint32_t)tid * per;
    uint32_t s = 0;
    for (uint32_t i = 0; i < per; i++)
        {
        const uint32_t b = b0 + i;
        if (b < nb)
            {
            const uint32_t c = scnt[b];
            s += c;
            uint32_t g = 0;
            if (c)
                {
                g = atomicAdd(cursor + (size_t)b * cstride, c);
                if (g + c > capc)
                    {
                    flag[0] = 1;
                    g = 0xffffffffu;
                    }
                }
            sbase[b] = g;
            }
        }
    uint32_t inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
        {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d)
            inc += t;
        }
    if (lane == 31)
        wsum[w] = inc;
    __syncthreads();
    if (w == 0)
        {
        const uint32_t v = lane < NW ? wsum[lane] : 0u;
        uint32_t iv = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
            {
            const uint32_t t = __shfl_up_sync(0xffffffffu, iv, d);
            if (lane >= d)
                iv += t;
            }
        if (lane < NW)
            wsum[lane] = iv - v;
        }
    __syncthreads();
    uint32_t run = wsum[w] + inc - s;
    for (uint32_t i = 0; i < per; i++)
        {
        const uint32_t b = b0 + i;
        if (b < nb)
            {
            sofs[b] = run;
            run += scnt[b];
            }
        }
    __syncthreads();
    // (4) sorted position and destination of every row
    uint32_t placed = 0;
#pragma unroll
    for (int k = 0; k < PER; k++)
        {
        const uint32_t r = (uint32_t)tid + (uint32_t)k * NT;
        if (r < tile_n)
            {
            raw[r] = key[k];
            if (bk[k] != 0xffffffffu)
                {
                const uint32_t g = sbase[bk[k]];
                order[sofs[bk[k]] + rk[k]] = (uint16_t)r;
                sdst[r] = g == 0xffffffffu ? 0xffffffffu : bk[k] * capc + g + rk[k];
                placed++;
                }
            }
        }
    (void)placed;
    if (bulk)
        {
        const bool ok = mbar_wait(bar, 0);
        if (__syncthreads_or(!ok))
            {
            if (tid == 0)
                flag[1] = 3;
            return;
            }
        }
    else
        __syncthreads();
    // (5) records out in bucket order: a group of lanes per record, consecutive records of a run are neighbours
    const uint32_t nsorted = sofs[nb - 1] + scnt[nb - 1]; // rows with a bucket (all of them unless a key was out of range)
    constexpr int U = 4;
    if ((RW & 1u) == 0)
        {
        const uint32_t R2 = RW / 2, G = 32u / R2;
        const uint32_t g = (uint32_t)lane / R2, c2 = (uint32_t)lane - g * R2;
        if (g < G)
            {
            const uint32_t ca = col[2 * c2], cb = col[2 * c2 + 1];
            const uint32_t Wa = ca & 255u, fa = ca >> 8, Wb = cb & 255u, fb = cb >> 8;
            uint2* rec2 = reinterpret_cast<uint2*>(rec);
            const uint32_t step = NW * G;
            for (uint32_t j0 = (uint32_t)w * G + g; j0 < nsorted; j0 += step * U)
                {
                uint32_t dd[U];
                uint2 v[U];
#pragma unroll
                for (int u = 0; u < U; u++)
                    {
                    const uint32_t j = j0 + (uint32_t)u * step;
                    dd[u] = 0xffffffffu;
                    if (j < nsorted)
                        {
                        const uint32_t r = order[j];
                        dd[u] = sdst[r];
                        v[u].x = Wa ? raw[fa + r * Wa] : (uint32_t)(tile0 + r);
                        v[u].y = Wb ? raw[fb + r * Wb] : (uint32_t)(tile0 + r);
                        }
                    }
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (dd[u] != 0xffffffffu)
                        rec2[(uint64_t)dd[u] * R2 + c2] = v[u];
                }
            }
        }
    else
        {
        const uint32_t G = 32u / RW;
        const uint32_t g = (uint32_t)lane / RW, c = (uint32_t)lane - g * RW;
        if (g < G)
            {
            const uint32_t cc = col[c];
            const uint32_t W = cc & 255u, fb = cc >> 8;
            const uint32_t step = NW * G;
            for (uint32_t j0 = (uint32_t)w * G + g; j0 < nsorted; j0 += step * U)
                {
                uint32_t dd[U], v[U];
#pragma unroll
                for (int u = 0; u < U; u++)
                    {
                    const uint32_t j = j0 + (uint32_t)u * step;
                    dd[u] = 0xffffffffu;
                    if (j < nsorted)
                        {
                        const uint32_t r = order[j];
                        dd[u] = sdst[r];
                        v[u] = W ? raw[fb + r * W] : (uint32_t)(tile0 + r);
                        }
                    }
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (dd[u] != 0xffffffffu)
                        rec[(uint64_t)dd[u] * RW + c] = v[u];
                }
            }
        }
    }

// ---- pass 2: clusters of 8 CTAs, one coarse bucket at a time: records -> owner CTA -> slot order -> fields out SoA ----
template <int W>
__device__ __forceinline__ void cl_emit(uint32_t* __restrict__ out, const uint32_t* __restrict__ src,
                                        const uint16_t* __restrict__ row_at, uint32_t total, uint32_t RW, int tid, int nt)
    {
    for (uint32_t q = tid; q < total; q += nt)
        {
        const uint32_t j = q / W, c = q - j * W;
        out[q] = src[(uint32_t)row_at[j] * RW + c];
        }
    }

// Persistent: cluster `cid` of `ncl` takes the coarse buckets cid, cid + ncl, ...  CTA r of the cluster OWNS the slots
// [r * 2^LF, (r + 1) * 2^LF) of the bucket.  Per bucket:
//   A  one TMA bulk copy brings my eighth of the bucket's records (any ids) into shared memory
//   B  rows are counted per owner (8 counters); the 8 x 8 count matrix is exchanged through distributed shared memory
//   C  my records leave sorted by owner as contiguous runs into the owners' inboxes -- a scratch area of this cluster
//      in global memory that is re-used for every bucket and therefore stays in the L2 (ncl * 8 * 2^LF records in all)
//   D  my inbox (the <= 2^LF records of my slots, any order) comes back with 16-byte cp.async
//   E  slot = key & (2^LF - 1) is the row's rank (occupancy bitmap: duplicate -> flag, gaps -> popc prefix); the
//      fields go out SoA, fully coalesced, at the bucket's own output range
// The exchange goes through the L2 and not through st.shared::cluster because scattered 40-byte records cost ~4.1
// cycles each as remote shared-memory stores (tools/dsmem_bench.cu, profiles/r3_dsmem_bench.txt: 9.7 B/cycle/SM;
// the first version of this kernel spent 0.68 ms that way), while runs sorted by owner are plain coalesced traffic.
// Shared memory: [records: CAPF * RW words + 32 B][row_of u16 CAPF][row_at u16 CAPF][bitmap CAPF / 32][wpref CAPF / 32]
template <int CS>
__global__ void __launch_bounds__(1024) k7_cluster_place(int LF, uint32_t nb, uint32_t ncl, const uint32_t* __restrict__ cursor,
                                                        uint32_t cstride, const uint32_t* __restrict__ rec,
                                                        uint32_t* __restrict__ scratch, uint32_t* __restrict__ flag,
                                                        const __grid_constant__ SlotArgs args)
    {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar_mem;
    __shared__ uint32_t red[32];
    __shared__ uint32_t cntm[2][CS][CS]; // [iteration parity][source CTA][owner CTA]: records source -> owner (written by the peers)
    __shared__ uint32_t ocnt[CS], oofs[CS + 1], ioff[CS], itot[CS];
    if (flag[0] != 0 || flag[2] != 0) // written by pass 1 only: uniform over the grid
        return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NT = 1024, NW = NT / 32, PER = 4; // a share is at most 2^12 records
    const uint32_t crank = cl_ctarank();
    const uint32_t cid = blockIdx.x / CS;
    const uint32_t CAPF = 1u << LF, capc = CAPF * CS, RW = args.row_words, nwords = CAPF / 32;
    uint32_t* buf = reinterpret_cast<uint32_t*>(smem_raw);
    uint16_t* row_of = reinterpret_cast<uint16_t*>(smem_raw + (size_t)CAPF * RW * 4 + 32);
    uint16_t* row_at = row_of + CAPF; // step C: position sorted by owner -> record; step E: sorted position -> record
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(row_at + CAPF);
    uint32_t* wpref = bitmap + nwords;
    uint32_t* inbox0 = scratch + (uint64_t)cid * capc * RW; // owner o's inbox: CAPF records from inbox0 + o * CAPF * RW
    const uint32_t bar = smem_u32(&bar_mem);
    if (tid == 0)
        mbar_init(bar, 1);
    __syncthreads();
    uint32_t parity = 0, it = 0;
    for (uint32_t c = cid; c < nb; c += ncl, it++)
        {
        const uint32_t cnt = min(__ldg(cursor + (size_t)c * cstride), capc);
        if (cnt == 0) // uniform over the cluster
            continue;
        // ---- A: my share of the bucket's records
        const uint32_t S = ((cnt + CS - 1) / CS + 3u) & ~3u; // multiples of 4 records keep every share 16-byte aligned
        const uint32_t lo = min(cnt, crank * S), hi = min(cnt, lo + S), m = hi - lo;
        if (tid == 0 && m)
            {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the buffer was last written by cp.async
            const uint32_t bytes = (m * RW * 4u + 15u) & ~15u;
            mbar_expect_tx(bar, bytes);
            bulk_g2s(smem_u32(buf), rec + ((uint64_t)c * capc + lo) * RW, bytes, bar);
            }
        // output rows of the bucket start after the rows of all lower buckets
        uint32_t part = 0;
        for (uint32_t b = tid; b < c; b += NT)
            part += min(__ldg(cursor + (size_t)b * cstride), capc);
        part = __reduce_add_sync(0xffffffffu, part);
        if (lane == 0)
            red[w] = part;
        if (tid < CS)
            ocnt[tid] = 0;
        __syncthreads();
        uint32_t base = 0;
#pragma unroll
        for (int i = 0; i < NW; i++)
            base += red[i];
        if (m)
            {
            if (!mbar_wait(bar, parity))
                flag[1] = 3; // reported by the host; the cluster keeps its barrier sequence
            parity ^= 1u;
            }
        // ---- B: rows per owner, rank of every row among my rows for the same owner
        uint32_t own[PER], rk[PER];
#pragma unroll
        for (int k = 0; k < PER; k++)
            {
            const uint32_t r = (uint32_t)tid + (uint32_t)k * NT;
            own[k] = 0xffffffffu;
            if (r < m)
                own[k] = (buf[r * RW] >> LF) & (uint32_t)(CS - 1);
            const unsigned mm = __match_any_sync(0xffffffffu, own[k]);
            const int leader = __ffs(mm) - 1;
            uint32_t b0 = 0;
            if (lane == leader && own[k] != 0xffffffffu)
                b0 = atomicAdd(&ocnt[own[k]], (uint32_t)__popc(mm));
            rk[k] = __shfl_sync(0xffffffffu, b0, leader) + (uint32_t)__popc(mm & cl_lanemask_lt());
            }
        __syncthreads();
        if (tid < CS * CS) // my counts into every peer's matrix
            {
            const uint32_t peer = (uint32_t)tid / CS, o = (uint32_t)tid % CS;
            cl_st_u32(cl_map(smem_u32(&cntm[it & 1u][crank][o]), peer), ocnt[o]);
            }
        if (tid == 0)
            {
            uint32_t run = 0;
            for (int o = 0; o < CS; o++)
                {
                oofs[o] = run;
                run += ocnt[o];
                }
            oofs[CS] = run;
            }
        cl_sync();
        if (tid < CS)
            {
            uint32_t before = 0, total = 0;
            for (uint32_t sr = 0; sr < (uint32_t)CS; sr++)
                {
                const uint32_t v = cntm[it & 1u][sr][tid];
                total += v;
                if (sr < crank)
                    before += v;
                }
            ioff[tid] = before;
            itot[tid] = total;
            }
#pragma unroll
        for (int k = 0; k < PER; k++)
            if (own[k] != 0xffffffffu)
                row_at[oofs[own[k]] + rk[k]] = (uint16_t)((uint32_t)tid + (uint32_t)k * NT);
        __syncthreads();
        bool over = false;
        uint32_t before_me = 0;
#pragma unroll
        for (int o = 0; o < CS; o++)
            {
            over = over || itot[o] > CAPF;
            if ((uint32_t)o < crank)
                before_me += itot[o];
            }
        if (over) // more records than slots in one CTA's range: duplicate ids.  Same matrix everywhere: the cluster agrees
            {
            if (tid == 0 && crank == 0)
                flag[3] = 1; // not flag[0]: that word decides the early exit above and must not change while clusters start
            continue;
            }
        // ---- C: records out, sorted by owner, to the owners' inboxes
        if ((RW & 1u) == 0)
            {
            const uint32_t R2 = RW / 2, G = 32u / R2;
            const uint32_t g = (uint32_t)lane / R2, c2 = (uint32_t)lane - g * R2;
            if (g < G)
                {
                const uint2* buf2 = reinterpret_cast<const uint2*>(buf);
                uint2* dst2 = reinterpret_cast<uint2*>(inbox0);
                for (uint32_t j = (uint32_t)w * G + g; j < m; j += (uint32_t)NW * G)
                    {
                    const uint32_t r = row_at[j];
                    const uint32_t o = (buf[r * RW] >> LF) & (uint32_t)(CS - 1);
                    dst2[(uint64_t)(o * CAPF + ioff[o] + (j - oofs[o])) * R2 + c2] = buf2[r * R2 + c2];
                    }
                }
            }
        else
            {
            const uint32_t G = 32u / RW;
            const uint32_t g = (uint32_t)lane / RW, cc = (uint32_t)lane - g * RW;
            if (g < G)
                for (uint32_t j = (uint32_t)w * G + g; j < m; j += (uint32_t)NW * G)
                    {
                    const uint32_t r = row_at[j];
                    const uint32_t o = (buf[r * RW] >> LF) & (uint32_t)(CS - 1);
                    inbox0[(uint64_t)(o * CAPF + ioff[o] + (j - oofs[o])) * RW + cc] = buf[r * RW + cc];
                    }
            }
        __threadfence();
        cl_sync(); // every record of the bucket is in its owner's inbox
        // ---- D: my inbox -> shared memory
        const uint32_t tot = itot[crank];
            {
            const uint32_t pieces = (tot * RW * 4u + 15u) >> 4;
            const unsigned char* src = reinterpret_cast<const unsigned char*>(inbox0 + (uint64_t)crank * CAPF * RW);
            const uint32_t sb = smem_u32(buf);
            for (uint32_t q = tid; q < pieces; q += NT)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb + q * 16u), "l"(src + (size_t)q * 16u) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            for (uint32_t i = tid; i < nwords; i += NT)
                bitmap[i] = 0;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
        cl_sync(); // all inboxes are read: the scratch area and the other half of the count matrix may be written again
        if (tot == 0)
            continue;
        // ---- E: slot of every row; a slot taken twice is a duplicate key
        int dup = 0;
        for (uint32_t r = tid; r < tot; r += NT)
            {
            const uint32_t sl = buf[r * RW] & (CAPF - 1u);
            const uint32_t bit = 1u << (sl & 31u);
            if (atomicOr(&bitmap[sl >> 5], bit) & bit)
                dup = 1;
            row_of[sl] = (uint16_t)r;
            }
        if (__syncthreads_or(dup))
            {
            if (tid == 0)
                flag[1] = 1;
            continue;
            }
        const uint16_t* ra = row_of; // every slot taken (dense ids): position == slot
        if (tot != CAPF)
            {
            if (w == 0)
                {
                uint32_t run = 0;
                for (uint32_t i0 = 0; i0 < nwords; i0 += 32)
                    {
                    const uint32_t pc = __popc(bitmap[i0 + lane]); // nwords is a multiple of 32 (LF >= 10)
                    uint32_t inc = pc;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1)
                        {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                        if (lane >= d)
                            inc += t;
                        }
                    wpref[i0 + lane] = run + inc - pc;
                    run += __shfl_sync(0xffffffffu, inc, 31);
                    }
                }
            __syncthreads();
            for (uint32_t sl = tid; sl < CAPF; sl += NT)
                {
                const uint32_t wd = bitmap[sl >> 5];
                if ((wd >> (sl & 31u)) & 1u)
                    row_at[wpref[sl >> 5] + __popc(wd & ((1u << (sl & 31u)) - 1u))] = row_of[sl];
                }
            __syncthreads();
            ra = row_at;
            }
        const uint64_t r0 = (uint64_t)base + before_me;
        for (int fi = 0; fi < args.nfields; fi++)
            {
            const SlotField f = args.f[fi];
            if (f.out == nullptr)
                continue;
            const uint32_t W = f.words;
            uint32_t* out = f.out + r0 * W;
            const uint32_t* src = buf + f.off;
            const uint32_t total = tot * W;
            if (W == 1)
                cl_emit<1>(out, src, ra, total, RW, tid, NT);
            else if (W == 3)
                cl_emit<3>(out, src, ra, total, RW, tid, NT);
            else if (W == 2)
                cl_emit<2>(out, src, ra, total, RW, tid, NT);
            else if (W == 4)
                cl_emit<4>(out, src, ra, total, RW, tid, NT);
            else
                for (uint32_t q = tid; q < total; q += NT)
                    {
                    const uint32_t j = q / W, cc = q - j * W;
                    out[q] = src[(uint32_t)ra[j] * RW + cc];
                    }
            }
        __syncthreads(); // the record buffer is free for the next bucket's bulk copy
        }
    }

// self-test of the bounded mbarrier wait (tests/test_gpu_kernels.py): a barrier that expects bytes nobody sends
__global__ void k7_selftest_mbar_timeout(uint32_t* __restrict__ flag, long long limit)
    {
    __shared__ __align__(8) unsigned long long bar_mem;
    const uint32_t bar = smem_u32(&bar_mem);
    if (threadIdx.x == 0)
        mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0)
        mbar_expect_tx(bar, 16);
    const bool ok = mbar_wait(bar, 0, limit);
    if (__syncthreads_or(!ok))
        {
        if (threadIdx.x == 0)
            flag[1] = 3;
        return;
        }
    if (threadIdx.x == 0)
        flag[1] = 7; // must not happen
    }
    } // namespace

// ---- host side -------------------------------------------------------------------------------------------
static void* g_cl_ws = nullptr;
static size_t g_cl_ws_bytes = 0;
static uint32_t* g_cl_flag_host = nullptr; // pinned

void cluster_release_workspace()
    {
    if (g_cl_ws)
        cudaFree(g_cl_ws);
    g_cl_ws = nullptr;
    g_cl_ws_bytes = 0;
    if (g_cl_flag_host)
        cudaFreeHost(g_cl_flag_host);
    g_cl_flag_host = nullptr;
    }

static inline size_t cl_up256(size_t v) { return (v + 255) / 256 * 256; }

template <int T, int NT>
static cudaError_t cl_launch_scatter(uint64_t n, int LC, uint32_t nb, uint32_t key_const, uint32_t* cursor, uint32_t cstride,
                                     uint32_t* rec, uint32_t* flag, int aggregate, const SlotArgs& a, size_t smem, cudaStream_t st)
    {
    static size_t attr = 0;
    if (smem > attr)
        {
        cudaError_t e = cudaFuncSetAttribute(k7_coarse_scatter<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
        attr = smem;
        }
    const uint32_t tiles = (uint32_t)((n + T - 1) / T);
    k7_coarse_scatter<T, NT><<<tiles, NT, smem, st>>>(n, LC, nb, key_const, cursor, cstride, rec, flag, aggregate, a);
    dev_stats().kernel_launches++;
    return cudaGetLastError();
    }

static size_t cl_scatter_smem(uint32_t in_words, int tile, uint32_t nb)
    {
    return ((size_t)in_words + 1) * tile * 4 + SLOT_MAX_FIELDS * SLOT_SKEW * 4 + (size_t)tile * 2 + 3 * (size_t)nb * 4;
    }

// Tries the cluster path (same contract as dev_reorder_slot).  *handled = 0: geometry does not fit, nothing was
// launched -- the caller continues with the slot path.
int dev_reorder_cluster(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm, int nfields,
                        const ReorderField* fields, int topbit, uint32_t key_const, void* stream_v, int* done,
                        int* out_of_range, void (*mark)(int, cudaStream_t), int* handled)
    {
    *handled = 0;
    // Opt-in (PGSD_B200_CLUSTER=1): measured on B200 at 16 Mi particles this path takes 0.33 + 0.67 ms against
    // 0.50 + 0.21 ms of the slot path (profiles/r3_time_cluster_*.txt) -- see the header.
    const char* en = getenv("PGSD_B200_CLUSTER");
    if (!(en && en[0] == '1'))
        return 0;
    uint64_t min_rows = 1ull << 20;
    if (const char* em = getenv("PGSD_B200_CLUSTER_MIN_ROWS"))
        min_rows = (uint64_t)atoll(em);
    if (n < min_rows || n >= 0xffffffffull || nfields + 2 > SLOT_MAX_FIELDS || keys_sorted == keys)
        return 0;
    cudaStream_t st = (cudaStream_t)stream_v;

    SlotArgs a;
    memset(&a, 0, sizeof(a));
    int nf = 0;
    uint32_t off = 0, in_words = 0;
    bool aligned = ((uintptr_t)keys & 15u) == 0;
    a.f[nf++] = SlotField { keys, keys_sorted, 1u, off };
    off += 1;
    in_words += 1;
    for (int i = 0; i < nfields; i++)
        {
        const ReorderField& f = fields[i];
        if (f.row_bytes == 0 || f.row_bytes % 4 != 0 || (((uintptr_t)f.in | (uintptr_t)f.out) & 3u) != 0 || f.in == nullptr
            || f.out == nullptr)
            return 0;
        if (((uintptr_t)f.in & 15u) != 0)
            aligned = false;
        a.f[nf++] = SlotField { (const uint32_t*)f.in, (uint32_t*)f.out, f.row_bytes / 4, off };
        off += f.row_bytes / 4;
        in_words += f.row_bytes / 4;
        if (off > SLOT_MAX_ROW_WORDS)
            return 0;
        }
    if (perm)
        {
        a.f[nf++] = SlotField { nullptr, perm, 1u, off };
        off += 1;
        }
    if (off > SLOT_MAX_ROW_WORDS)
        return 0;
    a.nfields = nf;
    a.row_words = off;
    a.nranks = 1;
    const char* eb = getenv("PGSD_B200_SLOT_BULK");
    a.bulk = (aligned && !(eb && eb[0] == '0')) ? 1 : 0;

    // slots per CTA of the placement cluster: as many as fit its shared memory
    int LF = 12;
    if (const char* el = getenv("PGSD_B200_CLUSTER_BITS"))
        LF = atoi(el);
    if (LF > 12)
        LF = 12;
    if (LF < 10)
        LF = 10;
    auto place_smem_of = [&](int lf) { return ((size_t)a.row_words * 4 + 4) * ((size_t)1 << lf) + 32 + 2 * (((size_t)1 << lf) / 32) * 4; };
    while (LF > 10 && place_smem_of(LF) > 200 * 1024)
        LF--;
    const size_t place_smem = place_smem_of(LF);
    if (place_smem > 200 * 1024)
        return 0;
    const int LC = LF + 3; // CL_SIZE = 8 CTAs
    static_assert(CL_SIZE == 8, "LC = LF + log2(CL_SIZE)");
    const int bbits = topbit > LC ? topbit - LC : 0;
    if (bbits > 11)
        return 0;
    const uint32_t nb = 1u << bbits;
    static_assert(CL_MAX_BUCKETS == 2048, "bbits <= 11");
    const uint64_t capc = 1ull << LC;
    if (n > (uint64_t)nb * capc)
        return 0; // more keys than slots: duplicates for certain
    if ((uint64_t)nb * capc > 4 * n + (1ull << 20))
        return 0; // sparse ids: the copy would be mostly holes

    int tile = 4096;
    if (const char* et = getenv("PGSD_B200_CLUSTER_TILE"))
        tile = atoi(et);
    if (tile != 1024 && tile != 2048 && tile != 4096)
        tile = 4096;
    while (tile > 1024 && cl_scatter_smem(in_words, tile, nb) > 220 * 1024)
        tile /= 2;
    const size_t scatter_smem = cl_scatter_smem(in_words, tile, nb);
    if (scatter_smem > 220 * 1024)
        return 0;
    int threads = 0;
    if (const char* eth = getenv("PGSD_B200_CLUSTER_THREADS"))
        threads = atoi(eth);
    int aggregate = 1;
    if (const char* ea = getenv("PGSD_B200_CLUSTER_AGG"))
        aggregate = atoi(ea) != 0;

    // how many placement clusters are resident at once (persistent grid; each has its own scratch area)
    static size_t place_attr = 0;
    if (place_smem > place_attr)
        {
        if (cudaFuncSetAttribute(k7_cluster_place<CL_SIZE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)place_smem) != cudaSuccess)
            {
            cudaGetLastError();
            return 0;
            }
        place_attr = place_smem;
        }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(CL_SIZE * 64, 1, 1);
    cfg.blockDim = dim3(1024, 1, 1);
    cfg.dynamicSmemBytes = place_smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL_SIZE;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    static size_t ncl_smem = 0;
    static int ncl_max = 0;
    if (ncl_smem != place_smem)
        {
        int k = 0;
        if (cudaOccupancyMaxActiveClusters(&k, k7_cluster_place<CL_SIZE>, &cfg) != cudaSuccess || k <= 0)
            {
            cudaGetLastError();
            return 0; // clusters of 8 CTAs with this much shared memory cannot be scheduled here
            }
        ncl_max = k;
        ncl_smem = place_smem;
        }
    uint32_t ncl = (uint32_t)ncl_max < nb ? (uint32_t)ncl_max : nb;
    if (const char* ec = getenv("PGSD_B200_CLUSTER_COUNT"))
        {
        const int k = atoi(ec);
        if (k >= 1 && (uint32_t)k < ncl)
            ncl = (uint32_t)k;
        }

    // workspace: [flag 256 B][cursors nb * cstride][records nb * 2^LC * RW words][scratch ncl * 2^LC * RW words]
    const uint32_t cstride = 32;
    const size_t cb = cl_up256((size_t)nb * cstride * 4);
    const size_t copy_bytes = cl_up256((size_t)nb * capc * a.row_words * 4) + 256;
    const size_t scratch_bytes = cl_up256((size_t)ncl_max * capc * a.row_words * 4) + 256;
    const size_t need = 256 + cb + copy_bytes + scratch_bytes;
    if (g_cl_ws_bytes < need)
        {
        if (g_cl_ws)
            cudaFree(g_cl_ws);
        g_cl_ws = nullptr;
        g_cl_ws_bytes = 0;
        if (cudaMalloc(&g_cl_ws, need) != cudaSuccess)
            {
            cudaGetLastError();
            return 0;
            }
        g_cl_ws_bytes = need;
        }
    if (!g_cl_flag_host && cudaHostAlloc((void**)&g_cl_flag_host, 256, cudaHostAllocDefault) != cudaSuccess)
        {
        cudaGetLastError();
        return 0;
        }
    unsigned char* p = (unsigned char*)g_cl_ws;
    uint32_t* flag = (uint32_t*)p;
    uint32_t* cursor = (uint32_t*)(p + 256);
    uint32_t* rec = (uint32_t*)(p + 256 + cb);
    uint32_t* scratch = (uint32_t*)(p + 256 + cb + copy_bytes);

    *handled = 1;
    *done = 0;
    *out_of_range = 0;
    cudaMemsetAsync(p, 0, 256 + cb, st); // flags + cursors
    cudaError_t e;
    if (tile == 4096)
        e = threads == 512 ? cl_launch_scatter<4096, 512>(n, LC, nb, key_const, cursor, cstride, rec, flag, aggregate, a, scatter_smem, st)
                           : cl_launch_scatter<4096, 1024>(n, LC, nb, key_const, cursor, cstride, rec, flag, aggregate, a, scatter_smem, st);
    else if (tile == 2048)
        e = threads == 256 ? cl_launch_scatter<2048, 256>(n, LC, nb, key_const, cursor, cstride, rec, flag, aggregate, a, scatter_smem, st)
                           : cl_launch_scatter<2048, 512>(n, LC, nb, key_const, cursor, cstride, rec, flag, aggregate, a, scatter_smem, st);
    else
        e = cl_launch_scatter<1024, 256>(n, LC, nb, key_const, cursor, cstride, rec, flag, aggregate, a, scatter_smem, st);
    if (mark)
        mark(0, st);
    if (e == cudaSuccess)
        {
        cfg.gridDim = dim3(ncl * CL_SIZE, 1, 1);
        e = cudaLaunchKernelEx(&cfg, k7_cluster_place<CL_SIZE>, LF, nb, ncl, (const uint32_t*)cursor, cstride, (const uint32_t*)rec,
                               scratch, flag, a);
        dev_stats().kernel_launches++;
        }
    if (mark)
        mark(1, st);
    if (e != cudaSuccess)
        {
        set_last_error(std::string("reorder cluster path launch: ") + cudaGetErrorString(e));
        cudaGetLastError();
        return -1;
        }
    cudaMemcpyAsync(g_cl_flag_host, flag, 16, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
        {
        set_last_error(std::string("reorder cluster path: ") + cudaGetErrorString(e));
        return -1;
        }
    if (g_cl_flag_host[1] == 3)
        {
        set_last_error("reorder cluster path: a bulk copy did not complete");
        return -1;
        }
    *done = (g_cl_flag_host[0] == 0 && g_cl_flag_host[1] == 0 && g_cl_flag_host[2] == 0 && g_cl_flag_host[3] == 0) ? 1 : 0;
    *out_of_range = g_cl_flag_host[2] != 0 ? 1 : 0;
    return 0;
    }

// 0: the bounded wait gave up and the kernel reported it (flag 3); anything else is a failure
int dev_selftest_mbar_timeout()
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    uint32_t* d = nullptr;
    if (cudaMalloc((void**)&d, 16) != cudaSuccess)
        return -6;
    cudaMemset(d, 0, 16);
    k7_selftest_mbar_timeout<<<1, 64>>>(d, 2000000ll);
    uint32_t h[4] = { 0, 0, 0, 0 };
    cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess)
        {
        set_last_error(std::string("selftest: ") + cudaGetErrorString(e));
        return -1;
        }
    return h[1] == 3 ? 0 : 1;
    }
} // namespace pgsdb
