// kernels_slot.cu -- K4/K5 "slot path": the reorder of a decoded frame when the particle ids are UNIQUE
// (the normal case: ids are a permutation of 0..N-1; hoomd.py:885-893 reads them
// from log/particles/id).
//
// Same oracle as kernels_sort.cu:  o = numpy.argsort(ids, kind='stable');  out_f = in_f[o].
// With unique keys the stable order is simply the key order, so no ranking is needed at all:
//
//   k6_slot_hist     histogram of the key bits above the low L ("slot") bits: <= 32768 buckets of at most
//                    CAP = 2^L keys each.  A bucket with more than CAP keys proves a duplicate -> flag.
//   k6_slot_scan     exclusive scan of the bucket counts (one CTA) -> bucket bases (output rows).
//   k6_slot_scatter  rows (key + all fields [+ original index]) are staged SoA -> shared memory with one
//                    TMA bulk copy per field (the keys go straight to registers, so the atomics below are
//                    in flight while the payload arrives), every row takes the next free position of its
//                    bucket (one global atomic per row on a cursor that has a 128-byte line of its own),
//                    and leaves as one interleaved record (a group of lanes per record, 64-bit stores).
//   k6_slot_place    one CTA per bucket: the bucket (<= CAP records, 40 KB for the SPH schema) comes into
//                    shared memory, every record is placed at slot = key & (CAP-1) (an occupancy bitmap
//                    detects duplicates and compacts gaps), and the fields are written back SoA, fully
//                    coalesced, at the bucket's own output range.
//
// Layout of the interleaved copy ("lines", default): unit j (256 bytes; PGSD_B200_SLOT_UNIT = log2) of bucket b is
// unit j * nb + b.  The
// cursors of all buckets advance at about the same pace, so the lines being filled form ONE compact window
// that moves through the copy, and the L2 write-backs of completed lines fall into few open DRAM rows.
// With contiguous buckets ("flat", PGSD_B200_SLOT_LAYOUT=flat: one TMA bulk copy per bucket in
// k6_slot_place) the nb write frontiers are spread over the whole copy: scatter 0.61 ms instead of 0.50 ms
// at 16 Mi particles.
//
// Bytes moved: 4 (hist) + 40 + 40 (scatter) + 40 + 40 (place) = 164 B/particle for the 40-byte SPH row,
// against 200 B/particle of the general path (bucket pass + 2 segmented pair passes + gather), and none
// of its ballot ranking.  When a duplicate key is found (flag), nothing of the result is trusted: the
// caller falls back to the stable general path of kernels_sort.cu.
//
// Measured on B200, 16 Mi particles: place 0.21 ms = 0.98 of the measured HBM copy rate; scatter 0.43-0.44 ms with
// 512 threads per 1024-row tile (2 rows per thread, 4 CTAs = 64 warps per SM, 196 KB carve-out; profiles/r4_*: the
// source-level ncu capture showed 41 % of the stall samples on the first use of an atomic's result at 50 % occupancy).
// With 256 threads (profiles/README.md, r2) it was 0.50 ms, which decomposes (round-1 timing experiments, profiles/r2_time_slot_scatter_decomposition.txt) into
// staging 0.12 + cursor atomics 0.16-0.20 + record stores 0.11 (+ 0.17 when they go to their scattered
// places) -- additive although no unit is above 35 % busy in ncu: DRAM spends its time opening rows for the
// write-backs, not transferring.  Tried without gain: a persistent double-buffered variant that overlaps
// TMA, atomics and stores inside the CTA (0.55 ms), L2 evict-first loads (0.57 ms), packed / 32-byte /
// 256-byte cursor strides (+-3 %), tiles of 512 / 2048 rows (+-5 %); and a separate kernel that only takes the
// positions (keys -> atomics -> 4 bytes per row; 0.16 ms = the L2's ~100 G atomics/s) with an atomic-free
// scatter after it: 0.82 ms, because records then no longer arrive in position order and half-filled lines
// are evicted and fetched again (DRAM 1.34 GB read / 0.99 GB written instead of 0.74 / 0.67).  Round 2, also without
// gain: batched atomics, a fifth CTA per SM / 228 KB carve-out, a kernel specialised for
// the 40-byte row with a third of the instructions, staging gaps without bank conflicts, a software-pipelined
// persistent kernel.  tools/atomic_overlap_bench.cu: a coalesced copy of the same bytes with one cursor atomic per
// row takes 0.30 ms -- the floor of this pass.
//
// Distributed (one frame partitioned over the ranks, pgsd_b200_reorder_distributed, bottom of this file): the same
// kernels on each owner's share, fed by k6_part_scatter (records appended to the owner's inbox in peer memory as
// contiguous runs) + k6_slot_scatter_rec (the scatter above on already interleaved records), or -- fused mode --
// by k6_slot_scatter storing straight into the owners' bucketed copies.
//
// sm_100a only (cp.async.bulk + mbarrier).  No CPU fallback.
#include "slot_common.cuh"

#include "comm.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace pgsdb
{
using namespace slotk;
namespace
    {
// flag words (device): [2] set by k6_slot_hist: a key lies outside the range guessed from n
//                      [0] set by k6_slot_scan: [2], or a bucket holds more than CAP keys (duplicate ids)
//                      [1] set by k6_slot_place: 1 = two rows of a bucket share a slot (duplicate ids),
//                          3 = a bulk copy never arrived
// ---- bucket histogram ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k6_slot_hist(const uint32_t* __restrict__ keys, uint64_t n, int L, uint32_t bmask,
                                                    uint32_t nb, uint32_t high_mask, uint32_t key_limit,
                                                    uint32_t* __restrict__ counts, uint32_t* __restrict__ flag)
    {
    extern __shared__ uint32_t hist[];
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x)
        hist[i] = 0;
    __syncthreads();
    const uint64_t n4 = n / 4;
    const uint4* k4 = reinterpret_cast<const uint4*>(keys);
    uint32_t high = 0; // key bits above the assumed range (high_mask != 0: the range was guessed from n, not measured)
    uint32_t kmax = 0; // largest key (key_limit != 0: ids must be < key_limit, distributed reorder)
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x)
        {
        const uint4 v = __ldg(k4 + i);
        high |= v.x | v.y | v.z | v.w;
        kmax = max(max(kmax, max(v.x, v.y)), max(v.z, v.w));
        atomicAdd(&hist[(v.x >> L) & bmask], 1u);
        atomicAdd(&hist[(v.y >> L) & bmask], 1u);
        atomicAdd(&hist[(v.z >> L) & bmask], 1u);
        atomicAdd(&hist[(v.w >> L) & bmask], 1u);
        }
    if (blockIdx.x == 0)
        for (uint64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x)
            {
            const uint32_t k = keys[i];
            high |= k;
            kmax = max(kmax, k);
            atomicAdd(&hist[(k >> L) & bmask], 1u);
            }
    if ((high & high_mask) || (key_limit != 0 && kmax >= key_limit))
        flag[2] = 1;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x)
        {
        const uint32_t c = hist[i];
        if (c)
            atomicAdd(counts + i, c);
        }
    }

// ---- exclusive scan of <= 32768 bucket counts, one CTA ---------------------------------------------------
// Global accesses are coalesced (counts in, bases out, through shared memory); thread t scans the `per`
// consecutive buckets from t * per in between.  cursor == NULL: the cursors were zeroed by a memset
// ("lines" layout: positions are relative to the bucket).
__global__ void __launch_bounds__(1024) k6_slot_scan(const uint32_t* __restrict__ counts, uint32_t nb, uint32_t cap,
                                                    uint32_t n, uint32_t* __restrict__ base, uint32_t* __restrict__ cursor,
                                                    uint32_t cstride, uint32_t* __restrict__ flag)
    {
    extern __shared__ uint32_t sc[]; // nb + nb / 32 (one pad word per 32: the per-thread chunks start 16 or 32 words apart)
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t over;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (uint32_t i = tid; i < nb; i += 1024)
        sc[i + (i >> 5)] = __ldg(counts + i);
    if (tid == 0)
        over = 0;
    __syncthreads();
    const uint32_t per = (nb + 1023u) / 1024u;
    const uint32_t b0 = (uint32_t)tid * per;
    uint32_t s = 0, mx = 0;
    for (uint32_t i = 0; i < per; i++)
        {
        const uint32_t b = b0 + i;
        if (b < nb)
            {
            const uint32_t c = sc[b + (b >> 5)];
            s += c;
            mx = c > mx ? c : mx;
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
        const uint32_t v = wsum[lane];
        uint32_t iv = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
            {
            const uint32_t t = __shfl_up_sync(0xffffffffu, iv, d);
            if (lane >= d)
                iv += t;
            }
        wsum[lane] = iv - v;
        }
    __syncthreads();
    if (mx > cap)
        over = 1;
    uint32_t run = wsum[w] + inc - s;
    for (uint32_t i = 0; i < per; i++)
        {
        const uint32_t b = b0 + i;
        if (b < nb)
            {
            const uint32_t c = sc[b + (b >> 5)];
            sc[b + (b >> 5)] = run;
            run += c;
            }
        }
    __syncthreads();
    for (uint32_t i = tid; i < nb; i += 1024)
        {
        const uint32_t v = sc[i + (i >> 5)];
        base[i] = v;
        if (cursor)
            cursor[(size_t)i * cstride] = v;
        }
    if (tid == 0)
        {
        base[nb] = n;
        if (over || flag[2] != 0)
            flag[0] = 1;
        }
    }

// ---- scatter: every row to the next free position of its bucket, as one interleaved record -------------
// Field tiles sit field-major in shared memory, tile i shifted by 4 * i words so that the columns of one
// record fall into different banks.  The keys (field 0) are loaded straight into registers: their atomics
// are in flight while the TMA engine brings the payload tiles.

template <int T, int NT>
__global__ void __launch_bounds__(NT, (2048 / NT > 8 ? 8 : 2048 / NT)) k6_slot_scatter(uint64_t n, uint32_t tile_first, int L, uint32_t bmask, uint32_t* __restrict__ cursor,
                                                     uint32_t cstride, uint32_t* __restrict__ aos, uint32_t* __restrict__ flag,
                                                     const __grid_constant__ SlotArgs args)
    {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar_mem;
    __shared__ uint32_t col[SLOT_MAX_ROW_WORDS]; // column c of the record: (word offset of its field tile + c') << 8 | W; 0: original index
    uint32_t* raw = reinterpret_cast<uint32_t*>(smem_raw); // field tiles: T * W words each (+ skew)
    if (flag[0] != 0) // written by k6_slot_scan only: uniform
        return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint64_t tile0 = ((uint64_t)blockIdx.x + tile_first) * T;
    const uint32_t tile_n = (uint32_t)((n - tile0) < (uint64_t)T ? (n - tile0) : T);
    const uint32_t RW = args.row_words;
    constexpr int PER = T / NT;

    // keys of my rows: registers
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
    uint32_t* sdst = raw + fbase; // T : row -> record position in the interleaved copy

    // (1) the payload tiles start their way global -> shared
    const bool bulk = args.bulk && tile_n == (uint32_t)T && tx_bytes != 0;
    const uint32_t bar = smem_u32(&bar_mem);
    if (bulk && tid == 0)
        mbar_init(bar, 1);
    __syncthreads();
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

    // (2) one position per row from the bucket's cursor.  (Issuing all of a thread's atomics before the first result is
    // used was measured slower: 0.53 instead of 0.49 ms at 4 rows per thread, no difference at 2 -- profiles/r4_ab_slot.txt.)
    uint32_t d[PER];
#pragma unroll
    for (int k = 0; k < PER; k++)
        {
        const uint32_t r = (uint32_t)tid + (uint32_t)k * NT;
        d[k] = 0;
        if (r < tile_n)
            {
            const uint32_t b = (key[k] >> L) & bmask;
            d[k] = atomicAdd(cursor + (size_t)b * cstride, 1u);
            if (args.nbl)
                {
                // position inside the bucket (< 4096), the bucket (< 32768) and, distributed, its owner
                uint32_t owner = 0, bl = b;
                if (args.nranks > 1)
                    {
                    owner = b / args.nbr;
                    bl = b - owner * args.nbr;
                    }
                d[k] = (d[k] & 4095u) | (bl << 12) | (owner << 27);
                }
            }
        }
#pragma unroll
    for (int k = 0; k < PER; k++)
        {
        const uint32_t r = (uint32_t)tid + (uint32_t)k * NT;
        sdst[r] = d[k];
        raw[r] = key[k];
        }
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

    // (3) records out
    slot_records_out<NT>(raw, sdst, col, tile0, tile_n, RW, aos, args, lane, w);
    }

// ---- distributed reorder, step 1: partition by OWNER ------------------------------------------------------------
// Same staging as k6_slot_scatter, but a row only needs to reach the rank that owns its id range (<= 8
// destinations).  The rows of a tile are counted per owner in shared memory, one atomic per (tile, owner) reserves
// a run in the owner's inbox, and the records leave sorted by owner: every tile writes a few contiguous runs of
// about T / ranks records (5 KB for the SPH row) -- into its own HBM or, through the IPC mapping, over NVLink.
template <int T, int NT>
__global__ void __launch_bounds__(NT) k6_part_scatter(uint64_t n, int L, uint32_t bmask, uint32_t* __restrict__ owner_cursor,
                                                     uint32_t* __restrict__ flag, const __grid_constant__ SlotArgs args)
    {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar_mem;
    __shared__ uint32_t col[SLOT_MAX_ROW_WORDS];
    __shared__ uint32_t scnt[8], sofs[9], sbase[8];
    uint32_t* raw = reinterpret_cast<uint32_t*>(smem_raw);
    if (flag[0] != 0) // set before this kernel starts (kd_plan: the step was called off on every rank): uniform
        return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint64_t tile0 = (uint64_t)blockIdx.x * T;
    const uint32_t tile_n = (uint32_t)((n - tile0) < (uint64_t)T ? (n - tile0) : T);
    const uint32_t RW = args.row_words;
    constexpr int PER = T / NT;

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
    uint32_t* sdst = raw + fbase;                                  // T : row -> owner << 27 | record index in its inbox
    uint16_t* order = reinterpret_cast<uint16_t*>(sdst + T);       // T : position sorted by owner -> row
    if (tid < 8)
        scnt[tid] = 0;
    const bool bulk = args.bulk && tile_n == (uint32_t)T && tx_bytes != 0;
    const uint32_t bar = smem_u32(&bar_mem);
    if (bulk && tid == 0)
        mbar_init(bar, 1);
    __syncthreads();
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
    // rank of every row among the tile's rows for the same owner
    uint32_t own[PER], rk[PER];
#pragma unroll
    for (int k = 0; k < PER; k++)
        {
        const uint32_t r = (uint32_t)tid + (uint32_t)k * NT;
        own[k] = 0;
        rk[k] = 0;
        if (r < tile_n)
            {
            own[k] = ((key[k] >> L) & bmask) / args.nbr;
            rk[k] = atomicAdd(&scnt[own[k]], 1u);
            }
        }
    __syncthreads();
    if (tid < args.nranks)
        sbase[tid] = scnt[tid] ? atomicAdd(owner_cursor + tid, scnt[tid]) : 0u; // one run per (tile, owner)
    if (tid == 0)
        {
        uint32_t run = 0;
        for (int o = 0; o < 8; o++)
            {
            sofs[o] = run;
            run += scnt[o];
            }
        sofs[8] = run;
        }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; k++)
        {
        const uint32_t r = (uint32_t)tid + (uint32_t)k * NT;
        if (r < tile_n)
            {
            order[sofs[own[k]] + rk[k]] = (uint16_t)r;
            sdst[r] = (own[k] << 27) | (sbase[own[k]] + rk[k]);
            raw[r] = key[k];
            }
        }
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
    slot_records_out<NT>(raw, sdst, col, tile0, tile_n, RW, nullptr, args, lane, w, order);
    }

// ---- distributed reorder, step 2: the slot scatter on records that are already interleaved (the inbox) -----------
// One bulk copy per tile; bucket = key bits above the slot bits minus this rank's first bucket; positions from the
// local cursors; records go into the "lines" copy k6_slot_place reads.
template <int T, int NT>
__global__ void __launch_bounds__(NT) k6_slot_scatter_rec(const uint32_t* __restrict__ rec, uint64_t n_host,
                                                         const uint32_t* __restrict__ n_dev, int L, uint32_t bmask,
                                                         uint32_t bucket0, uint32_t* __restrict__ cursor, uint32_t cstride,
                                                         uint32_t* __restrict__ aos, uint32_t* __restrict__ flag,
                                                         const __grid_constant__ SlotArgs args)
    {
    // n_dev != NULL: the row count was decided on the device (kd_plan); the grid covers the largest possible inbox
    const uint64_t n = n_dev ? (uint64_t)*n_dev : n_host;
    if ((uint64_t)blockIdx.x * T >= n)
        return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar_mem;
    __shared__ uint32_t col[SLOT_MAX_ROW_WORDS];
    uint32_t* rows = reinterpret_cast<uint32_t*>(smem_raw); // T records of RW words
    if (flag[0] != 0)
        return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint64_t tile0 = (uint64_t)blockIdx.x * T;
    const uint32_t tile_n = (uint32_t)((n - tile0) < (uint64_t)T ? (n - tile0) : T);
    const uint32_t RW = args.row_words;
    uint32_t* sdst = rows + (size_t)T * RW;
    if (tid < (int)RW)
        col[tid] = ((uint32_t)tid << 8) | RW; // column c of record r: rows[c + r * RW]
    const uint32_t bar = smem_u32(&bar_mem);
    const uint32_t bytes = tile_n * RW * 4u;
    const bool bulk = (bytes & 15u) == 0; // tile0 * RW * 4 is a multiple of 16 (T is a multiple of 4)
    if (bulk && tid == 0)
        mbar_init(bar, 1);
    __syncthreads();
    if (bulk)
        {
        if (tid == 0)
            {
            mbar_expect_tx(bar, bytes);
            bulk_g2s(smem_u32(rows), rec + tile0 * RW, bytes, bar);
            }
        const bool ok = mbar_wait(bar, 0);
        if (__syncthreads_or(!ok))
            {
            if (tid == 0)
                flag[1] = 3;
            return;
            }
        }
    else
        {
        for (uint32_t q = tid; q < tile_n * RW; q += NT)
            rows[q] = __ldg(rec + tile0 * RW + q);
        __syncthreads();
        }
    constexpr int PER = T / NT;
    uint32_t d[PER];
#pragma unroll
    for (int k = 0; k < PER; k++)
        {
        const uint32_t r = (uint32_t)tid + (uint32_t)k * NT;
        d[k] = 0;
        if (r < tile_n)
            {
            const uint32_t bl = ((rows[r * RW] >> L) & bmask) - bucket0;
            d[k] = (atomicAdd(cursor + (size_t)bl * cstride, 1u) & 4095u) | (bl << 12);
            }
        }
#pragma unroll
    for (int k = 0; k < PER; k++)
        sdst[(uint32_t)tid + (uint32_t)k * NT] = d[k];
    __syncthreads();
    slot_records_out<NT>(rows, sdst, col, tile0, tile_n, RW, aos, args, lane, w);
    }

// ---- place: one CTA per bucket; slot = key & (CAP-1) is the row's rank among the bucket's (unique) keys ---
template <int W>
__device__ __forceinline__ void slot_emit(uint32_t* __restrict__ out, const uint32_t* __restrict__ src,
                                          const uint16_t* __restrict__ row_at, uint32_t total, uint32_t RW, int tid, int nt)
    {
    for (uint32_t q = tid; q < total; q += nt)
        {
        const uint32_t j = q / W, c = q - j * W;
        out[q] = src[(uint32_t)row_at[j] * RW + c];
        }
    }

__global__ void __launch_bounds__(1024) k6_slot_place(const uint32_t* __restrict__ base, int L, const uint32_t* __restrict__ aos,
                                                     uint32_t* __restrict__ flag, const __grid_constant__ SlotArgs args)
    {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar_mem;
    if (flag[0] != 0) // written by k6_slot_scan only: uniform
        return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nt = blockDim.x;
    const uint32_t r0 = base[blockIdx.x];
    const uint32_t cnt = base[blockIdx.x + 1] - r0;
    if (cnt == 0)
        return;
    const uint32_t CAP = 1u << L, RW = args.row_words, nwords = CAP / 32;
    // shared memory: [records: CAP*RW words + 32 B][row_of u16 CAP][row_at u16 CAP][bitmap nwords][wpref nwords]
    const uint64_t byte0 = args.nbl ? 0ull : (uint64_t)r0 * RW * 4u;
    const uint32_t lead = (uint32_t)(byte0 & 15u);
    const uint32_t* rows = reinterpret_cast<const uint32_t*>(smem_raw + lead);
    unsigned char* p = smem_raw + (size_t)CAP * RW * 4 + 128;
    uint16_t* row_of = reinterpret_cast<uint16_t*>(p);
    uint16_t* row_at2 = row_of + CAP;
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(row_at2 + CAP);
    uint32_t* wpref = bitmap + nwords;

    // (1) the whole bucket with one bulk copy (source start rounded down, size rounded up to 16 bytes)
    const uint32_t bar = smem_u32(&bar_mem);
    if (args.nbl)
        {
        // "lines" layout: the bucket's 128-byte lines are nb lines apart; 16-byte cp.async pieces
        const uint32_t pieces = (cnt * RW * 4u + 15u) >> 4;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(aos);
        const uint32_t sbase = smem_u32(smem_raw);
        const uint32_t us = (uint32_t)args.ushift - 4u; // 16-byte pieces per unit, log2
        for (uint32_t q = tid; q < pieces; q += nt)
            {
            const uint64_t at = ((uint64_t)((q >> us) * args.nbl + blockIdx.x) << args.ushift) + (q & ((1u << us) - 1u)) * 16u;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + q * 16u), "l"(src + at) : "memory");
            }
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (uint32_t i = tid; i < nwords; i += nt)
            bitmap[i] = 0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        }
    else if (args.bulk)
        {
        if (tid == 0)
            mbar_init(bar, 1);
        for (uint32_t i = tid; i < nwords; i += nt)
            bitmap[i] = 0;
        __syncthreads();
        if (tid == 0)
            {
            const uint32_t bytes = (lead + cnt * RW * 4u + 15u) & ~15u;
            mbar_expect_tx(bar, bytes);
            bulk_g2s(smem_u32(smem_raw), reinterpret_cast<const unsigned char*>(aos) + (byte0 - lead), bytes, bar);
            }
        const bool ok = mbar_wait(bar, 0);
        if (__syncthreads_or(!ok))
            {
            if (tid == 0)
                flag[1] = 3;
            return;
            }
        }
    else
        {
        for (uint32_t i = tid; i < nwords; i += nt)
            bitmap[i] = 0;
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem_raw + lead);
        const uint32_t* src = aos + (uint64_t)r0 * RW;
        for (uint32_t q = tid; q < cnt * RW; q += nt)
            dst[q] = __ldg(src + q);
        __syncthreads();
        }

    // (2) slot of every row; a slot taken twice is a duplicate key
    int dup = 0;
    for (uint32_t r = tid; r < cnt; r += nt)
        {
        const uint32_t s = rows[r * RW] & (CAP - 1u);
        const uint32_t bit = 1u << (s & 31u);
        if (atomicOr(&bitmap[s >> 5], bit) & bit)
            dup = 1;
        row_of[s] = (uint16_t)r;
        }
    if (__syncthreads_or(dup))
        {
        if (tid == 0)
            flag[1] = 1;
        return;
        }

    // (3) sorted position -> row.  A full bucket has every slot taken: position == slot.
    const uint16_t* row_at = row_of;
    if (cnt != CAP)
        {
        if (w == 0)
            {
            uint32_t run = 0;
            for (uint32_t i0 = 0; i0 < nwords; i0 += 32)
                {
                const uint32_t pc = __popc(bitmap[i0 + lane]); // nwords is a multiple of 32
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
        for (uint32_t s = tid; s < CAP; s += nt)
            {
            const uint32_t wd = bitmap[s >> 5];
            if ((wd >> (s & 31u)) & 1u)
                row_at2[wpref[s >> 5] + __popc(wd & ((1u << (s & 31u)) - 1u))] = row_of[s];
            }
        __syncthreads();
        row_at = row_at2;
        }

    // (4) fields out, SoA, coalesced: the bucket's output rows are [r0, r0 + cnt) of every field
    for (int fi = 0; fi < args.nfields; fi++)
        {
        const SlotField f = args.f[fi];
        if (f.out == nullptr)
            continue;
        const uint32_t W = f.words;
        uint32_t* out = f.out + (uint64_t)r0 * W;
        const uint32_t* src = rows + f.off;
        const uint32_t total = cnt * W;
        if (W == 1)
            slot_emit<1>(out, src, row_at, total, RW, tid, nt);
        else if (W == 3)
            slot_emit<3>(out, src, row_at, total, RW, tid, nt);
        else if (W == 2)
            slot_emit<2>(out, src, row_at, total, RW, tid, nt);
        else if (W == 4)
            slot_emit<4>(out, src, row_at, total, RW, tid, nt);
        else
            for (uint32_t q = tid; q < total; q += nt)
                {
                const uint32_t j = q / W, c = q - j * W;
                out[q] = src[(uint32_t)row_at[j] * RW + c];
                }
        }
    }
// ---- distributed reorder without the host in the loop ------------------------------------------------------------
// Every rank's IPC-exported buffer starts with a control block and two parity copies of a count table; the ranks
// exchange their per-bucket counts, "my records have arrived" and "my buckets are placed" by plain stores into the
// peers' buffers (NVLink) followed by a system-scope release of a flag word that the peer's next kernel spins on.
// Parity = call number & 1: a rank can be at most one call ahead of a peer it has not heard from, so two copies
// suffice.  Flags carry the call number (epoch); a wait gives up after ~2 s worth of cycles and reports it.
struct DistCtl
    {
    uint32_t ready[2][8]; // [parity][source rank]: counts published
    uint32_t done[2][8];  // records of the source rank have reached my inbox
    uint32_t fin[2][8];   // the source rank has placed its buckets ...
    uint32_t dup[2][8];   // ... and found (1) / did not find (0) two records with one id
    };
constexpr size_t DIST_CTL_BYTES = 4096;
constexpr uint32_t DIST_ROW_WORDS = (1u << SLOT_MAX_BUCKET_BITS) + 8; // counts + {n_local lo, hi, out_capacity lo, hi, status}
constexpr size_t DIST_HDR_BYTES = DIST_CTL_BYTES + (size_t)2 * 8 * DIST_ROW_WORDS * 4; // in front of the inbox
static_assert(sizeof(DistCtl) <= DIST_CTL_BYTES && DIST_HDR_BYTES % 256 == 0, "header layout");

struct DistPeers
    {
    void* shared[8]; // base of every rank's buffer (own memory or IPC mapping)
    };
__device__ __forceinline__ uint32_t* dist_row(void* shared, int par, int src)
    {
    return reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(shared) + DIST_CTL_BYTES) + ((size_t)par * 8 + src) * DIST_ROW_WORDS;
    }
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
    {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
    {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
    }
__device__ __forceinline__ bool dist_spin(const uint32_t* p, uint32_t epoch)
    {
    const long long t0 = clock64();
    while (ld_acquire_sys(p) != epoch)
        {
        if (clock64() - t0 > 4000000000ll)
            return false;
        __nanosleep(200);
        }
    return true;
    }

// status bits of a distributed call (device words and the pinned result the host reads at the end)
enum : uint32_t
    {
    DST_RANGE = 1,    // an id outside [0, buckets * 2^L)
    DST_OVERFLOW = 2, // more ids than slots in a bucket: duplicates
    DST_RESOURCE = 4, // a rank could not allocate / map memory
    DST_FITS = 8,     // out_capacity too small on some rank
    DST_TIMEOUT = 16, // a peer never signalled
    DST_DUP = 32      // two records of a bucket share a slot
    };

// block p: my counts + meta into rank p's table, then the flag
__global__ void __launch_bounds__(256) kd_publish(DistPeers peers, int me, int par, uint32_t epoch, const uint32_t* __restrict__ counts,
                                                 uint32_t nbp, const uint32_t* __restrict__ flag, uint64_t n_local,
                                                 uint64_t out_capacity, uint32_t status0)
    {
    void* dst = peers.shared[blockIdx.x];
    if (dst == nullptr) // this peer could not be mapped: it will report a timeout
        return;
    uint32_t* row = dist_row(dst, par, me);
    for (uint32_t i = threadIdx.x; i < nbp; i += blockDim.x)
        row[i] = counts[i];
    if (threadIdx.x == 0)
        {
        row[nbp + 0] = (uint32_t)n_local;
        row[nbp + 1] = (uint32_t)(n_local >> 32);
        row[nbp + 2] = (uint32_t)out_capacity;
        row[nbp + 3] = (uint32_t)(out_capacity >> 32);
        row[nbp + 4] = status0 | (flag[2] != 0 ? DST_RANGE : 0u);
        }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
        st_release_sys(&reinterpret_cast<DistCtl*>(dst)->ready[par][me], epoch);
    }

// One CTA: waits for the counts of all ranks, then decides everything the host used to decide -- identically on
// every rank: who owns how many rows, whether a bucket overflows, where my run starts in every owner's inbox,
// the output rows of my buckets.
__global__ void __launch_bounds__(1024) kd_plan(void* myshared, int G, int me, int par, uint32_t epoch, uint32_t nbp, uint32_t nb_used,
                                               uint32_t nbr, uint32_t cap, uint32_t* __restrict__ owner_cursor,
                                               uint32_t* __restrict__ base, uint32_t* __restrict__ n_owned, uint32_t* __restrict__ flag,
                                               uint32_t* __restrict__ host_out)
    {
    extern __shared__ uint32_t tot[]; // nbr + 1: rows of my buckets
    __shared__ uint32_t s_status, s_all[8], s_low[8], wsum[32], s_red[2][32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    DistCtl* ctl = reinterpret_cast<DistCtl*>(myshared);
    if (tid == 0)
        s_status = 0;
    __syncthreads();
    if (tid < G && !dist_spin(&ctl->ready[par][tid], epoch))
        atomicOr(&s_status, (uint32_t)DST_TIMEOUT);
    __syncthreads();
    const uint32_t* rows[8];
    for (int p = 0; p < 8; p++)
        rows[p] = dist_row(myshared, par, p < G ? p : 0);
    if (tid < G)
        atomicOr(&s_status, rows[tid][nbp + 4]);
    const uint32_t my_b0 = min((uint32_t)me * nbr, nb_used), my_b1 = min(my_b0 + nbr, nb_used);
    uint32_t bad = 0;
    for (int o = 0; o < G; o++)
        {
        const uint32_t b0 = min((uint32_t)o * nbr, nb_used), b1 = min(b0 + nbr, nb_used);
        uint32_t all = 0, low = 0;
        for (uint32_t b = b0 + tid; b < b1; b += 1024)
            {
            uint32_t t = 0;
            for (int p = 0; p < G; p++)
                {
                const uint32_t c = rows[p][b];
                t += c;
                if (p < me)
                    low += c;
                }
            all += t;
            if (t > cap)
                bad |= DST_OVERFLOW;
            if (o == me)
                tot[b - my_b0] = t;
            }
        all = __reduce_add_sync(0xffffffffu, all);
        low = __reduce_add_sync(0xffffffffu, low);
        if (lane == 0)
            {
            s_red[0][w] = all;
            s_red[1][w] = low;
            }
        __syncthreads();
        if (tid == 0)
            {
            uint32_t a2 = 0, l2 = 0;
            for (int i = 0; i < 32; i++)
                {
                a2 += s_red[0][i];
                l2 += s_red[1][i];
                }
            s_all[o] = a2;
            s_low[o] = l2;
            }
        __syncthreads();
        }
    for (uint32_t b = nb_used + tid; b < nbp; b += 1024) // ids beyond the last used bucket (the histogram flags them too)
        for (int p = 0; p < G; p++)
            if (rows[p][b])
                bad |= DST_RANGE;
    if (bad)
        atomicOr(&s_status, bad);
    if (tid < G)
        {
        const unsigned long long capo = (unsigned long long)rows[tid][nbp + 2] | ((unsigned long long)rows[tid][nbp + 3] << 32);
        if ((unsigned long long)s_all[tid] > capo)
            atomicOr(&s_status, (uint32_t)DST_FITS);
        }
    // exclusive prefix of my buckets' rows -> base[0 .. nbr + 1] (entries past my last bucket repeat the total)
    const uint32_t nmine = my_b1 - my_b0;
    const uint32_t per = (nbr + 2 + 1023u) / 1024u;
    const uint32_t i0 = (uint32_t)tid * per;
    uint32_t sum = 0;
    for (uint32_t i = 0; i < per; i++)
        if (i0 + i < nmine)
            sum += tot[i0 + i];
    uint32_t inc = sum;
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
        const uint32_t v = wsum[lane];
        uint32_t iv = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
            {
            const uint32_t t = __shfl_up_sync(0xffffffffu, iv, d);
            if (lane >= d)
                iv += t;
            }
        wsum[lane] = iv - v;
        }
    __syncthreads();
    uint32_t run = wsum[w] + inc - sum;
    for (uint32_t i = 0; i < per; i++)
        {
        const uint32_t k = i0 + i;
        if (k <= nbr + 1)
            {
            base[k] = run;
            if (k < nmine)
                run += tot[k];
            }
        }
    if (tid < 8)
        owner_cursor[tid] = tid < G ? s_low[tid] : 0u;
    __syncthreads();
    if (tid == 0)
        {
        const uint32_t st = s_status;
        *n_owned = st ? 0u : s_all[me];
        flag[0] = st ? 1u : 0u; // the kernels that follow do nothing then
        host_out[0] = st;
        host_out[1] = s_all[me];
        __threadfence_system();
        }
    }

// lanes p < G: tell rank p that my part of the step is complete (which = 0: records delivered, 1: buckets placed)
__global__ void kd_signal(DistPeers peers, int G, int me, int par, uint32_t epoch, int which, const uint32_t* __restrict__ flag)
    {
    const int p = threadIdx.x;
    if (p >= G || peers.shared[p] == nullptr)
        return;
    DistCtl* ctl = reinterpret_cast<DistCtl*>(peers.shared[p]);
    __threadfence_system();
    if (which == 0)
        st_release_sys(&ctl->done[par][me], epoch);
    else
        {
        ctl->dup[par][me] = flag[1];
        __threadfence_system();
        st_release_sys(&ctl->fin[par][me], epoch);
        }
    }

// lanes p < G wait for rank p's signal; which = 1 also folds the peers' duplicate flags into the result the host reads
__global__ void kd_wait(void* myshared, int G, int par, uint32_t epoch, int which, uint32_t* __restrict__ flag, uint32_t* __restrict__ host_out)
    {
    const int p = threadIdx.x;
    DistCtl* ctl = reinterpret_cast<DistCtl*>(myshared);
    uint32_t st = 0;
    if (p < G)
        {
        if (!dist_spin(which == 0 ? &ctl->done[par][p] : &ctl->fin[par][p], epoch))
            st |= DST_TIMEOUT;
        else if (which == 1)
            {
            const uint32_t d = ctl->dup[par][p];
            if (d == 3)
                st |= DST_RESOURCE;
            else if (d != 0)
                st |= DST_DUP;
            }
        }
    st = __reduce_or_sync(0xffffffffu, st);
    if (p == 0)
        {
        if (st & DST_TIMEOUT)
            flag[0] = 1;
        host_out[2 + which] = st;
        __threadfence_system();
        }
    }

// cursors of the distributed reorder: cursor[b * cstride] = first position of this rank's records in bucket b
__global__ void k6_slot_spread(const uint32_t* __restrict__ start, uint32_t nb, uint32_t* __restrict__ cursor, uint32_t cstride)
    {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb)
        cursor[(size_t)b * cstride] = start[b];
    }
    } // namespace

// ---- host side -------------------------------------------------------------------------------------------
static void* g_slot_ws = nullptr;
static size_t g_slot_ws_bytes = 0;
static uint32_t* g_slot_flag_host = nullptr; // pinned, 2 words

void dist_release_workspace();
void slot_release_workspace()
    {
    dist_release_workspace();
    cluster_release_workspace();
    if (g_slot_ws)
        cudaFree(g_slot_ws);
    g_slot_ws = nullptr;
    g_slot_ws_bytes = 0;
    if (g_slot_flag_host)
        cudaFreeHost(g_slot_flag_host);
    g_slot_flag_host = nullptr;
    }

static inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

static size_t slot_place_smem(int L, uint32_t rw)
    {
    const size_t cap = (size_t)1 << L;
    return cap * rw * 4 + 128 + 2 * cap * 2 + 2 * (cap / 32) * 4;
    }

template <int T, int NT>
static cudaError_t launch_scatter(uint64_t n, uint32_t tile_first, uint32_t tiles, int L, uint32_t bmask, uint32_t* cursor,
                                  uint32_t cstride, uint32_t* aos, uint32_t* flag, const SlotArgs& a, uint32_t in_words,
                                  cudaStream_t st)
    {
    // Shared-memory carve-out in KB (PGSD_B200_SLOT_CARVEOUT, read once; default 196).  With 4 CTAs of 46.5 KB the driver
    // picks 200 KB on its own, but as soon as a fifth CTA would fit it takes 228 KB, and the L1 that is left is too
    // small for the stores and atomics in flight: scatter 0.56 instead of 0.45 ms (profiles/r4_ab_slot2.txt).
    static int carve = -1;
    if (carve < 0)
        {
        const char* e = getenv("PGSD_B200_SLOT_CARVEOUT");
        carve = e ? atoi(e) : 196;
        }
    // field tiles (+ skew behind every one) + the row -> position table
    const size_t smem = ((size_t)in_words + 1) * T * 4 + (size_t)(a.nfields + 1) * SLOT_SKEW * 4;
    static size_t attr = 0;
    if (smem > attr)
        {
        cudaError_t e = cudaFuncSetAttribute(k6_slot_scatter<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
        if (carve > 0)
            cudaFuncSetAttribute(k6_slot_scatter<T, NT>, cudaFuncAttributePreferredSharedMemoryCarveout, carve * 100 / 228);
        attr = smem;
        }
    if (tiles == 0)
        return cudaSuccess;
    k6_slot_scatter<T, NT><<<tiles, NT, smem, st>>>(n, tile_first, L, bmask, cursor, cstride, aos, flag, a);
    dev_stats().kernel_launches++;
    return cudaGetLastError();
    }

// Tries the slot path.  topbit: the keys differ only in bits [0, topbit) -- measured by the census, or
// (guessed != 0) assumed from n for dense ids and verified by k6_slot_hist: *out_of_range = 1 reports a miss.
// *done = 1: outputs are complete (stream-ordered work finished; the flag was read
// back).  *done = 0: not applicable or duplicate keys found -- the caller must run the general path (inputs
// are untouched).  phase: optional cudaEvent_t[2] recorded after the scatter and after the placement.
int dev_reorder_slot(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm, int nfields,
                     const ReorderField* fields, int topbit, int guessed, uint32_t key_const, void* stream_v, int* done,
                     int* out_of_range, void (*mark)(int, cudaStream_t))
    {
    *done = 0;
    *out_of_range = 0;
    const char* en = getenv("PGSD_B200_SLOT");
    if (en && en[0] == '0')
        return 0;
    // large frames: coarse partition + cluster placement (kernels_cluster.cu), two passes without per-row atomics
        {
        int handled = 0;
        const int rc = dev_reorder_cluster(n, keys, keys_sorted, perm, nfields, fields, topbit, guessed ? 0u : key_const, stream_v, done,
                                           out_of_range, mark, &handled);
        if (rc != 0 || handled)
            return rc;
        }
    if (n == 0 || n >= 0xffffffffull || nfields + 2 > SLOT_MAX_FIELDS || keys_sorted == keys)
        return 0;
    cudaStream_t st = (cudaStream_t)stream_v;

    SlotArgs a;
    memset(&a, 0, sizeof(a));
    int nf = 0;
    uint32_t off = 0, in_words = 0;
    bool aligned = ((uintptr_t)keys & 15u) == 0;
    if (!aligned)
        return 0; // the histogram reads the keys as uint4
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
    SlotArgs a_place = a; // the interleaved copy is always 16-byte aligned
    a_place.bulk = !(eb && eb[0] == '0') ? 1 : 0;

    // slot bits: as few as keep the bucket count <= 32768
    int L = SLOT_MIN_BITS;
    const char* el = getenv("PGSD_B200_SLOT_BITS");
    if (el)
        L = atoi(el);
    if (L < SLOT_MIN_BITS)
        L = SLOT_MIN_BITS;
    while (topbit - L > SLOT_MAX_BUCKET_BITS)
        L++;
    if (L > SLOT_MAX_BITS)
        return 0;
    const size_t place_smem = slot_place_smem(L, a.row_words);
    if (place_smem > 220 * 1024)
        return 0;
    const int bbits = topbit > L ? topbit - L : 0;
    const uint32_t nb = 1u << bbits;
    const uint32_t bmask = nb - 1u;
    const uint32_t cap = 1u << L;
    if (n > (uint64_t)nb * cap)
        return 0; // more keys than slots: duplicates for certain

    int tile = 1024;
    const char* et = getenv("PGSD_B200_SLOT_TILE");
    if (et)
        tile = atoi(et);
    if (tile != 512 && tile != 1024 && tile != 2048)
        tile = 1024;
    while (tile > 512 && ((size_t)in_words + 1) * tile * 4 > 200 * 1024)
        tile /= 2;
    // rows per thread of the scatter (tile / threads): fewer rows = more warps per SM for the same shared memory
    int per = 2; // 512 threads for a 1024-row tile: 4 CTAs = 64 warps per SM (0.445 ms; 4 rows per thread: 0.49 ms)
    if (const char* ep = getenv("PGSD_B200_SLOT_PER"))
        per = atoi(ep);
    if (per != 1 && per != 2 && per != 4)
        per = 2;
    if (((size_t)in_words + 1) * tile * 4 > 200 * 1024)
        return 0;

    // The bucket cursors are spread out (one per 128-byte line by default): the L2 atomic unit serialises
    // operations on the same sector, and 16384 packed cursors are only 2048 sectors.
    uint32_t cstride = 32;
    const char* ec = getenv("PGSD_B200_SLOT_CSTRIDE");
    if (ec)
        cstride = (uint32_t)atoi(ec);
    if (cstride != 1 && cstride != 8 && cstride != 16 && cstride != 32 && cstride != 64)
        cstride = 32;
    // workspace: [flag 256 B][counts nb][base nb+1][cursors nb * cstride][interleaved copy + 256 B]
    const char* ely = getenv("PGSD_B200_SLOT_LAYOUT");
    const bool lines = !(ely && !strcmp(ely, "flat"));
    a.nbl = lines ? nb : 0u;
    a_place.nbl = a.nbl;
    const size_t tb = up256((size_t)(nb + 1) * 4);
    const size_t cb = up256((size_t)nb * cstride * 4);
    int ushift = 8; // 256-byte units: 2 % faster than 128-byte lines, 1 KB and more is slower (profiles/README.md)
    const char* eu = getenv("PGSD_B200_SLOT_UNIT");
    if (eu)
        ushift = atoi(eu);
    if (ushift < 7 || ushift > 12)
        ushift = 8;
    a.ushift = ushift;
    a_place.ushift = ushift;
    const size_t unit = (size_t)1 << ushift;
    const size_t copy_bytes = lines ? (size_t)nb * (((size_t)cap * a.row_words * 4 + unit - 1) / unit) * unit : (size_t)n * a.row_words * 4;
    const size_t need = 256 + 2 * tb + cb + up256(copy_bytes) + 256;
    if (g_slot_ws_bytes < need)
        {
        if (g_slot_ws)
            cudaFree(g_slot_ws);
        g_slot_ws = nullptr;
        g_slot_ws_bytes = 0;
        if (cudaMalloc(&g_slot_ws, need) != cudaSuccess)
            {
            cudaGetLastError();
            return 0; // the general path reports its own allocation failures
            }
        g_slot_ws_bytes = need;
        }
    if (!g_slot_flag_host && cudaHostAlloc((void**)&g_slot_flag_host, 256, cudaHostAllocDefault) != cudaSuccess)
        {
        cudaGetLastError();
        return 0;
        }
    unsigned char* p = (unsigned char*)g_slot_ws;
    uint32_t* flag = (uint32_t*)p;
    uint32_t* counts = (uint32_t*)(p + 256);
    uint32_t* base = (uint32_t*)(p + 256 + tb);
    uint32_t* cursor = (uint32_t*)(p + 256 + 2 * tb);
    uint32_t* aos = (uint32_t*)(p + 256 + 2 * tb + cb);

    static bool attr_done = false;
    if (!attr_done)
        {
        cudaFuncSetAttribute(k6_slot_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << SLOT_MAX_BUCKET_BITS) * 4);
        cudaFuncSetAttribute(k6_slot_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, ((1 << SLOT_MAX_BUCKET_BITS) / 32 * 33 + 1) * 4);
        cudaFuncSetAttribute(k6_slot_place, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr_done = true;
        }
    cudaMemsetAsync(p, 0, 256 + tb, st); // flag + counts
    int hgrid = dev_sm_count() * (nb <= 16384 ? 2 : 1);
    const uint64_t want = (n / 4 + 1023) / 1024;
    if ((uint64_t)hgrid > want)
        hgrid = want ? (int)want : 1;
    const uint32_t high_mask = (guessed && topbit < 32) ? ~((1u << topbit) - 1u) : 0u;
    k6_slot_hist<<<hgrid, 1024, (size_t)nb * 4, st>>>(keys, n, L, bmask, nb, high_mask, 0u, counts, flag);
    if (lines)
        cudaMemsetAsync(cursor, 0, (size_t)nb * cstride * 4, st);
    k6_slot_scan<<<1, 1024, (size_t)(nb + nb / 32 + 1) * 4, st>>>(counts, nb, cap, (uint32_t)n, base, lines ? nullptr : cursor, cstride, flag);
    dev_stats().kernel_launches += 2;
    cudaError_t e = cudaGetLastError();
    const uint32_t tiles_all = (uint32_t)((n + tile - 1) / tile), tile_first = 0;
    if (e == cudaSuccess && tiles_all > 0)
        {
        if (tile == 512)
            e = per == 2 ? launch_scatter<512, 256>(n, tile_first, tiles_all, L, bmask, cursor, cstride, aos, flag, a, in_words, st)
                         : launch_scatter<512, 128>(n, tile_first, tiles_all, L, bmask, cursor, cstride, aos, flag, a, in_words, st);
        else if (tile == 2048)
            e = per == 2 ? launch_scatter<2048, 1024>(n, tile_first, tiles_all, L, bmask, cursor, cstride, aos, flag, a, in_words, st)
                         : launch_scatter<2048, 512>(n, tile_first, tiles_all, L, bmask, cursor, cstride, aos, flag, a, in_words, st);
        else if (per == 1)
            e = launch_scatter<1024, 1024>(n, tile_first, tiles_all, L, bmask, cursor, cstride, aos, flag, a, in_words, st);
        else if (per == 2)
            e = launch_scatter<1024, 512>(n, tile_first, tiles_all, L, bmask, cursor, cstride, aos, flag, a, in_words, st);
        else
            e = launch_scatter<1024, 256>(n, tile_first, tiles_all, L, bmask, cursor, cstride, aos, flag, a, in_words, st);
        }
    if (mark)
        mark(0, st);
    if (e == cudaSuccess)
        {
        k6_slot_place<<<nb, cap / 4, place_smem, st>>>(base, L, aos, flag, a_place);
        e = cudaGetLastError();
        }
    if (mark)
        mark(1, st);
    dev_stats().kernel_launches++;
    if (e != cudaSuccess)
        {
        set_last_error(std::string("reorder slot path launch: ") + cudaGetErrorString(e));
        return -1;
        }
    cudaMemcpyAsync(g_slot_flag_host, flag, 12, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
        {
        set_last_error(std::string("reorder slot path: ") + cudaGetErrorString(e));
        return -1;
        }
    if (g_slot_flag_host[1] == 3)
        {
        set_last_error("reorder slot path: a bulk copy did not complete");
        return -1;
        }
    *done = (g_slot_flag_host[0] == 0 && g_slot_flag_host[1] == 0) ? 1 : 0;
    *out_of_range = g_slot_flag_host[2] != 0 ? 1 : 0;
    return 0;
    }

// ---- distributed reorder: ONE frame whose rows are partitioned over the ranks (config 3 read-back) -----------
// "GPU g loads file partition g; destination GPU = id / ceil(N/G); one all-to-all of
// 40-byte rows, then local sort + gather".  Here the all-to-all is not a separate collective: kernels store the
// records straight into the owner's memory -- its own, or a CUDA IPC mapping of the peer's, i.e. plain stores over
// NVLink -- and each owner finishes its buckets locally.
//   1 all-gather  n_local, output capacity, size + IPC handle of the shared buffer (-> geometry, same on every rank;
//                 mappings are cached between calls; a second exchange only when the buffer has to grow)
//   2 k6_slot_hist on the local keys; all-gather of the bucket counts (+ the range flag): every rank now knows
//     where its records start at every destination (sum of the counts of lower ranks: no remote atomics), how
//     many rows every rank will own, and whether a bucket overflows -- identical decisions everywhere
//   3 records to their owners: k6_part_scatter into the owners' inboxes ("partition", default) or k6_slot_scatter
//     into the owners' bucketed copies ("fused"); stream sync; barrier all-gather
//   4 owner: [partition: k6_slot_scatter_rec inbox -> bucketed copy] k6_slot_place; all-gather of the duplicate flags
// Requires unique ids below ceil(N / 2^L) * 2^L (dense ids 0..N-1 qualify); otherwise every rank returns 1
// and the caller gathers the frame to one GPU and uses pgsd_b200_reorder_device.
static void* g_dist_copy = nullptr; // this rank's IPC-exported buffer: [DistCtl + count tables (DIST_HDR_BYTES)][inbox / bucketed copy]
static uint32_t g_dist_epoch = 0;   // calls that reached the exchange; the same on every rank
static inline void* dist_inbox(void* base) { return base ? (void*)((unsigned char*)base + DIST_HDR_BYTES) : nullptr; }
static size_t g_dist_copy_bytes = 0;
static void* g_dist_peer[8] = { nullptr };
static cudaIpcMemHandle_t g_dist_peer_handle[8];
static bool g_dist_peer_open[8] = { false };

static void dist_close_peers()
    {
    for (int p = 0; p < 8; p++)
        {
        if (g_dist_peer_open[p] && g_dist_peer[p])
            cudaIpcCloseMemHandle(g_dist_peer[p]);
        g_dist_peer_open[p] = false;
        g_dist_peer[p] = nullptr;
        }
    }

void dist_release_workspace()
    {
    dist_close_peers();
    if (g_dist_copy)
        cudaFree(g_dist_copy);
    g_dist_copy = nullptr;
    g_dist_copy_bytes = 0;
    }

// PGSD_B200_DIST_TRACE=1: rank 0 prints the host wall time of every stage of a distributed reorder to stderr
struct DistTrace
    {
    bool on;
    std::chrono::steady_clock::time_point t0;
    std::string line;
    explicit DistTrace(bool enable) : on(enable), t0(std::chrono::steady_clock::now()) { }
    void mark(const char* what)
        {
        if (!on)
            return;
        const auto t1 = std::chrono::steady_clock::now();
        char buf[96];
        snprintf(buf, sizeof(buf), " %s %.0f us;", what, std::chrono::duration<double, std::micro>(t1 - t0).count());
        line += buf;
        t0 = t1;
        }
    ~DistTrace()
        {
        if (on)
            fprintf(stderr, "[pgsd_b200 reorder_distributed]%s\n", line.c_str());
        }
    };

int dev_reorder_distributed(uint64_t n_local, const uint32_t* keys, uint64_t out_capacity, uint64_t* n_out, uint64_t* id_first,
                            uint32_t* keys_sorted, int nfields, const ReorderField* fields, void* stream_v)
    {
    int rc = dev_init(-1);
    if (rc != 0)
        return rc;
    Comm* c = comm();
    const int G = c->nprocs, me = c->rank;
    const char* etr = getenv("PGSD_B200_DIST_TRACE");
    DistTrace trace(etr && etr[0] == '1' && me == 0);
    if (G > 8)
        {
        set_last_error("reorder_distributed: at most 8 ranks");
        return -2;
        }
    if (n_out == nullptr)
        {
        set_last_error("reorder_distributed: n_out is required");
        return -2;
        }
    // Rank-local argument problems must not make this rank leave before the collectives below (its peers would wait
    // for it forever): they are folded into the first all-gather and every rank returns the same error.
    // n_local == UINT64_MAX: the caller already knows its arguments are unusable and only takes part.
    bool bad_args = false;
    if (n_local == 0xffffffffffffffffull)
        {
        bad_args = true;
        n_local = 0;
        }
    if (nfields < 0 || (nfields > 0 && fields == nullptr) || nfields + 1 > SLOT_MAX_FIELDS || (n_local > 0 && keys == nullptr)
        || ((uintptr_t)keys & 15u) != 0)
        {
        set_last_error("reorder_distributed: bad arguments (keys must be 16-byte aligned)");
        bad_args = true;
        }
    cudaStream_t st = (cudaStream_t)stream_v;
    SlotArgs a;
    memset(&a, 0, sizeof(a));
    int nf = 0;
    uint32_t off = 0, in_words = 0;
    bool aligned = true;
    a.f[nf++] = SlotField { keys, keys_sorted, 1u, off };
    off += 1;
    in_words += 1;
    for (int i = 0; i < nfields && !bad_args; i++)
        {
        const ReorderField& f = fields[i];
        if (f.row_bytes == 0 || f.row_bytes % 4 != 0 || (((uintptr_t)f.in | (uintptr_t)f.out) & 3u) != 0
            || (n_local > 0 && f.in == nullptr) || (out_capacity > 0 && f.out == nullptr))
            {
            set_last_error("reorder_distributed: fields must be word sized and word aligned");
            bad_args = true;
            break;
            }
        if (((uintptr_t)f.in & 15u) != 0)
            aligned = false;
        a.f[nf++] = SlotField { (const uint32_t*)f.in, (uint32_t*)f.out, f.row_bytes / 4, off };
        off += f.row_bytes / 4;
        in_words += f.row_bytes / 4;
        }
    if (off > SLOT_MAX_ROW_WORDS)
        {
        set_last_error("reorder_distributed: rows of at most 31 words");
        bad_args = true;
        }
    if (bad_args)
        {
        n_local = 0; // nothing of this rank's data is touched
        off = off > SLOT_MAX_ROW_WORDS || off == 0 ? 1 : off;
        }
    a.nfields = nf;
    a.row_words = off;
    a.nranks = G;
    a.ushift = 7;
    a.bulk = aligned ? 1 : 0;

    // (1) sizes, and the IPC handle of the shared copy each rank has right now (valid unless somebody must grow)
    auto export_copy = [&](uint64_t* w9)
        {
        memset(w9, 0, 9 * sizeof(uint64_t));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        w9[8] = g_dist_copy ? 1 : 0;
        if (g_dist_copy && G > 1)
            {
            cudaIpcMemHandle_t h;
            if (cudaIpcGetMemHandle(&h, g_dist_copy) != cudaSuccess)
                {
                cudaGetLastError();
                w9[8] = 0;
                }
            else
                memcpy(w9, &h, 64);
            }
        };
    constexpr size_t W1 = 14;
    uint64_t mine[W1] = { n_local, out_capacity, (uint64_t)g_dist_copy_bytes };
    export_copy(mine + 3);
    mine[12] = bad_args ? 1 : 0;
        {
        // which physical GPU this rank runs on: ranks that SHARE a device (tests, oversubscribed hosts) must not wait
        // for one another inside kernels -- nothing guarantees that their kernels run at the same time
        // (cudaGetDeviceProperties takes milliseconds, more with 8 processes in the driver at once: asked once)
        static uint64_t dev_hash = 0;
        if (dev_hash == 0)
            {
            int dev = 0;
            cudaDeviceProp prop;
            uint64_t hh = 1469598103934665603ull;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess)
                for (size_t i = 0; i < sizeof(prop.uuid.bytes); i++)
                    hh = (hh ^ (unsigned char)prop.uuid.bytes[i]) * 1099511628211ull;
            else
                cudaGetLastError();
            dev_hash = hh | 1ull;
            }
        const uint64_t h = dev_hash;
        mine[13] = h;
        }
    std::vector<uint64_t> all((size_t)G * W1);
    if (c->allgather(mine, all.data(), W1) != 0)
        return -1;
    trace.mark("allgather(sizes+handles)");
    uint64_t N = 0;
    bool any_bad = false;
    for (int p = 0; p < G; p++)
        {
        N += all[(size_t)p * W1];
        any_bad = any_bad || all[(size_t)p * W1 + 12] != 0;
        }
    if (any_bad)
        {
        if (!bad_args)
            set_last_error("reorder_distributed: another rank was called with invalid arguments");
        return -2;
        }
    bool shared_device = false;
    for (int p = 0; p < G; p++)
        for (int q = p + 1; q < G; q++)
            if (all[(size_t)p * W1 + 13] == all[(size_t)q * W1 + 13])
                shared_device = true;
    *n_out = 0;
    if (id_first)
        *id_first = 0;
    if (N == 0)
        return 0;
    if (N >= 0xffffffffull)
        {
        set_last_error("reorder_distributed: fewer than 2^32 - 1 rows in total");
        return -2;
        }
    DistPlan pl;
    if (dist_plan(N, G, &pl) != 0)
        {
        set_last_error("reorder_distributed: frame too large for the slot geometry");
        return -2;
        }
    const int L = pl.L;
    const size_t place_smem = slot_place_smem(L, a.row_words);
    if (place_smem > 220 * 1024)
        {
        set_last_error("reorder_distributed: rows too wide for the slot geometry");
        return -2;
        }
    const uint32_t cap = pl.cap;
    const uint32_t nbp = pl.nbp; // histogram size (power of two)
    const uint32_t bmask = nbp - 1u;
    const uint32_t nb_used = pl.nb_used;
    const uint32_t nbr = pl.nbr; // buckets per owner
    const uint64_t key_limit = (uint64_t)nb_used * cap;             // <= 2^27
    const uint32_t my_b0 = (uint32_t)me * nbr < nb_used ? (uint32_t)me * nbr : nb_used;
    const uint32_t my_b1 = my_b0 + nbr < nb_used ? my_b0 + nbr : nb_used;
    if (id_first)
        *id_first = (uint64_t)me * nbr * cap;
    a.nbl = nbr;
    a.nbr = nbr;
    int tile = 1024;
    while (tile > 512 && ((size_t)in_words + 1) * tile * 4 > 200 * 1024)
        tile /= 2;

    // Two ways to get the records to their owners (same on every rank):
    //   partition (default for > 1 rank): k6_part_scatter appends every record to the owner's INBOX -- a tile writes
    //       one contiguous run per owner, which is what NVLink wants -- then the owner runs k6_slot_scatter_rec +
    //       k6_slot_place on its inbox.  Three passes over the rows, all of them streaming.
    //   fused (PGSD_B200_DIST_MODE=fused, and always for 1 rank): k6_slot_scatter stores every record directly at its
    //       place in the owner's bucketed copy.  Two passes, but the link carries single 40-byte stores
    //       (measured 300 GB/s per direction at 2 GPUs: 3.7 ms against 3.4 ms on one GPU).
    const char* emode = getenv("PGSD_B200_DIST_MODE");
    const bool part = G > 1 && !(emode && !strcmp(emode, "fused"));
    // (2) the shared buffer (inbox / bucketed copy): same size on every rank; reallocation is a collective decision
    const size_t nlines = ((size_t)cap * a.row_words * 4 + 127) / 128;
    const size_t lines_bytes = up256((size_t)nbr * nlines * 128) + 256;
    const size_t copy_need = part ? up256((size_t)nbr * cap * a.row_words * 4) + 256 : lines_bytes;
    bool grow = false;
    for (int p = 0; p < G; p++)
        if (all[(size_t)p * W1 + 2] < copy_need + DIST_HDR_BYTES)
            grow = true;
    std::vector<uint64_t> hall((size_t)G * 9);
    if (grow)
        {
        dist_close_peers();
        const bool synced = cudaStreamSynchronize(st) == cudaSuccess; // a local failure is reported through the handle exchange
        if (!synced)
            cudaGetLastError();
        if (c->barrier() != 0) // nobody maps the old copies any more
            return -1;
        if (g_dist_copy)
            cudaFree(g_dist_copy);
        g_dist_copy = nullptr;
        g_dist_copy_bytes = 0;
        bool ok = synced && cudaMalloc(&g_dist_copy, copy_need + DIST_HDR_BYTES) == cudaSuccess;
        if (ok)
            {
            // flags of the new buffer start at 0 (no call number is 0); the handle exchange below is also the barrier
            // that keeps every peer from writing into it before the memset has finished
            ok = cudaMemset(g_dist_copy, 0, DIST_HDR_BYTES) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
            g_dist_copy_bytes = copy_need + DIST_HDR_BYTES;
            }
        if (!ok)
            cudaGetLastError();
        uint64_t hsend[9];
        export_copy(hsend);
        if (c->allgather(hsend, hall.data(), 9) != 0)
            return -1;
        }
    else
        for (int p = 0; p < G; p++)
            memcpy(&hall[(size_t)p * 9], &all[(size_t)p * W1 + 3], 9 * sizeof(uint64_t));
    for (int p = 0; p < G; p++)
        if (hall[(size_t)p * 9 + 8] == 0)
            {
            set_last_error("reorder_distributed: a rank could not allocate / export its part of the copy");
            return -6;
            }
    for (int p = 0; p < G; p++)
        {
        if (p == me)
            {
            a.peer[p] = (uint32_t*)dist_inbox(g_dist_copy);
            continue;
            }
        cudaIpcMemHandle_t h;
        memcpy(&h, &hall[(size_t)p * 9], 64);
        if (!g_dist_peer_open[p] || memcmp(&h, &g_dist_peer_handle[p], 64) != 0)
            {
            if (g_dist_peer_open[p] && g_dist_peer[p])
                cudaIpcCloseMemHandle(g_dist_peer[p]);
            g_dist_peer_open[p] = false;
            void* ptr = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
                {
                set_last_error(std::string("reorder_distributed: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
                cudaGetLastError();
                // keep the collective sequence intact: the failure is reported through the counts all-gather below
                ptr = nullptr;
                }
            g_dist_peer[p] = ptr;
            g_dist_peer_handle[p] = h;
            g_dist_peer_open[p] = ptr != nullptr;
            }
        a.peer[p] = (uint32_t*)dist_inbox(g_dist_peer[p]);
        }
    bool peers_ok = true;
    for (int p = 0; p < G; p++)
        if (a.peer[p] == nullptr)
            peers_ok = false;

    // local workspace: [flag 256 B][counts nbp][base nbr+1][start nbp][cursors nbp * cstride]
    const uint32_t cstride = 32;
    const size_t tb = up256((size_t)(nbp + 1) * 4);
    const size_t bb = up256((size_t)(nbr + 2) * 4);
    const size_t cb = up256((size_t)nbp * cstride * 4);
    const size_t need = 256 + 2 * tb + bb + cb + 256 + (part ? lines_bytes : 0);
    bool ws_ok = true;
    if (g_slot_ws_bytes < need)
        {
        if (g_slot_ws)
            cudaFree(g_slot_ws);
        g_slot_ws = nullptr;
        g_slot_ws_bytes = 0;
        if (cudaMalloc(&g_slot_ws, need) != cudaSuccess)
            {
            cudaGetLastError();
            ws_ok = false;
            }
        else
            g_slot_ws_bytes = need;
        }
    if (!g_slot_flag_host && cudaHostAlloc((void**)&g_slot_flag_host, 256, cudaHostAllocDefault) != cudaSuccess)
        {
        cudaGetLastError();
        ws_ok = false;
        }
    static bool attr_done = false;
    if (!attr_done)
        {
        cudaFuncSetAttribute(k6_slot_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << SLOT_MAX_BUCKET_BITS) * 4);
        cudaFuncSetAttribute(k6_slot_place, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr_done = true;
        }
    unsigned char* p8 = (unsigned char*)g_slot_ws;
    uint32_t* flag = (uint32_t*)p8;
    uint32_t* counts = (uint32_t*)(p8 + 256);
    uint32_t* base = (uint32_t*)(p8 + 256 + tb);
    uint32_t* start = (uint32_t*)(p8 + 256 + tb + bb);
    uint32_t* cursor = (uint32_t*)(p8 + 256 + 2 * tb + bb);
    uint32_t* owner_cursor = (uint32_t*)(p8 + 256 + 2 * tb + bb + cb);          // 8 words
    uint32_t* lines_copy = (uint32_t*)(p8 + 256 + 2 * tb + bb + cb + 256);      // partition mode only

    trace.mark("peers+workspace");
    // ---- device-driven exchange (default for > 1 rank): counts, "records delivered" and "buckets placed" travel as
    // stores + flags between the GPUs; the host launches everything at once and reads one result at the end.
    // PGSD_B200_DIST_HOST=1 keeps the exchange on the host communicator (3 more all-gathers, host-built tables).
    // Only with one GPU per rank: the kernels of the exchange spin on flags that the peers' kernels write.
    const char* ehost = getenv("PGSD_B200_DIST_HOST");
    if (part && !shared_device && !(ehost && ehost[0] == '1'))
        {
        const uint32_t epoch = ++g_dist_epoch;
        const int par = (int)(epoch & 1u);
        if (!ws_ok)
            {
            set_last_error("reorder_distributed: workspace allocation failed (the peers will report a timeout)");
            return -6;
            }
        DistPeers peers;
        memset(&peers, 0, sizeof(peers));
        for (int p = 0; p < G; p++)
            peers.shared[p] = p == me ? g_dist_copy : g_dist_peer[p];
        uint32_t* n_owned = owner_cursor + 16;
        uint32_t* host_out = g_slot_flag_host + 16; // pinned: [0] plan status, [1] rows owned, [2] delivery wait, [3] final wait
        host_out[0] = host_out[1] = host_out[2] = host_out[3] = 0;
        static bool plan_attr = false;
        if (!plan_attr)
            {
            cudaFuncSetAttribute(kd_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024);
            plan_attr = true;
            }
        cudaMemsetAsync(p8, 0, 256 + tb, st);
        if (n_local > 0)
            {
            int hgrid = dev_sm_count() * (nbp <= 16384 ? 2 : 1);
            const uint64_t want = (n_local / 4 + 1023) / 1024;
            if ((uint64_t)hgrid > want)
                hgrid = want ? (int)want : 1;
            k6_slot_hist<<<hgrid, 1024, (size_t)nbp * 4, st>>>(keys, n_local, L, bmask, nbp, 0u, (uint32_t)key_limit, counts, flag);
            dev_stats().kernel_launches++;
            }
        kd_publish<<<G, 256, 0, st>>>(peers, me, par, epoch, counts, nbp, flag, n_local, out_capacity, peers_ok ? 0u : (uint32_t)DST_RESOURCE);
        kd_plan<<<1, 1024, (size_t)(nbr + 2) * 4, st>>>(g_dist_copy, G, me, par, epoch, nbp, nb_used, nbr, cap, owner_cursor, base, n_owned, flag, host_out);
        cudaMemsetAsync(cursor, 0, (size_t)nbr * cstride * 4, st);
        dev_stats().kernel_launches += 2;
        cudaError_t e = cudaGetLastError();
        SlotArgs ap = a;
        ap.nbl = 0;
        if (e == cudaSuccess && n_local > 0)
            {
            const size_t smem = ((size_t)in_words + 1) * tile * 4 + SLOT_MAX_FIELDS * SLOT_SKEW * 4 + (size_t)tile * 2;
            const uint32_t tiles = (uint32_t)((n_local + tile - 1) / tile);
            if (tile == 512)
                {
                cudaFuncSetAttribute(k6_part_scatter<512, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                k6_part_scatter<512, 128><<<tiles, 128, smem, st>>>(n_local, L, bmask, owner_cursor, flag, ap);
                }
            else
                {
                cudaFuncSetAttribute(k6_part_scatter<1024, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                k6_part_scatter<1024, 256><<<tiles, 256, smem, st>>>(n_local, L, bmask, owner_cursor, flag, ap);
                }
            dev_stats().kernel_launches++;
            e = cudaGetLastError();
            }
        kd_signal<<<1, 32, 0, st>>>(peers, G, me, par, epoch, 0, flag);
        kd_wait<<<1, 32, 0, st>>>(g_dist_copy, G, par, epoch, 0, flag, host_out);
        dev_stats().kernel_launches += 2;
        if (e == cudaSuccess && my_b1 > my_b0)
            {
            SlotArgs aq = a;
            aq.bulk = 1;
            aq.nranks = 1;
            // my inbox (rows decided by kd_plan, any order) -> bucketed copy -> fields in id order
            const int rt = ((size_t)a.row_words + 1) * 1024 * 4 <= 200 * 1024 ? 1024 : 512;
            const size_t smem = ((size_t)a.row_words + 1) * rt * 4;
            const uint32_t tiles = (uint32_t)(((uint64_t)(my_b1 - my_b0) * cap + rt - 1) / rt); // the fullest inbox possible
            if (rt == 512)
                {
                cudaFuncSetAttribute(k6_slot_scatter_rec<512, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                k6_slot_scatter_rec<512, 128><<<tiles, 128, smem, st>>>((const uint32_t*)dist_inbox(g_dist_copy), 0, n_owned, L, bmask,
                                                                       my_b0, cursor, cstride, lines_copy, flag, aq);
                }
            else
                {
                cudaFuncSetAttribute(k6_slot_scatter_rec<1024, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                k6_slot_scatter_rec<1024, 256><<<tiles, 256, smem, st>>>((const uint32_t*)dist_inbox(g_dist_copy), 0, n_owned, L, bmask,
                                                                        my_b0, cursor, cstride, lines_copy, flag, aq);
                }
            k6_slot_place<<<my_b1 - my_b0, cap / 4, place_smem, st>>>(base, L, lines_copy, flag, aq);
            dev_stats().kernel_launches += 2;
            if (e == cudaSuccess)
                e = cudaGetLastError();
            }
        kd_signal<<<1, 32, 0, st>>>(peers, G, me, par, epoch, 1, flag);
        kd_wait<<<1, 32, 0, st>>>(g_dist_copy, G, par, epoch, 1, flag, host_out);
        dev_stats().kernel_launches += 2;
        if (e == cudaSuccess)
            e = cudaGetLastError();
        const cudaError_t es = cudaStreamSynchronize(st);
        trace.mark("device-driven step");
        if (e != cudaSuccess || es != cudaSuccess)
            {
            set_last_error(std::string("reorder_distributed: ") + cudaGetErrorString(e != cudaSuccess ? e : es));
            cudaGetLastError();
            return -1;
            }
        const uint32_t stt = host_out[0] | host_out[2] | host_out[3];
        if (stt & DST_RESOURCE)
            {
            set_last_error("reorder_distributed: a rank failed to allocate or map memory, or a bulk copy did not complete");
            return -6;
            }
        if (stt & DST_TIMEOUT)
            {
            set_last_error("reorder_distributed: a peer did not signal within the time limit");
            return -1;
            }
        if (stt & (DST_RANGE | DST_OVERFLOW))
            return 1; // ids not dense / more ids than slots in a bucket: not applicable, on every rank
        if (stt & DST_FITS)
            {
            set_last_error("reorder_distributed: out_capacity too small on some rank (rows owned: ceil(N / 2^L / ranks) * 2^L at most)");
            return -2;
            }
        if (stt & DST_DUP)
            return 1; // two records with one id: nothing of the outputs is valid, on every rank
        *n_out = host_out[1];
        return 0;
        }
    // (3) local histogram, counts of all ranks
    const size_t cw = ((size_t)nbp + 1) / 2 + 1; // u64 words: packed counts + status word
    std::vector<uint64_t> csend(cw, 0), call((size_t)G * cw);
    std::vector<uint32_t> hcounts((size_t)nbp + 1, 0);
    uint64_t status = (ws_ok && peers_ok) ? 0 : 4; // 1: key out of range, 4: resource failure
    if (ws_ok)
        {
        cudaMemsetAsync(p8, 0, 256 + tb, st);
        if (n_local > 0)
            {
            int hgrid = dev_sm_count() * (nbp <= 16384 ? 2 : 1);
            const uint64_t want = (n_local / 4 + 1023) / 1024;
            if ((uint64_t)hgrid > want)
                hgrid = want ? (int)want : 1;
            k6_slot_hist<<<hgrid, 1024, (size_t)nbp * 4, st>>>(keys, n_local, L, bmask, nbp, 0u, (uint32_t)key_limit, counts, flag);
            dev_stats().kernel_launches++;
            }
        cudaMemcpyAsync(hcounts.data(), counts, (size_t)nbp * 4, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(g_slot_flag_host, flag, 12, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess)
            {
            cudaGetLastError();
            status |= 4;
            }
        else if (g_slot_flag_host[2] != 0)
            status |= 1;
        }
    trace.mark("hist+D2H");
    memcpy(csend.data(), hcounts.data(), (size_t)nbp * 4);
    csend[cw - 1] = status;
    if (c->allgather(csend.data(), call.data(), cw) != 0)
        return -1;
    trace.mark("allgather(counts)");
    uint64_t any = 0;
    for (int p = 0; p < G; p++)
        any |= call[(size_t)p * cw + cw - 1];
    if (any & 4)
        {
        set_last_error("reorder_distributed: a rank failed to allocate or map memory");
        return -6;
        }
    if (any & 1)
        return 1; // ids outside [0, key_limit): not applicable
    // where my records start in every bucket, how full every bucket gets, who owns how many rows
    std::vector<uint32_t> hstart(nbp, 0), hbase((size_t)nbr + 2, 0);
    std::vector<uint64_t> owned(G, 0);
    bool overflow = false;
    std::vector<uint32_t> htot(nbp, 0);
    for (int p = 0; p < G; p++) // rank by rank over contiguous rows of the gathered table
        {
        const uint32_t* cp = reinterpret_cast<const uint32_t*>(&call[(size_t)p * cw]);
        if (p == me)
            memcpy(hstart.data(), htot.data(), (size_t)nbp * 4);
        for (uint32_t b = 0; b < nbp; b++)
            htot[b] += cp[b];
        }
    for (uint32_t b = 0; b < nbp; b++)
        if (htot[b] > cap)
            overflow = true;
    for (uint32_t b = 0; b < nb_used; b++)
        {
        const uint32_t o = b / nbr;
        owned[o] += htot[b];
        if (o == (uint32_t)me)
            hbase[b - my_b0 + 1] = htot[b];
        }
    if (overflow)
        return 1; // more ids than slots in a bucket: duplicates
    bool fits = true;
    for (int p = 0; p < G; p++)
        if (owned[p] > all[(size_t)p * W1 + 1])
            fits = false;
    if (!fits)
        {
        set_last_error("reorder_distributed: out_capacity too small on some rank (rows owned: ceil(N / 2^L / ranks) * 2^L at most)");
        return -2;
        }
    for (uint32_t i = 1; i <= nbr; i++) // counts -> exclusive prefix; entries past my last bucket repeat the total
        hbase[i] += hbase[i - 1];
    hbase[nbr + 1] = hbase[nbr];
    *n_out = owned[me];
    cudaMemcpyAsync(base, hbase.data(), (size_t)(nbr + 1) * 4, cudaMemcpyHostToDevice, st);
    cudaError_t e = cudaSuccess;
    uint32_t hown[8] = { 0 };
    if (part)
        {
        // where my run starts in every owner's inbox: the rows of lower ranks for that owner come first
        for (int o = 0; o < G; o++)
            {
            const uint32_t b0 = (uint32_t)o * nbr < nb_used ? (uint32_t)o * nbr : nb_used;
            const uint32_t b1 = b0 + nbr < nb_used ? b0 + nbr : nb_used;
            for (int p = 0; p < me; p++)
                {
                const uint32_t* cp = reinterpret_cast<const uint32_t*>(&call[(size_t)p * cw]);
                for (uint32_t b = b0; b < b1; b++)
                    hown[o] += cp[b];
                }
            }
        cudaMemcpyAsync(owner_cursor, hown, sizeof(hown), cudaMemcpyHostToDevice, st);
        trace.mark("tables+H2D");
        // (4) records to their owners' inboxes
        SlotArgs ap = a;
        ap.nbl = 0;
        if (n_local > 0)
            {
            const size_t smem = ((size_t)in_words + 1) * tile * 4 + SLOT_MAX_FIELDS * SLOT_SKEW * 4 + (size_t)tile * 2;
            const uint32_t tiles = (uint32_t)((n_local + tile - 1) / tile);
            if (tile == 512)
                {
                cudaFuncSetAttribute(k6_part_scatter<512, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                k6_part_scatter<512, 128><<<tiles, 128, smem, st>>>(n_local, L, bmask, owner_cursor, flag, ap);
                }
            else
                {
                cudaFuncSetAttribute(k6_part_scatter<1024, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                k6_part_scatter<1024, 256><<<tiles, 256, smem, st>>>(n_local, L, bmask, owner_cursor, flag, ap);
                }
            dev_stats().kernel_launches++;
            e = cudaGetLastError();
            }
        }
    else
        {
        cudaMemcpyAsync(start, hstart.data(), (size_t)nbp * 4, cudaMemcpyHostToDevice, st);
        k6_slot_spread<<<(nbp + 255) / 256, 256, 0, st>>>(start, nbp, cursor, cstride);
        dev_stats().kernel_launches++;
        trace.mark("tables+H2D");
        // (4) records to their places in the owners' bucketed copies
        e = cudaGetLastError();
        if (e == cudaSuccess && n_local > 0)
            {
            const uint32_t tiles_all = (uint32_t)((n_local + tile - 1) / tile);
            if (tile == 512)
                e = launch_scatter<512, 256>(n_local, 0, tiles_all, L, bmask, cursor, cstride, (uint32_t*)dist_inbox(g_dist_copy), flag, a, in_words, st);
            else
                e = launch_scatter<1024, 512>(n_local, 0, tiles_all, L, bmask, cursor, cstride, (uint32_t*)dist_inbox(g_dist_copy), flag, a, in_words, st);
            }
        }
    cudaMemcpyAsync(g_slot_flag_host, flag, 12, cudaMemcpyDeviceToHost, st);
    uint64_t st4 = 0;
    if (e != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
        {
        set_last_error(std::string("reorder_distributed scatter: ") + cudaGetErrorString(e != cudaSuccess ? e : cudaGetLastError()));
        st4 = 4;
        }
    else if (g_slot_flag_host[1] == 3)
        st4 = 4;
    trace.mark("scatter+sync");
    std::vector<uint64_t> sall(G);
    if (c->allgather(&st4, sall.data(), 1) != 0) // also the barrier: every record has reached its owner
        return -1;
    trace.mark("allgather(barrier)");
    for (int p = 0; p < G; p++)
        if (sall[p] != 0)
            {
            if (st4 == 0)
                set_last_error("reorder_distributed: the scatter failed on another rank");
            return -1;
            }

    // (5) my buckets
    uint64_t st5 = 0;
    if (my_b1 > my_b0)
        {
        SlotArgs ap = a;
        ap.bulk = 1;
        ap.nranks = 1;
        const uint32_t* bucketed = (const uint32_t*)dist_inbox(g_dist_copy);
        if (part)
            {
            // my inbox (owned[me] records, any order) -> bucketed copy, with the cursors starting at 0
            cudaMemsetAsync(cursor, 0, (size_t)nbr * cstride * 4, st);
            const int rt = ((size_t)a.row_words + 1) * 1024 * 4 <= 200 * 1024 ? 1024 : 512;
            const size_t smem = ((size_t)a.row_words + 1) * rt * 4;
            const uint32_t tiles = (uint32_t)((owned[me] + rt - 1) / rt);
            if (tiles)
                {
                if (rt == 512)
                    {
                    cudaFuncSetAttribute(k6_slot_scatter_rec<512, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    k6_slot_scatter_rec<512, 128><<<tiles, 128, smem, st>>>((const uint32_t*)dist_inbox(g_dist_copy), owned[me], nullptr, L, bmask, my_b0,
                                                                           cursor, cstride, lines_copy, flag, ap);
                    }
                else
                    {
                    cudaFuncSetAttribute(k6_slot_scatter_rec<1024, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    k6_slot_scatter_rec<1024, 256><<<tiles, 256, smem, st>>>((const uint32_t*)dist_inbox(g_dist_copy), owned[me], nullptr, L, bmask, my_b0,
                                                                            cursor, cstride, lines_copy, flag, ap);
                    }
                dev_stats().kernel_launches++;
                }
            bucketed = lines_copy;
            }
        k6_slot_place<<<my_b1 - my_b0, cap / 4, place_smem, st>>>(base, L, bucketed, flag, ap);
        dev_stats().kernel_launches++;
        e = cudaGetLastError();
        cudaMemcpyAsync(g_slot_flag_host, flag, 12, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
            {
            set_last_error(std::string("reorder_distributed place: ") + cudaGetErrorString(e != cudaSuccess ? e : cudaGetLastError()));
            st5 = 4;
            }
        else if (g_slot_flag_host[1] != 0)
            st5 = 1; // two records of a bucket share a slot
        }
    trace.mark("place+sync");
    if (c->allgather(&st5, sall.data(), 1) != 0)
        return -1;
    trace.mark("allgather(flags)");
    any = 0;
    for (int p = 0; p < G; p++)
        any |= sall[p];
    if (any & 4)
        return -1;
    return (any & 1) ? 1 : 0;
    }
} // namespace pgsdb
