// slot_common.cuh -- declarations shared by the reorder kernels for unique ids (kernels_slot.cu: bucket scatter +
// per-bucket placement, distributed variants; kernels_cluster.cu: coarse partition + cluster placement).
#pragma once
#include "device_internal.h"

namespace pgsdb
{
namespace slotk
    {
constexpr int SLOT_MAX_FIELDS = 18;    // key + 16 caller fields + original index
constexpr int SLOT_MAX_ROW_WORDS = 32; // one warp store covers >= 1 row
constexpr int SLOT_MIN_BITS = 10;
constexpr int SLOT_MAX_BITS = 12;
constexpr int SLOT_MAX_BUCKET_BITS = 15; // 32768 buckets: 128 KB histogram in shared memory
constexpr uint32_t SLOT_SKEW = 4;        // words between the field tiles of a staged tile: the columns of one record fall
                                         // into different banks, and every tile stays 16-byte aligned for the bulk copies

struct SlotField
    {
    const uint32_t* in; // n rows of `words` words; NULL: the row's original index
    uint32_t* out;      // destination of the reordered field; NULL: not wanted
    uint32_t words;
    uint32_t off;       // word offset inside the interleaved row
    };
struct SlotArgs
    {
    SlotField f[SLOT_MAX_FIELDS];
    int nfields;
    uint32_t row_words;
    int bulk; // every input is 16-byte aligned: full tiles are staged with cp.async.bulk
    uint32_t nbl; // 0: buckets are contiguous in the interleaved copy ("flat").  nb: "lines" layout -- 128-byte line j of
                  // bucket b lives at line j * nb + b, so that the lines the buckets are currently filling (about the
                  // same j for all of them) form one compact, advancing window instead of nb windows spread over the
                  // whole copy: the L2 write-backs then fall into few open DRAM rows
    int ushift; // "lines" layout: log2 of the interleaving unit in bytes (7: 128-byte lines)
    // distributed reorder: rank `o` owns the buckets [o * nbr, (o + 1) * nbr); peer[o] is its interleaved copy
    // (own memory, or a CUDA IPC mapping of the owner's memory: the record stores then travel over NVLink)
    uint32_t* peer[8];
    uint32_t nbr;
    int nranks; // 1: single-GPU reorder, peer[] unused
    };

// ---- PTX wrappers: mbarrier + 1-D bulk copy global -> shared (TMA engine) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
    {
    return (uint32_t)__cvta_generic_to_shared(p);
    }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
    {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
    {
    unsigned long long state;
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 %0, [%1], %2;"
                 : "=l"(state)
                 : "r"(bar), "r"(bytes)
                 : "memory");
    (void)state;
    }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
    {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
    }
// Waits for phase `parity`; gives up after `limit` cycles (default: ~2 s worth) so that a lost copy cannot hang the GPU.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, long long limit = 4000000000ll)
    {
    const long long t0 = clock64();
    for (;;)
        {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (done)
            return true;
        if (clock64() - t0 > limit)
            return false;
        }
    }

// records out: a group of lanes writes one record (consecutive words), G records per warp store
// order != NULL: the loop runs over tile positions j sorted by destination and row = order[j], so that the records
// of one warp store are neighbours at their destination (partition by owner: long contiguous runs).
template <int NT>
__device__ __forceinline__ void slot_records_out(const uint32_t* __restrict__ raw, const uint32_t* __restrict__ sdst,
                                                 const uint32_t* __restrict__ col, uint64_t tile0, uint32_t tile_n,
                                                 uint32_t RW, uint32_t* __restrict__ aos, const SlotArgs& args, int lane, int w,
                                                 const uint16_t* __restrict__ order = nullptr)
    {
    constexpr uint32_t NW = NT / 32;
    constexpr int U = 4;
    if ((RW & 1u) == 0)
        {
        const uint32_t R2 = RW / 2, G = 32u / R2;
        const uint32_t g = (uint32_t)lane / R2, c2 = (uint32_t)lane - g * R2;
        if (g < G)
            {
            const uint32_t ca = col[2 * c2], cb = col[2 * c2 + 1];
            const uint32_t Wa = ca & 255u, fa = ca >> 8, Wb = cb & 255u, fb = cb >> 8;
            uint2* aos2 = reinterpret_cast<uint2*>(aos);
            const uint32_t step = NW * G;
            for (uint32_t r0 = (uint32_t)w * G + g; r0 < tile_n; r0 += step * U)
                {
                uint32_t dd[U];
                uint2 v[U];
#pragma unroll
                for (int u = 0; u < U; u++)
                    {
                    const uint32_t j = r0 + (uint32_t)u * step;
                    if (j < tile_n)
                        {
                        const uint32_t r = order ? order[j] : j;
                        dd[u] = sdst[r];
                        v[u].x = Wa ? raw[fa + r * Wa] : (uint32_t)(tile0 + r);
                        v[u].y = Wb ? raw[fb + r * Wb] : (uint32_t)(tile0 + r);
                        }
                    }
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (r0 + (uint32_t)u * step < tile_n)
                        {
                        if (args.nbl)
                            {
                            const uint32_t o = (dd[u] & 4095u) * (RW * 4u) + c2 * 8u;
                            const uint64_t at = ((uint64_t)((o >> args.ushift) * args.nbl + ((dd[u] >> 12) & 32767u)) << args.ushift) + (o & ((1u << args.ushift) - 1u));
                            unsigned char* copy = reinterpret_cast<unsigned char*>(args.nranks > 1 ? args.peer[dd[u] >> 27] : aos);
                            *reinterpret_cast<uint2*>(copy + at) = v[u];
                            }
                        else if (args.nranks > 1) // record index in the owner's inbox
                            reinterpret_cast<uint2*>(args.peer[dd[u] >> 27])[(uint64_t)(dd[u] & 0x7ffffffu) * R2 + c2] = v[u];
                        else
                            aos2[(uint64_t)dd[u] * R2 + c2] = v[u];
                        }
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
            for (uint32_t r0 = (uint32_t)w * G + g; r0 < tile_n; r0 += step * U)
                {
                uint32_t dd[U], v[U];
#pragma unroll
                for (int u = 0; u < U; u++)
                    {
                    const uint32_t j = r0 + (uint32_t)u * step;
                    if (j < tile_n)
                        {
                        const uint32_t r = order ? order[j] : j;
                        dd[u] = sdst[r];
                        v[u] = W ? raw[fb + r * W] : (uint32_t)(tile0 + r);
                        }
                    }
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (r0 + (uint32_t)u * step < tile_n)
                        {
                        if (args.nbl)
                            {
                            const uint32_t o = (dd[u] & 4095u) * (RW * 4u) + c * 4u;
                            const uint64_t at = ((uint64_t)((o >> args.ushift) * args.nbl + ((dd[u] >> 12) & 32767u)) << args.ushift) + (o & ((1u << args.ushift) - 1u));
                            unsigned char* copy = reinterpret_cast<unsigned char*>(args.nranks > 1 ? args.peer[dd[u] >> 27] : aos);
                            *reinterpret_cast<uint32_t*>(copy + at) = v[u];
                            }
                        else if (args.nranks > 1)
                            args.peer[dd[u] >> 27][(uint64_t)(dd[u] & 0x7ffffffu) * RW + c] = v[u];
                        else
                            aos[(uint64_t)dd[u] * RW + c] = v[u];
                        }
                }
            }
        }
    }

    } // namespace slotk
} // namespace pgsdb
