// device_internal.h -- shared by the .cu translation units of libpgsd_b200 only.
#pragma once
#include "device.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <string>

namespace pgsdb
{
int dev_sm_count();               // valid after dev_init()
void sort_release_workspace();    // kernels_sort.cu
int dev_reorder_rows(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm, int nfields,
                     const ReorderField* fields, void* stream); // kernels_sort.cu

// kernels_slot.cu: reorder for unique keys; *done = 0 -> not applicable / duplicates, run the general path
void slot_release_workspace();
int dev_reorder_slot(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm, int nfields,
                     const ReorderField* fields, int topbit, int guessed, uint32_t key_const, void* stream, int* done,
                     int* out_of_range, void (*mark)(int, cudaStream_t));

// kernels_cluster.cu: coarse partition + cluster placement for large frames with unique keys; *handled = 0 ->
// geometry does not fit, nothing launched
void cluster_release_workspace();
int dev_reorder_cluster(uint64_t n, const uint32_t* keys, uint32_t* keys_sorted, uint32_t* perm, int nfields,
                        const ReorderField* fields, int topbit, uint32_t key_const, void* stream, int* done,
                        int* out_of_range, void (*mark)(int, cudaStream_t), int* handled);

// pgsd_type codes (include/pgsd.h; ref: /root/reference/pgsd/pgsd/pgsd.h:38-69)
enum : int
    {
    T_U8 = 1,
    T_U16 = 2,
    T_U32 = 3,
    T_U64 = 4,
    T_I8 = 5,
    T_I16 = 6,
    T_I32 = 7,
    T_I64 = 8,
    T_F32 = 9,
    T_F64 = 10
    };

// ---- K1 launch interface (kernels_pack.cu)
constexpr int PACK_MAX_COLS = 8;
constexpr int PACK_MAX_SEGS = 16;
struct PackSegment
    {
    void* dst;                          // packed (N, M) array of dst_type, device
    const void* base[PACK_MAX_COLS];    // device column bases
    long long stride[PACK_MAX_COLS];    // in elements of src_type
    unsigned long long N;
    unsigned int M;
    int src_type;
    int dst_type;
    };
// one launch packs all segments (the chunks of one frame)
int pack_launch(const PackSegment* segs, int nsegs, cudaStream_t st);
} // namespace pgsdb
