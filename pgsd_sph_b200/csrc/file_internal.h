// file_internal.h -- entry points of pgsd_file.cpp used by api_b200.cpp (pgsd_b200.h ABI).
#pragma once
#include "../../include/pgsd.h"
#include "device.h"

namespace pgsdb
{
// One chunk whose bytes are produced by K1 (pack + cast) from `cols`.
struct DeviceChunk
    {
    const char* name;
    int dst_type, src_type;
    uint64_t N; // rows packed by K1
    uint32_t M; // columns packed by K1
    uint64_t N_global;
    uint32_t M_global;
    uint64_t offset; // elements, or PGSD_B200_OFFSET_AUTO
    bool all;
    const Column* cols;
    bool host_columns;
    };
// pgsd_write_chunk for n chunks at once: ONE K1 launch packs them all into the frame arena.
int file_write_chunks_device(pgsd_handle* h, int n, const DeviceChunk* chunks);
// pgsd_write_chunk for a chunk whose bytes are produced by K1 from `cols` (device pointers, or
// host pointers when host_columns is set); same validation and return codes as pgsd_write_chunk.
int file_write_chunk_device(pgsd_handle* h, const char* name, int dst_type, uint64_t N, uint32_t M,
                            uint64_t N_global, uint32_t M_global, uint64_t offset, bool all, int src_type,
                            const Column* cols, bool host_columns);
int file_read_to_device(pgsd_handle* h, void* dev_dst, uint64_t bytes, uint64_t file_off);
} // namespace pgsdb
