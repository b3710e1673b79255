/* pgsd.h -- C ABI of the B200-native PGSD file layer (libpgsd_b200.so).
 *
 * Drop-in for the reference's libpgsd: every entry point, enum value and public
 * struct below has the same name, argument meaning, error behaviour and LP64 layout
 * as the reference interface it replaces (cited as `ref: file:line`, paths relative
 * to /root/reference/).  A caller of the reference (its Cython module fl.pyx, its C++
 * benchmarks, HOOMD-SPH's dump writer) recompiles against this header unchanged.
 *
 * Differences a caller can observe:
 *   - no <mpi.h>: ranks come from pgsd_b200_comm_init_*() (include/pgsd_b200.h), which
 *     stands where MPI_Init/MPI_COMM_WORLD stood; with no communicator the library
 *     runs as rank 0 of 1.
 *   - `fh` is an opaque pointer to library state instead of an MPI_File (same size,
 *     still NULL when no file is open -- ref: pgsd.c:1494).
 *   - `data` of pgsd_write_chunk may be a CUDA device pointer (detected with
 *     cudaPointerGetAttributes); it is then packed by the sm_100a kernels of
 *     pgsd_sph_b200/csrc and staged to the file through pinned buffers.
 *   - every rank keeps a replica of the index and namelist, so pgsd_find_chunk and
 *     pgsd_find_matching_chunk_name return usable pointers on all ranks (the reference
 *     returns a dangling pointer on non-root ranks, pgsd.c:2378).
 */
#ifndef PGSD_B200_PGSD_H
#define PGSD_B200_PGSD_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element type codes stored in pgsd_index_entry::type -- ref: pgsd/pgsd/pgsd.h:38-69 */
enum pgsd_type
    {
    PGSD_TYPE_UINT8 = 1,
    PGSD_TYPE_UINT16 = 2,
    PGSD_TYPE_UINT32 = 3,
    PGSD_TYPE_UINT64 = 4,
    PGSD_TYPE_INT8 = 5,
    PGSD_TYPE_INT16 = 6,
    PGSD_TYPE_INT32 = 7,
    PGSD_TYPE_INT64 = 8,
    PGSD_TYPE_FLOAT = 9,
    PGSD_TYPE_DOUBLE = 10
    };

/* ref: pgsd/pgsd/pgsd.h:72-82 */
enum pgsd_open_flag
    {
    PGSD_OPEN_READWRITE = 1,
    PGSD_OPEN_READONLY = 2,
    PGSD_OPEN_APPEND = 3
    };

/* ref: pgsd/pgsd/pgsd.h:85-120 */
enum pgsd_error
    {
    PGSD_SUCCESS = 0,
    PGSD_ERROR_IO = -1,
    PGSD_ERROR_INVALID_ARGUMENT = -2,
    PGSD_ERROR_NOT_A_PGSD_FILE = -3,
    PGSD_ERROR_INVALID_PGSD_FILE_VERSION = -4,
    PGSD_ERROR_FILE_CORRUPT = -5,
    PGSD_ERROR_MEMORY_ALLOCATION_FAILED = -6,
    PGSD_ERROR_NAMELIST_FULL = -7,
    PGSD_ERROR_FILE_MUST_BE_WRITABLE = -8,
    PGSD_ERROR_FILE_MUST_BE_READABLE = -9
    };

enum { PGSD_NAME_SIZE = 64 };      /* ref: pgsd.h:122-128 */
enum { PGSD_RESERVED_BYTES = 80 }; /* ref: pgsd.h:130-134 */

/* 256-byte file header at offset 0 -- ref: pgsd.h:143-174 */
struct pgsd_header
    {
    uint64_t magic;                      /*   0 */
    uint64_t index_location;             /*   8 */
    uint64_t index_allocated_entries;    /*  16 */
    uint64_t namelist_location;          /*  24 */
    uint64_t namelist_allocated_entries; /*  32  (bytes / PGSD_NAME_SIZE) */
    uint32_t schema_version;             /*  40 */
    uint32_t pgsd_version;               /*  44 */
    char application[PGSD_NAME_SIZE];    /*  48 */
    char schema[PGSD_NAME_SIZE];         /* 112 */
    char reserved[PGSD_RESERVED_BYTES];  /* 176 */
    };

/* 32-byte index entry; N and M are GLOBAL sizes, location == 0 ends the index
   -- ref: pgsd.h:182-204 */
struct pgsd_index_entry
    {
    uint64_t frame;
    uint64_t N;
    int64_t location;
    uint32_t M;
    uint16_t id;
    uint8_t type;
    uint8_t flags;
    };

/* The four structs below are public only because pgsd_handle embeds them
   (ref: pgsd.h:210-286).  name_map is unused by this implementation (kept for layout). */
struct pgsd_name_id_pair
    {
    char* name;
    struct pgsd_name_id_pair* next;
    uint16_t id;
    };

struct pgsd_name_id_map
    {
    struct pgsd_name_id_pair* v;
    size_t size;
    };

struct pgsd_index_buffer
    {
    struct pgsd_index_entry* data;
    size_t size;
    size_t reserved;
    void* mapped_data;
    size_t mapped_len;
    };

struct pgsd_byte_buffer
    {
    char* data;
    size_t size;
    size_t reserved;
    };

struct pgsd_name_buffer
    {
    struct pgsd_byte_buffer data;
    size_t n_names;
    };

/* Caller-allocated handle, 544 bytes on LP64 -- ref: pgsd.h:297-353.
   All members are read-only to the caller.  The buffers it points at are owned by the
   library state behind `fh` and are kept in sync at every API call boundary. */
struct pgsd_handle
    {
    void* fh;                              /*   0  opaque (reference: MPI_File) */
    struct pgsd_header header;             /*   8 */
    struct pgsd_index_buffer file_index;   /* 264  committed entries, sorted by (frame,id) */
    struct pgsd_index_buffer frame_index;  /* 304  entries waiting for the next flush */
    struct pgsd_index_buffer buffer_index; /* 344  entries of buffered small chunks */
    struct pgsd_byte_buffer write_buffer;  /* 384  this rank's buffered small chunks */
    struct pgsd_name_buffer file_names;    /* 408 */
    struct pgsd_name_buffer frame_names;   /* 440 */
    uint64_t cur_frame;                    /* 472 */
    long long int file_size;               /* 480 */
    enum pgsd_open_flag open_flags;        /* 488 */
    struct pgsd_name_id_map name_map;      /* 496 */
    uint64_t pending_index_entries;        /* 512 */
    uint64_t maximum_write_buffer_size;    /* 520 */
    uint64_t index_entries_to_buffer;      /* 528 */
    int rank;                              /* 536 */
    int nprocs;                            /* 540 */
    };

/* ref: pgsd.h:362, pgsd.c:1705-1708 */
uint32_t pgsd_make_version(unsigned int major, unsigned int minor);

/* Create (truncate) a file, write header + 128 zero index entries + 1024 zero namelist
   bytes on rank 0, then open it.  flags: READWRITE or APPEND.
   ref: pgsd.h:412-418, pgsd.c:1710-1773, :1414-1474 */
int pgsd_create_and_open(struct pgsd_handle* handle,
                         const char* fname,
                         const char* application,
                         const char* schema,
                         uint32_t schema_version,
                         enum pgsd_open_flag flags,
                         int exclusive_create);

/* ref: pgsd.h:440, pgsd.c:1775-1812, :1484-1703 */
int pgsd_open(struct pgsd_handle* handle, const char* fname, enum pgsd_open_flag flags);

/* Flush (writable files), release all library state, close.  ref: pgsd.h:480, pgsd.c:1814-1914 */
int pgsd_close(struct pgsd_handle* handle);

/* Finish the current frame; commits it to the file when it holds direct chunks or when
   more than index_entries_to_buffer buffered entries are waiting.
   ref: pgsd.h:498, pgsd.c:1916-1953 */
int pgsd_end_frame(struct pgsd_handle* handle);

/* Write names, buffered chunks and completed-frame index entries.  ref: pgsd.h:517, pgsd.c:1955-2070 */
int pgsd_flush(struct pgsd_handle* handle);

/* Add an N x M chunk of `type` to the current frame (collective, same order on all ranks).
     N_global, M_global : sizes recorded in the index entry
     offset             : this rank's start inside the chunk, in ELEMENTS (not rows/bytes)
     global_size        : ignored (dead in the reference, pgsd.c:2147-2151)
     all                : true  -> every rank writes its N x M slice at file_size+offset
                          false -> buffered when max-over-ranks size < maximum_write_buffer_size,
                                   otherwise only rank 0 writes
     data               : host pointer, or CUDA device pointer (B200 extension)
   ref: pgsd.h:551-564, pgsd.c:2072-2259 */
int pgsd_write_chunk(struct pgsd_handle* handle,
                     const char* name,
                     enum pgsd_type type,
                     uint64_t N,
                     uint32_t M,
                     uint64_t N_global,
                     uint32_t M_global,
                     uint64_t offset,
                     uint64_t global_size,
                     bool all,
                     uint8_t flags,
                     const void* data);

/* Look up (frame, name); pointer into handle->file_index.data, valid until the next
   flush/close.  ref: pgsd.h:581-582, pgsd.c:2295-2434 */
const struct pgsd_index_entry*
pgsd_find_chunk(struct pgsd_handle* handle, uint64_t frame, const char* name);

/* all == false: read the whole chunk (N, M, offset ignored).
   all == true : read N x M elements starting `offset` ROWS into the chunk.
   `data` may be a CUDA device pointer (B200 extension).
   ref: pgsd.h:604-610, pgsd.c:2436-2537 */
int pgsd_read_chunk(struct pgsd_handle* handle,
                    void* data,
                    const struct pgsd_index_entry* chunk,
                    uint64_t N,
                    uint32_t M,
                    uint32_t offset,
                    bool all);

uint64_t pgsd_get_nframes(struct pgsd_handle* handle); /* ref: pgsd.h:620, pgsd.c:2261-2276 */
uint64_t pgsd_get_nnames(struct pgsd_handle* handle);  /* ref: pgsd.h:630, pgsd.c:2279-2292 */
size_t pgsd_sizeof_type(enum pgsd_type type);          /* ref: pgsd.h:638, pgsd.c:2539-2555 */

/* Iterate names starting with `match`; `prev` = previous return value or NULL.
   ref: pgsd.h:659-660, pgsd.c:2557-2641 */
const char*
pgsd_find_matching_chunk_name(struct pgsd_handle* handle, const char* match, const char* prev);

/* ref: pgsd.h:686-729, pgsd.c:2643-2683 */
uint64_t pgsd_get_maximum_write_buffer_size(struct pgsd_handle* handle);
int pgsd_set_maximum_write_buffer_size(struct pgsd_handle* handle, uint64_t size);
uint64_t pgsd_get_index_entries_to_buffer(struct pgsd_handle* handle);
int pgsd_set_index_entries_to_buffer(struct pgsd_handle* handle, uint64_t number);

/* Broadcast one index entry from rank 0 (exported by the reference, unused by it).
   ref: pgsd.h:735, pgsd.c:152-172 */
void pgsd_bcast_index_entry(struct pgsd_index_entry* e);

/* ref: pgsd.h:737-742 (there: MPI_Comm_rank(MPI_COMM_WORLD) == 0; a function here
   because the communicator lives in the library, not in a header) */
bool is_root(void);

#ifdef __cplusplus
}
#endif
#endif /* PGSD_B200_PGSD_H */
