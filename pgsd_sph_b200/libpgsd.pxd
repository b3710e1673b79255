# libpgsd.pxd -- Cython declarations of the C ABI of libpgsd_b200.so (include/pgsd.h, include/pgsd_b200.h).
# Counterpart of the reference's /root/reference/pgsd/pgsd/libpgsd.pxd:9-144; no mpi4py cimport: the
# rank communicator is set through pgsd_b200_comm_init_*.
from libc.stdint cimport uint8_t, uint16_t, uint32_t, uint64_t, int64_t

cdef extern from "pgsd.h" nogil:
    cdef enum pgsd_type:
        PGSD_TYPE_UINT8
        PGSD_TYPE_UINT16
        PGSD_TYPE_UINT32
        PGSD_TYPE_UINT64
        PGSD_TYPE_INT8
        PGSD_TYPE_INT16
        PGSD_TYPE_INT32
        PGSD_TYPE_INT64
        PGSD_TYPE_FLOAT
        PGSD_TYPE_DOUBLE

    cdef enum pgsd_open_flag:
        PGSD_OPEN_READWRITE
        PGSD_OPEN_READONLY
        PGSD_OPEN_APPEND

    cdef enum pgsd_error:
        PGSD_SUCCESS
        PGSD_ERROR_IO
        PGSD_ERROR_INVALID_ARGUMENT
        PGSD_ERROR_NOT_A_PGSD_FILE
        PGSD_ERROR_INVALID_PGSD_FILE_VERSION
        PGSD_ERROR_FILE_CORRUPT
        PGSD_ERROR_MEMORY_ALLOCATION_FAILED
        PGSD_ERROR_NAMELIST_FULL
        PGSD_ERROR_FILE_MUST_BE_WRITABLE
        PGSD_ERROR_FILE_MUST_BE_READABLE

    cdef struct pgsd_header:
        uint64_t magic
        uint64_t index_location
        uint64_t index_allocated_entries
        uint64_t namelist_location
        uint64_t namelist_allocated_entries
        uint32_t schema_version
        uint32_t pgsd_version
        char application[64]
        char schema[64]
        char reserved[80]

    cdef struct pgsd_index_entry:
        uint64_t frame
        uint64_t N
        int64_t location
        uint32_t M
        uint16_t id
        uint8_t type
        uint8_t flags

    # the real layout comes from pgsd.h (544 bytes); only what the Python layer reads is named here
    cdef struct pgsd_handle:
        void* fh
        pgsd_header header
        uint64_t cur_frame
        long long file_size

    uint32_t pgsd_make_version(unsigned int major, unsigned int minor)
    int pgsd_create_and_open(pgsd_handle* handle, const char* fname, const char* application, const char* schema,
                             uint32_t schema_version, pgsd_open_flag flags, int exclusive_create)
    int pgsd_open(pgsd_handle* handle, const char* fname, pgsd_open_flag flags)
    int pgsd_close(pgsd_handle* handle)
    int pgsd_end_frame(pgsd_handle* handle)
    int pgsd_flush(pgsd_handle* handle)
    int pgsd_write_chunk(pgsd_handle* handle, const char* name, pgsd_type type, uint64_t N, uint32_t M,
                         uint64_t N_global, uint32_t M_global, uint64_t offset, uint64_t global_size, bint all,
                         uint8_t flags, const void* data)
    const pgsd_index_entry* pgsd_find_chunk(pgsd_handle* handle, uint64_t frame, const char* name)
    int pgsd_read_chunk(pgsd_handle* handle, void* data, const pgsd_index_entry* chunk, uint64_t N, uint32_t M,
                        uint32_t offset, bint all)
    uint64_t pgsd_get_nframes(pgsd_handle* handle)
    uint64_t pgsd_get_nnames(pgsd_handle* handle)
    size_t pgsd_sizeof_type(pgsd_type type)
    const char* pgsd_find_matching_chunk_name(pgsd_handle* handle, const char* match, const char* prev)
    uint64_t pgsd_get_maximum_write_buffer_size(pgsd_handle* handle)
    int pgsd_set_maximum_write_buffer_size(pgsd_handle* handle, uint64_t size)
    uint64_t pgsd_get_index_entries_to_buffer(pgsd_handle* handle)
    int pgsd_set_index_entries_to_buffer(pgsd_handle* handle, uint64_t number)

cdef extern from "pgsd_b200.h" nogil:
    cdef struct pgsd_b200_column:
        const void* base
        int64_t stride

    cdef struct pgsd_b200_chunk_desc:
        const char* name
        pgsd_type dst_type
        pgsd_type src_type
        uint64_t N
        uint32_t M
        uint64_t N_global
        uint32_t M_global
        uint64_t offset
        bint all
        const pgsd_b200_column* cols

    int pgsd_b200_comm_size()
    const char* pgsd_b200_last_error()
    int pgsd_b200_write_chunk_soa(pgsd_handle* handle, const char* name, pgsd_type dst_type, uint64_t N, uint32_t M,
                                  uint64_t N_global, uint32_t M_global, uint64_t offset, bint all,
                                  pgsd_type src_type, const pgsd_b200_column* cols)
    int pgsd_b200_write_chunks_soa(pgsd_handle* handle, int n_chunks, const pgsd_b200_chunk_desc* chunks)
