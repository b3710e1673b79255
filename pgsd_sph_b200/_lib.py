"""ctypes binding of libpgsd_b200.so -- the C ABI declared in include/pgsd.h and include/pgsd_b200.h.

This is the reference-side stub a maintainer would write for the ``pgsd.fl`` layer
(the reference binds the same functions from Cython: /root/reference/pgsd/pgsd/libpgsd.pxd:9-144).
The library is REQUIRED: there is no Python or CPU fallback for anything it does.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpgsd_b200.so")

PGSD_NAME_SIZE = 64
PGSD_RESERVED_BYTES = 80

# enum pgsd_type (ref: pgsd.h:38-69)
TYPE_UINT8, TYPE_UINT16, TYPE_UINT32, TYPE_UINT64 = 1, 2, 3, 4
TYPE_INT8, TYPE_INT16, TYPE_INT32, TYPE_INT64 = 5, 6, 7, 8
TYPE_FLOAT, TYPE_DOUBLE = 9, 10
# enum pgsd_open_flag (ref: pgsd.h:72-82)
OPEN_READWRITE, OPEN_READONLY, OPEN_APPEND = 1, 2, 3
# enum pgsd_error (ref: pgsd.h:85-120)
SUCCESS = 0
ERROR_IO = -1
ERROR_INVALID_ARGUMENT = -2
ERROR_NOT_A_PGSD_FILE = -3
ERROR_INVALID_PGSD_FILE_VERSION = -4
ERROR_FILE_CORRUPT = -5
ERROR_MEMORY_ALLOCATION_FAILED = -6
ERROR_NAMELIST_FULL = -7
ERROR_FILE_MUST_BE_WRITABLE = -8
ERROR_FILE_MUST_BE_READABLE = -9

OFFSET_AUTO = 2 ** 64 - 1
N_GLOBAL_AUTO = 2 ** 64 - 1


class Header(C.Structure):  # ref: pgsd.h:143-174
    _fields_ = [
        ("magic", C.c_uint64),
        ("index_location", C.c_uint64),
        ("index_allocated_entries", C.c_uint64),
        ("namelist_location", C.c_uint64),
        ("namelist_allocated_entries", C.c_uint64),
        ("schema_version", C.c_uint32),
        ("pgsd_version", C.c_uint32),
        ("application", C.c_char * PGSD_NAME_SIZE),
        ("schema", C.c_char * PGSD_NAME_SIZE),
        ("reserved", C.c_char * PGSD_RESERVED_BYTES),
    ]


class IndexEntry(C.Structure):  # ref: pgsd.h:182-204
    _fields_ = [
        ("frame", C.c_uint64),
        ("N", C.c_uint64),
        ("location", C.c_int64),
        ("M", C.c_uint32),
        ("id", C.c_uint16),
        ("type", C.c_uint8),
        ("flags", C.c_uint8),
    ]


class IndexBuffer(C.Structure):
    _fields_ = [("data", C.POINTER(IndexEntry)), ("size", C.c_size_t), ("reserved", C.c_size_t),
                ("mapped_data", C.c_void_p), ("mapped_len", C.c_size_t)]


class ByteBuffer(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t), ("reserved", C.c_size_t)]


class NameBuffer(C.Structure):
    _fields_ = [("data", ByteBuffer), ("n_names", C.c_size_t)]


class NameIdMap(C.Structure):
    _fields_ = [("v", C.c_void_p), ("size", C.c_size_t)]


class Handle(C.Structure):  # ref: pgsd.h:297-353 (544 bytes on LP64)
    _fields_ = [
        ("fh", C.c_void_p),
        ("header", Header),
        ("file_index", IndexBuffer),
        ("frame_index", IndexBuffer),
        ("buffer_index", IndexBuffer),
        ("write_buffer", ByteBuffer),
        ("file_names", NameBuffer),
        ("frame_names", NameBuffer),
        ("cur_frame", C.c_uint64),
        ("file_size", C.c_longlong),
        ("open_flags", C.c_int),
        ("name_map", NameIdMap),
        ("pending_index_entries", C.c_uint64),
        ("maximum_write_buffer_size", C.c_uint64),
        ("index_entries_to_buffer", C.c_uint64),
        ("rank", C.c_int),
        ("nprocs", C.c_int),
    ]


assert C.sizeof(Header) == 256 and C.sizeof(IndexEntry) == 32 and C.sizeof(Handle) == 544


class Column(C.Structure):  # struct pgsd_b200_column
    _fields_ = [("base", C.c_void_p), ("stride", C.c_int64)]


class ChunkDesc(C.Structure):  # struct pgsd_b200_chunk_desc
    _fields_ = [("name", C.c_char_p), ("dst_type", C.c_int), ("src_type", C.c_int), ("N", C.c_uint64),
                ("M", C.c_uint32), ("N_global", C.c_uint64), ("M_global", C.c_uint32), ("offset", C.c_uint64),
                ("all", C.c_bool), ("cols", C.POINTER(Column))]


class Field(C.Structure):  # struct pgsd_b200_field
    _fields_ = [("in_", C.c_void_p), ("out", C.c_void_p), ("row_bytes", C.c_uint32)]


class Stats(C.Structure):  # struct pgsd_b200_stats
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("file_bytes_written", C.c_uint64), ("file_bytes_read", C.c_uint64),
                ("collectives", C.c_uint64), ("commit_wait_s", C.c_double), ("d2h_busy_s", C.c_double),
                ("file_busy_s", C.c_double), ("pieces", C.c_uint64)]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_size_t)

_HP = C.POINTER(Handle)
_u64, _u32, _i = C.c_uint64, C.c_uint32, C.c_int
_vp = C.c_void_p

# name -> (restype, argtypes): every symbol include/pgsd.h and include/pgsd_b200.h declare
SIGNATURES = {
    # ---- include/pgsd.h
    "pgsd_make_version": (C.c_uint32, [C.c_uint, C.c_uint]),
    "pgsd_create_and_open": (_i, [_HP, C.c_char_p, C.c_char_p, C.c_char_p, _u32, _i, _i]),
    "pgsd_open": (_i, [_HP, C.c_char_p, _i]),
    "pgsd_close": (_i, [_HP]),
    "pgsd_end_frame": (_i, [_HP]),
    "pgsd_flush": (_i, [_HP]),
    "pgsd_write_chunk": (_i, [_HP, C.c_char_p, _i, _u64, _u32, _u64, _u32, _u64, _u64, C.c_bool, C.c_uint8, _vp]),
    "pgsd_find_chunk": (C.POINTER(IndexEntry), [_HP, _u64, C.c_char_p]),
    "pgsd_read_chunk": (_i, [_HP, _vp, C.POINTER(IndexEntry), _u64, _u32, _u32, C.c_bool]),
    "pgsd_get_nframes": (_u64, [_HP]),
    "pgsd_get_nnames": (_u64, [_HP]),
    "pgsd_sizeof_type": (C.c_size_t, [_i]),
    "pgsd_find_matching_chunk_name": (_vp, [_HP, C.c_char_p, _vp]),
    "pgsd_get_maximum_write_buffer_size": (_u64, [_HP]),
    "pgsd_set_maximum_write_buffer_size": (_i, [_HP, _u64]),
    "pgsd_get_index_entries_to_buffer": (_u64, [_HP]),
    "pgsd_set_index_entries_to_buffer": (_i, [_HP, _u64]),
    "pgsd_bcast_index_entry": (None, [C.POINTER(IndexEntry)]),
    "is_root": (C.c_bool, []),
    # ---- include/pgsd_b200.h
    "pgsd_b200_comm_init_host": (_i, [_i, _i, ALLGATHER_FN, _vp]),
    "pgsd_b200_comm_init_shm": (_i, [_i, _i, C.c_char_p]),
    "pgsd_b200_nccl_unique_id": (_i, [_vp]),
    "pgsd_b200_comm_init_nccl": (_i, [_i, _i, _vp, _i]),
    "pgsd_b200_comm_finalize": (_i, []),
    "pgsd_b200_comm_rank": (_i, []),
    "pgsd_b200_comm_size": (_i, []),
    "pgsd_b200_comm_kind": (C.c_char_p, []),
    "pgsd_b200_barrier": (_i, []),
    "pgsd_b200_partition": (_i, [_u64, C.POINTER(_u64), C.POINTER(_u64)]),
    "pgsd_b200_cuda_available": (_i, []),
    "pgsd_b200_device_init": (_i, [_i]),
    "pgsd_b200_set_stream": (_i, [_vp]),
    "pgsd_b200_last_error": (C.c_char_p, []),
    "pgsd_b200_configure_staging": (_i, [_u32, _u64, _u32]),
    "pgsd_b200_write_chunk_soa": (_i, [_HP, C.c_char_p, _i, _u64, _u32, _u64, _u32, _u64, C.c_bool, _i, C.POINTER(Column)]),
    "pgsd_b200_write_chunks_soa": (_i, [_HP, _i, C.POINTER(ChunkDesc)]),
    "pgsd_b200_pack_soa": (_i, [_vp, _i, _u64, _u32, _i, C.POINTER(Column), _vp]),
    "pgsd_b200_scan_sizes": (_i, [C.POINTER(_u64), _i, _i, _i, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "pgsd_b200_sort_ids": (_i, [_u64, _vp, _vp, _vp, _vp]),
    "pgsd_b200_gather": (_i, [_u64, _vp, _i, C.POINTER(Field), _vp]),
    "pgsd_b200_reorder_device": (_i, [_u64, _vp, _vp, _vp, _i, C.POINTER(Field), _vp]),
    "pgsd_b200_reorder_distributed_plan": (_i, [_u64, _i, _i, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pgsd_b200_reorder_distributed": (_i, [_u64, _vp, _u64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), _vp, _i,
                                            C.POINTER(Field), _vp]),
    "pgsd_b200_reorder_host": (_i, [_u64, _vp, _vp, _vp, _i, C.POINTER(Field)]),
    "pgsd_b200_get_stats": (_i, [C.POINTER(Stats)]),
    "pgsd_b200_reset_stats": (_i, []),
    "pgsd_b200_malloc": (_i, [C.POINTER(_vp), _u64]),
    "pgsd_b200_free": (_i, [_vp]),
    "pgsd_b200_host_alloc": (_i, [C.POINTER(_vp), _u64]),
    "pgsd_b200_host_free": (_i, [_vp]),
    "pgsd_b200_memcpy": (_i, [_vp, _vp, _u64, _i]),
    "pgsd_b200_synchronize": (_i, []),
    "pgsd_b200_drain": (_i, []),
    "pgsd_b200_shutdown": (_i, []),
    "pgsd_b200_timer_create": (_i, [C.POINTER(_vp)]),
    "pgsd_b200_timer_start": (_i, [_vp]),
    "pgsd_b200_timer_stop": (_i, [_vp, C.POINTER(C.c_float)]),
    "pgsd_b200_timer_destroy": (_i, [_vp]),
    "pgsd_b200_flush_l2": (_i, []),
    "pgsd_b200_pack_profiling": (_i, [_i]),
    "pgsd_b200_pack_last_ms": (_i, [C.POINTER(C.c_float)]),
    "pgsd_b200_reorder_profiling": (_i, [_i]),
    "pgsd_b200_reorder_phase_ms": (_i, [C.POINTER(C.c_float)]),
    "pgsd_b200_selftest": (_i, [_i]),
    "pgsd_b200_file_stage_write": (_i, [_i, _vp, _u64, _u64, _i]),
    "pgsd_b200_read_ahead_stats": (_i, [C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "pgsd_b200_read_ahead_host_read": (_i, [_i, _vp, _u64, _u64]),
    "pgsd_b200_read_ahead_host_reset": (_i, []),
    "pgsd_b200_read_ahead_host_stats": (_i, [C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "pgsd_b200_file_stage_ceiling": (_i, [C.c_char_p, _u64, _u64, C.POINTER(C.c_double), C.POINTER(_i), C.POINTER(_i)]),
}

_lib = None


def load():
    """Return the loaded library; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C pgsd_sph_b200/csrc`. pgsd_sph_b200 has no pure-Python or CPU fallback.")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL, use_errno=True)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    msg = load().pgsd_b200_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what=""):
    """Raise for a negative return code of a pgsd_b200_* device/communicator call."""
    if rc != 0:
        raise RuntimeError(f"{what} failed with code {rc}: {last_error()}")
