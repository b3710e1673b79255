# cython: language_level=3, embedsignature=True
"""PGSD file layer API -- drop-in for the reference's ``pgsd.fl`` over libpgsd_b200.

Cython binding of the C ABI (include/pgsd.h, include/pgsd_b200.h), same surface as the reference's
Cython module (/root/reference/pgsd/pgsd/fl.pyx:149-1052): :py:func:`open`, :py:class:`PGSDFile` with
``write_chunk / end_frame / flush / read_chunk / chunk_exists / find_matching_chunk_names / close``,
the same properties, error mapping (fl.pyx:35-61) and context-manager / pickle behaviour.  The C calls
run ``with nogil`` like the reference's.  What is new:

* ``data`` of :py:meth:`PGSDFile.write_chunk` may live on the GPU (anything exposing
  ``__cuda_array_interface__`` or ``__dlpack__``): contiguous arrays are handed to the C ABI as
  device pointers; strided ones are made contiguous by the K1 pack kernel on the device -- the
  device-side ``numpy.ascontiguousarray`` of fl.pyx:571.  No PyTorch dependency.
* :py:meth:`PGSDFile.write_chunk_soa` / :py:meth:`PGSDFile.write_frame_soa` pack + dtype-cast separate
  component arrays into (N, M) chunks on the device (the ``ParticleData.validate`` contract,
  hoomd.py:206-270), a whole frame with one kernel launch.
* ``offset='auto'`` lets the library place this rank's rows at the exclusive prefix over ranks
  (K2) instead of the caller passing all ranks' row counts (fl.pyx:596-598).
* :py:meth:`PGSDFile.read_chunk` can read straight into device memory (``device=True``).
"""
import errno as _errno
import logging
import os
from pickle import PickleError

import numpy

from libc.errno cimport errno
from libc.stdint cimport uint8_t, uint32_t, uint64_t, int64_t, uintptr_t
from libc.stdlib cimport malloc, free
from libc.string cimport memset
from cpython.buffer cimport PyObject_GetBuffer, PyBuffer_Release, PyBUF_SIMPLE, PyBUF_WRITABLE

cimport libpgsd

from .devmem import DeviceArray, as_device_view, is_device_array

logger = logging.getLogger('pgsd.fl')

_NP_TO_PGSD = {
    numpy.dtype(numpy.uint8): libpgsd.PGSD_TYPE_UINT8, numpy.dtype(numpy.uint16): libpgsd.PGSD_TYPE_UINT16,
    numpy.dtype(numpy.uint32): libpgsd.PGSD_TYPE_UINT32, numpy.dtype(numpy.uint64): libpgsd.PGSD_TYPE_UINT64,
    numpy.dtype(numpy.int8): libpgsd.PGSD_TYPE_INT8, numpy.dtype(numpy.int16): libpgsd.PGSD_TYPE_INT16,
    numpy.dtype(numpy.int32): libpgsd.PGSD_TYPE_INT32, numpy.dtype(numpy.int64): libpgsd.PGSD_TYPE_INT64,
    numpy.dtype(numpy.float32): libpgsd.PGSD_TYPE_FLOAT, numpy.dtype(numpy.float64): libpgsd.PGSD_TYPE_DOUBLE,
}
_NP_TO_PGSD = {k: int(v) for k, v in _NP_TO_PGSD.items()}
_PGSD_TO_NP = {v: k for k, v in _NP_TO_PGSD.items()}

cdef uint64_t _AUTO = 0xFFFFFFFFFFFFFFFFULL  # PGSD_B200_OFFSET_AUTO / N_global "auto"


cdef str _last_error():
    cdef const char* m = libpgsd.pgsd_b200_last_error()
    return m.decode('utf-8', 'replace') if m != NULL else ''


cdef _raise_on_error(int retval, extra, int err=0):
    """Raise the appropriate error type (ref: fl.pyx:35-61)."""
    if retval == 0:
        return
    if retval == libpgsd.PGSD_ERROR_IO:
        e = err or _errno.EIO
        raise IOError(e, os.strerror(e), extra)
    elif retval == libpgsd.PGSD_ERROR_NOT_A_PGSD_FILE:
        raise RuntimeError("Not a PGSD file: " + extra)
    elif retval == libpgsd.PGSD_ERROR_INVALID_PGSD_FILE_VERSION:
        raise RuntimeError("Unsupported PGSD file version: " + extra)
    elif retval == libpgsd.PGSD_ERROR_FILE_CORRUPT:
        raise RuntimeError("Corrupt PGSD file: " + extra)
    elif retval == libpgsd.PGSD_ERROR_MEMORY_ALLOCATION_FAILED:
        raise MemoryError("Memory allocation failed: " + extra)
    elif retval == libpgsd.PGSD_ERROR_NAMELIST_FULL:
        raise RuntimeError("PGSD namelist is full: " + extra)
    elif retval == libpgsd.PGSD_ERROR_FILE_MUST_BE_WRITABLE:
        raise RuntimeError("File must be writable: " + extra)
    elif retval == libpgsd.PGSD_ERROR_FILE_MUST_BE_READABLE:
        raise RuntimeError("File must be readable: " + extra)
    elif retval == libpgsd.PGSD_ERROR_INVALID_ARGUMENT:
        raise RuntimeError("Invalid pgsd argument: " + extra + " " + _last_error())
    else:
        raise RuntimeError("Unknown error: " + extra + " " + _last_error())


def open(name, mode, application=None, schema=None, schema_version=None):
    """Open a PGSD file and return a :py:class:`PGSDFile` (ref: fl.pyx:149-228).

    Modes: ``'r'`` read, ``'r+'`` read/write existing, ``'w'`` create/overwrite, ``'x'`` create
    exclusively, ``'a'`` read/write, created if missing.
    """
    return PGSDFile(str(name), mode, application, schema, schema_version)


cdef class _PreparedFrame:
    """C descriptor table of a frame's SoA chunks (see PGSDFile.prepare_frame_soa)."""
    cdef libpgsd.pgsd_b200_chunk_desc* descs
    cdef libpgsd.pgsd_b200_column* cols
    cdef int n
    cdef object keep

    def __cinit__(self):
        self.descs = NULL
        self.cols = NULL
        self.n = 0

    def __dealloc__(self):
        if self.descs != NULL:
            free(self.descs)
        if self.cols != NULL:
            free(self.cols)


cdef struct _HostChunk:
    const char* name
    int pgsd_type
    uint64_t N
    uint32_t M
    uint64_t N_global
    uint64_t stride
    bint write_all
    const void* data


cdef class _PreparedHost:
    """Argument table of host chunks written every frame from the same buffers (see PGSDFile.prepare_chunks)."""
    cdef _HostChunk* chunks
    cdef int n
    cdef object keep

    def __cinit__(self):
        self.chunks = NULL
        self.n = 0

    def __dealloc__(self):
        if self.chunks != NULL:
            free(self.chunks)


cdef class PGSDFile:
    """PGSD file access interface (ref: fl.pyx:231-380)."""

    cdef libpgsd.pgsd_handle _handle   # caller-owned handle, embedded by value (ref: fl.pyx:284)
    cdef bint _is_open
    cdef str _mode
    cdef str _name

    def __init__(self, name, mode, application, schema, schema_version):
        cdef libpgsd.pgsd_open_flag c_flags
        cdef int exclusive_create = 0
        cdef int overwrite = 0
        cdef int retval = 0
        cdef int err = 0
        cdef uint32_t version
        cdef bytes bname, bapp, bschema
        cdef const char* c_name
        cdef const char* c_app
        cdef const char* c_schema
        self._is_open = False
        self._mode = mode
        if mode == 'w':
            c_flags = libpgsd.PGSD_OPEN_READWRITE
            overwrite = 1
        elif mode == 'r':
            c_flags = libpgsd.PGSD_OPEN_READONLY
        elif mode == 'r+':
            c_flags = libpgsd.PGSD_OPEN_READWRITE
        elif mode == 'x':
            c_flags = libpgsd.PGSD_OPEN_READWRITE
            overwrite = 1
            exclusive_create = 1
        elif mode == 'a':
            c_flags = libpgsd.PGSD_OPEN_READWRITE
            if not os.path.exists(name):
                overwrite = 1
        else:
            raise ValueError("Invalid mode: " + mode)
        self._name = name
        bname = name.encode('utf-8')
        c_name = bname

        if overwrite:
            if application is None:
                raise ValueError("Provide application when creating a file")
            if schema is None:
                raise ValueError("Provide schema when creating a file")
            if schema_version is None:
                raise ValueError("Provide schema_version when creating a file")
            logger.info('overwriting file: ' + name + ' with mode: ' + mode + ', application: ' + application
                        + ', schema: ' + schema + ', and schema_version: ' + str(schema_version))
            if exclusive_create and os.path.exists(name) and libpgsd.pgsd_b200_comm_size() == 1:
                raise FileExistsError(_errno.EEXIST, os.strerror(_errno.EEXIST), name)
            version = libpgsd.pgsd_make_version(int(schema_version[0]), int(schema_version[1]))
            bapp = application.encode('utf-8')
            bschema = schema.encode('utf-8')
            c_app = bapp
            c_schema = bschema
            with nogil:
                retval = libpgsd.pgsd_create_and_open(&self._handle, c_name, c_app, c_schema, version, c_flags,
                                                      exclusive_create)
                err = errno
        else:
            logger.info('opening file: ' + name + ' with mode: ' + mode)
            if not os.path.exists(name):
                raise FileNotFoundError(_errno.ENOENT, os.strerror(_errno.ENOENT), name)
            with nogil:
                retval = libpgsd.pgsd_open(&self._handle, c_name, c_flags)
                err = errno
        _raise_on_error(retval, name, err)
        self._is_open = True

        if schema is not None:
            schema_truncated = schema
            if len(schema_truncated) > 64:
                schema_truncated = schema_truncated[0:63]
            if self.schema != schema_truncated:
                file_schema = self.schema
                self.close()
                raise RuntimeError('file ' + name + ' has incorrect schema: ' + file_schema)

    # ------------------------------------------------------------------ life cycle
    def close(self, write_all=True):
        """Close the file; further operations raise ValueError (ref: fl.pyx:382-458)."""
        cdef int retval, err
        if self._is_open:
            logger.info('closing file: ' + self._name)
            with nogil:
                retval = libpgsd.pgsd_close(&self._handle)
                err = errno
            self._is_open = False
            _raise_on_error(retval, self._name, err)

    def end_frame(self, write_all=True):
        """Complete the current frame (ref: fl.pyx:460-505).

        With several ranks this is where the frame's chunks get their file offsets: one
        all-gather of the chunk sizes + exclusive scan (K2) instead of per-chunk collectives.
        """
        cdef int retval, err
        self._check_open()
        with nogil:
            retval = libpgsd.pgsd_end_frame(&self._handle)
            err = errno
        _raise_on_error(retval, self._name, err)

    def flush(self, write_all=True):
        """Flush all buffered frames to the file and wait for queued device writes (ref: fl.pyx:507-524)."""
        cdef int retval, err
        self._check_open()
        with nogil:
            retval = libpgsd.pgsd_flush(&self._handle)
            err = errno
        _raise_on_error(retval, self._name, err)

    # ------------------------------------------------------------------ write
    cdef _offset_args(self, N, M, offset, rank):
        # ref: fl.pyx:594-598 -- `offset` holds the row counts of all ranks
        if offset is None:
            return N, 0
        if isinstance(offset, str):
            if offset != 'auto':
                raise ValueError("offset must be None, 'auto' or an array of per-rank row counts")
            return _AUTO, _AUTO
        offset = numpy.asarray(offset)
        return int(offset.sum()), int(M) * int(offset[0:rank].sum())

    cdef int _c_write(self, bytes bname, int pgsd_type, uint64_t N, uint32_t M, uint64_t N_global, uint64_t stride,
                      bint write_all, uintptr_t ptr, int* err) noexcept:
        cdef const char* c_name = bname
        cdef uint64_t gsize = 0 if N_global == _AUTO else N_global * M
        cdef int retval
        with nogil:
            retval = libpgsd.pgsd_write_chunk(&self._handle, c_name, <libpgsd.pgsd_type>pgsd_type, N, M, N_global, M,
                                              stride, gsize, write_all, 0, <const void*>ptr)
            err[0] = errno
        return retval

    def write_chunk(self, name, data, offset=None, rank=0, write_all=True):
        """Write a data chunk to the current frame (ref: fl.pyx:526-654).

        Args:
            name (str): Name of the chunk.
            data: numpy array / array-like, or a CUDA array (``__cuda_array_interface__`` /
                ``__dlpack__``), with 2 or fewer dimensions.
            offset: per-rank row counts (N_global = offset.sum(), this rank starts at
                ``offset[:rank].sum()`` rows), ``'auto'`` (library computes both), or None.
            rank (int): this rank's index into ``offset``.
            write_all (bool): every rank writes its rows (True, the reference default) or the
                chunk is replicated/small and goes through the write buffer (False).
        """
        cdef Py_buffer buf
        cdef int retval, err = 0
        cdef uintptr_t ptr = 0
        self._check_open()
        if is_device_array(data):
            return self._write_chunk_device(name, data, offset, rank, write_all)

        data_array = numpy.ascontiguousarray(data)
        if data_array is not data:
            logger.warning('implicit data copy when writing chunk: ' + name)
        data_array = data_array.view()
        if len(data_array.shape) > 2:
            raise ValueError("PGSD can only write 1 or 2 dimensional arrays: " + name)
        if len(data_array.shape) == 1:
            data_array = data_array.reshape([data_array.shape[0], 1])
        N, M = data_array.shape
        N_global, stride = self._offset_args(N, M, offset, rank)
        pgsd_type = _NP_TO_PGSD.get(data_array.dtype)
        if pgsd_type is None:
            raise ValueError("invalid type for chunk: " + name)
        if data_array.size:
            PyObject_GetBuffer(data_array, &buf, PyBUF_SIMPLE)
            ptr = <uintptr_t>buf.buf
            try:
                retval = self._c_write(name.encode('utf-8'), pgsd_type, N, M, N_global, stride, bool(write_all), ptr, &err)
            finally:
                PyBuffer_Release(&buf)
        else:
            retval = self._c_write(name.encode('utf-8'), pgsd_type, N, M, N_global, stride, bool(write_all), 0, &err)
        _raise_on_error(retval, self._name, err)

    def prepare_chunks(self, chunks, rank=0):
        """Validate a list of host chunks once and keep their ``pgsd_write_chunk`` arguments.

        ``chunks``: sequence of ``(name, array, offset, write_all)`` with the meaning of :py:meth:`write_chunk`'s
        arguments; the arrays must be C-contiguous numpy arrays with 1 or 2 dimensions.  The returned object keeps
        the arrays alive; :py:meth:`write_prepared` writes their CURRENT contents, so a simulation that logs the
        same scalars every step updates the arrays in place and pays the Python argument handling once instead of
        once per chunk and frame (3.5 us each, which is most of a 4096-particle frame's cost).
        """
        cdef _PreparedHost ph = _PreparedHost()
        cdef int n = len(chunks), i
        keep = []
        ph.chunks = <_HostChunk*>malloc(max(n, 1) * sizeof(_HostChunk))
        if ph.chunks == NULL:
            raise MemoryError()
        i = 0
        for (name, data, offset, write_all) in chunks:
            if not isinstance(data, numpy.ndarray) or not data.flags['C_CONTIGUOUS']:
                raise ValueError("prepare_chunks needs C-contiguous numpy arrays: " + name)
            if data.ndim > 2:
                raise ValueError("PGSD can only write 1 or 2 dimensional arrays: " + name)
            N = data.shape[0] if data.ndim >= 1 else 1
            M = data.shape[1] if data.ndim == 2 else 1
            N_global, stride = self._offset_args(N, M, offset, rank)
            pgsd_type = _NP_TO_PGSD.get(data.dtype)
            if pgsd_type is None:
                raise ValueError("invalid type for chunk: " + name)
            bname = name.encode('utf-8')
            keep.extend([bname, data])
            ph.chunks[i].name = <const char*>bname
            ph.chunks[i].pgsd_type = pgsd_type
            ph.chunks[i].N = N
            ph.chunks[i].M = M
            ph.chunks[i].N_global = N_global
            ph.chunks[i].stride = stride
            ph.chunks[i].write_all = bool(write_all)
            ph.chunks[i].data = <const void*><uintptr_t>(data.ctypes.data if data.size else 0)
            i += 1
        ph.n = n
        ph.keep = keep
        return ph

    def write_prepared(self, _PreparedHost prepared):
        """``pgsd_write_chunk`` for every chunk of ``prepared`` (:py:meth:`prepare_chunks`), in order, from the arrays'
        current contents.  Equivalent to calling :py:meth:`write_chunk` for each."""
        cdef int retval = 0, err = 0, i
        cdef _HostChunk* c
        cdef uint64_t gsize
        self._check_open()
        with nogil:
            for i in range(prepared.n):
                c = &prepared.chunks[i]
                gsize = 0 if c.N_global == _AUTO else c.N_global * c.M
                retval = libpgsd.pgsd_write_chunk(&self._handle, c.name, <libpgsd.pgsd_type>c.pgsd_type, c.N, c.M, c.N_global,
                                                  c.M, c.stride, gsize, c.write_all, 0, c.data)
                if retval != 0:
                    err = errno
                    break
        _raise_on_error(retval, self._name, err)

    def _write_chunk_device(self, name, data, offset, rank, write_all):
        cdef libpgsd.pgsd_b200_column cols[8]
        cdef int retval, err = 0, j
        cdef const char* c_name
        cdef uint64_t cN, cNg, cstride
        cdef uint32_t cM
        cdef int ctype
        cdef bint call
        ptr, shape, dtype, strides, keep = as_device_view(data)
        if len(shape) > 2:
            raise ValueError("PGSD can only write 1 or 2 dimensional arrays: " + name)
        pgsd_type = _NP_TO_PGSD.get(dtype)
        if pgsd_type is None:
            raise ValueError("invalid type for chunk: " + name)
        N = int(shape[0]) if len(shape) else 1
        M = int(shape[1]) if len(shape) == 2 else 1
        N_global, stride = self._offset_args(N, M, offset, rank)
        item = dtype.itemsize
        contiguous = strides is None or tuple(strides) == ((M * item, item) if len(shape) == 2 else (item,))
        bname = name.encode('utf-8')
        if contiguous or N == 0:
            retval = self._c_write(bname, pgsd_type, N, M, N_global, stride, bool(write_all), <uintptr_t>(ptr if N else 0),
                                   &err)
        else:
            # strided device array: K1 makes it contiguous (device-side ascontiguousarray)
            if M > 8 or any(s % item for s in strides):
                raise ValueError("strided device arrays need M <= 8 and element-aligned strides: " + name)
            logger.warning('implicit device pack when writing chunk: ' + name)
            col_stride = strides[1] if len(shape) == 2 else item
            for j in range(M):
                cols[j].base = <const void*><uintptr_t>(ptr + j * col_stride)
                cols[j].stride = strides[0] // item
            c_name = bname
            cN, cM, cNg, cstride, ctype, call = N, M, N_global, stride, pgsd_type, bool(write_all)
            with nogil:
                retval = libpgsd.pgsd_b200_write_chunk_soa(&self._handle, c_name, <libpgsd.pgsd_type>ctype, cN, cM, cNg, cM,
                                                           cstride, call, <libpgsd.pgsd_type>ctype, cols)
                err = errno
        del keep
        _raise_on_error(retval, self._name, err)

    def _soa_views(self, name, columns):
        """-> (M x (ptr, N, dtype, stride, on_device), N, src dtype, keep-alive list)."""
        M = len(columns)
        if M < 1 or M > 8:
            raise ValueError("write_chunk_soa takes 1..8 columns: " + name)
        views, keep = [], []
        for c in columns:
            if is_device_array(c):
                ptr, shape, cdt, strides, k = as_device_view(c)
                keep.append(k)
                on_device = True
            else:
                a = numpy.asarray(c)
                ptr, shape, cdt, strides = a.ctypes.data, a.shape, a.dtype, a.strides
                keep.append(a)
                on_device = False
            if len(shape) != 1:
                raise ValueError("write_chunk_soa columns must be 1-dimensional: " + name)
            st = cdt.itemsize if strides is None else strides[0]
            if st % cdt.itemsize:
                raise ValueError("column stride is not a multiple of the item size: " + name)
            views.append((ptr, int(shape[0]), cdt, st // cdt.itemsize, on_device))
        N, src_dt = views[0][1], views[0][2]
        if any(v[1] != N or v[2] != src_dt or v[4] != views[0][4] for v in views):
            raise ValueError("write_chunk_soa columns must share length, dtype and memory space: " + name)
        return views, N, src_dt, keep

    def write_chunk_soa(self, name, columns, dtype=None, offset=None, rank=0, write_all=True):
        """Pack M component arrays into one (N, M) chunk on the device and write it (K1).

        ``columns`` is a sequence of M equally long 1-D arrays of one dtype -- CUDA arrays (hot
        path) or numpy arrays (uploaded first).  ``dtype`` is the chunk's dtype (default: the
        columns' dtype); the cast follows ``numpy.astype``.  This is the device form of
        ``numpy.ascontiguousarray(numpy.stack(columns, 1), dtype)`` -- what the reference's
        callers do on the host before ``write_chunk`` (fl.pyx:571, hoomd.py:206-270).
        """
        self._check_open()
        prepared = self.prepare_frame_soa([(name, columns, dtype, offset, write_all)], rank=rank)
        self.write_frame_soa(prepared)

    def prepare_frame_soa(self, chunks, rank=0):
        """Build the reusable C descriptor table for :py:meth:`write_frame_soa`.

        ``chunks``: sequence of ``(name, columns, dtype, offset, write_all)`` with the meaning of
        :py:meth:`write_chunk_soa`'s arguments.  The returned object keeps the arrays alive and can
        be written any number of times (one simulation's buffers, one frame per time step).
        """
        cdef _PreparedFrame pf = _PreparedFrame()
        cdef int n = len(chunks), i, j, ncols = 0, c0 = 0
        keep = []
        parsed = []
        for (name, columns, dtype, offset, write_all) in chunks:
            views, N, src_dt, k = self._soa_views(name, columns)
            src_type = _NP_TO_PGSD.get(src_dt)
            dst_type = _NP_TO_PGSD.get(numpy.dtype(dtype) if dtype is not None else src_dt)
            if src_type is None or dst_type is None:
                raise ValueError("invalid type for chunk: " + name)
            N_global, stride = self._offset_args(N, len(views), offset, rank)
            bname = name.encode('utf-8')
            keep.extend([k, bname])
            parsed.append((bname, views, N, src_type, dst_type, N_global, stride, bool(write_all)))
            ncols += len(views)
        pf.descs = <libpgsd.pgsd_b200_chunk_desc*>malloc(max(n, 1) * sizeof(libpgsd.pgsd_b200_chunk_desc))
        pf.cols = <libpgsd.pgsd_b200_column*>malloc(max(ncols, 1) * sizeof(libpgsd.pgsd_b200_column))
        if pf.descs == NULL or pf.cols == NULL:
            raise MemoryError()
        memset(pf.descs, 0, max(n, 1) * sizeof(libpgsd.pgsd_b200_chunk_desc))
        for i in range(n):
            bname, views, N, src_type, dst_type, N_global, stride, wa = parsed[i]
            for j in range(len(views)):
                pf.cols[c0 + j].base = <const void*><uintptr_t>(views[j][0] if N else 0)
                pf.cols[c0 + j].stride = views[j][3]
            pf.descs[i].name = <const char*>bname
            pf.descs[i].dst_type = <libpgsd.pgsd_type><int>dst_type
            pf.descs[i].src_type = <libpgsd.pgsd_type><int>src_type
            pf.descs[i].N = N
            pf.descs[i].M = len(views)
            pf.descs[i].N_global = N_global
            pf.descs[i].M_global = len(views)
            pf.descs[i].offset = stride
            pf.descs[i].all = wa
            pf.descs[i].cols = pf.cols + c0
            c0 += len(views)
        pf.n = n
        pf.keep = keep
        return pf

    def write_frame_soa(self, _PreparedFrame prepared):
        """Write all SoA chunks of a frame with ONE K1 launch (``pgsd_b200_write_chunks_soa``);
        ``prepared`` comes from :py:meth:`prepare_frame_soa`.  Equivalent to calling
        :py:meth:`write_chunk_soa` for each chunk in order."""
        cdef int retval, err
        self._check_open()
        with nogil:
            retval = libpgsd.pgsd_b200_write_chunks_soa(&self._handle, prepared.n, prepared.descs)
            err = errno
        _raise_on_error(retval, self._name, err)

    # ------------------------------------------------------------------ read
    def chunk_exists(self, frame, name, write_all=True):
        """Test if a chunk exists (ref: fl.pyx:656-715)."""
        cdef const libpgsd.pgsd_index_entry* entry
        cdef bytes bname = name.encode('utf-8')
        cdef const char* c_name = bname
        cdef uint64_t c_frame = frame
        self._check_open()
        with nogil:
            entry = libpgsd.pgsd_find_chunk(&self._handle, c_frame, c_name)
        return entry != NULL

    def read_chunk(self, frame, name, N=0, M=0, offset=0, r_all=False, device=False):
        """Read a data chunk (ref: fl.pyx:717-874).

        ``r_all=False`` reads the whole (N_global, M) chunk.  ``r_all=True`` reads ``N`` rows of
        ``M`` values starting ``offset`` rows into the chunk and returns exactly those rows
        (the reference returns an (N_global, M) array whose first N rows are filled).
        ``device=True`` returns a :py:class:`~pgsd_sph_b200.devmem.DeviceArray`.
        Raises KeyError if the chunk does not exist.
        """
        cdef const libpgsd.pgsd_index_entry* entry_p
        cdef libpgsd.pgsd_index_entry entry
        cdef bytes bname = name.encode('utf-8')
        cdef const char* c_name = bname
        cdef uint64_t c_frame = frame
        cdef Py_buffer buf
        cdef uintptr_t ptr
        cdef uint64_t c_rows
        cdef uint32_t c_cols, c_off
        cdef bint c_all = bool(r_all)
        cdef int retval = 0, err = 0
        self._check_open()
        with nogil:
            entry_p = libpgsd.pgsd_find_chunk(&self._handle, c_frame, c_name)
        if entry_p == NULL:
            raise KeyError("frame " + str(frame) + " / chunk " + name + " not found in: " + self._name)
        entry = entry_p[0]
        dtype = _PGSD_TO_NP.get(entry.type)
        if dtype is None:
            raise ValueError("invalid type for chunk: " + name)
        rows = int(N) if r_all else int(entry.N)
        cols = int(M) if r_all else int(entry.M)
        if r_all and cols != entry.M:
            raise ValueError("M must equal the chunk's M for a partial read: " + name)
        c_rows, c_cols, c_off = rows, cols, int(offset)
        if device:
            out = DeviceArray((rows, cols), dtype)
            ptr = <uintptr_t>out.ptr
            if rows != 0 and cols != 0:
                with nogil:
                    retval = libpgsd.pgsd_read_chunk(&self._handle, <void*>ptr, &entry, c_rows, c_cols, c_off, c_all)
                    err = errno
        else:
            out = numpy.empty(dtype=dtype, shape=[rows, cols])
            # only read chunk if we have data
            if rows != 0 and cols != 0:
                PyObject_GetBuffer(out, &buf, PyBUF_WRITABLE)
                ptr = <uintptr_t>buf.buf
                try:
                    with nogil:
                        retval = libpgsd.pgsd_read_chunk(&self._handle, <void*>ptr, &entry, c_rows, c_cols, c_off, c_all)
                        err = errno
                finally:
                    PyBuffer_Release(&buf)
        _raise_on_error(retval, self._name, err)
        if entry.M == 1:
            return out.reshape([rows])
        return out

    def find_matching_chunk_names(self, match, write_all=True):
        """All chunk names in the file that start with ``match`` (ref: fl.pyx:876-945)."""
        cdef bytes bmatch = match.encode('utf-8')
        cdef const char* c_match = bmatch
        cdef const char* found
        self._check_open()
        retval = []
        with nogil:
            found = libpgsd.pgsd_find_matching_chunk_name(&self._handle, c_match, NULL)
        while found != NULL:
            retval.append(found.decode('utf-8'))
            with nogil:
                found = libpgsd.pgsd_find_matching_chunk_name(&self._handle, c_match, found)
        return retval

    # ------------------------------------------------------------------ protocol / properties
    def _check_open(self):
        if not self._is_open:
            raise ValueError("File is not open")

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.close()

    def __reduce__(self):
        """Allows read-only files to be pickled (ref: fl.pyx:968-977)."""
        if self._mode not in ['rb', 'r']:
            raise PickleError("Only read only GSDFiles can be pickled.")
        return (PGSDFile, (self._name, self._mode, self.application, self.schema, self.schema_version))

    @property
    def name(self):
        return self._name

    @property
    def mode(self):
        return self._mode

    @property
    def pgsd_version(self):
        cdef uint32_t v = self._handle.header.pgsd_version
        return (v >> 16, v & 0xffff)

    @property
    def schema_version(self):
        cdef uint32_t v = self._handle.header.schema_version
        return (v >> 16, v & 0xffff)

    @property
    def schema(self):
        return self._handle.header.schema.decode('utf-8')

    @property
    def application(self):
        return self._handle.header.application.decode('utf-8')

    @property
    def nframes(self):
        self._check_open()
        return libpgsd.pgsd_get_nframes(&self._handle)

    @property
    def nnames(self):
        self._check_open()
        return libpgsd.pgsd_get_nnames(&self._handle)

    @property
    def maximum_write_buffer_size(self):
        self._check_open()
        return libpgsd.pgsd_get_maximum_write_buffer_size(&self._handle)

    @maximum_write_buffer_size.setter
    def maximum_write_buffer_size(self, size):
        self._check_open()
        _raise_on_error(libpgsd.pgsd_set_maximum_write_buffer_size(&self._handle, int(size)), self._name)

    @property
    def index_entries_to_buffer(self):
        self._check_open()
        return libpgsd.pgsd_get_index_entries_to_buffer(&self._handle)

    @index_entries_to_buffer.setter
    def index_entries_to_buffer(self, number):
        self._check_open()
        _raise_on_error(libpgsd.pgsd_set_index_entries_to_buffer(&self._handle, int(number)), self._name)

    def __dealloc__(self):
        if self._is_open:
            libpgsd.pgsd_close(&self._handle)
            self._is_open = False
