"""PGSD file layer API -- drop-in for the reference's ``pgsd.fl`` over libpgsd_b200.

Same surface as the reference Cython module (/root/reference/pgsd/pgsd/fl.pyx:149-1052):
:py:func:`open`, :py:class:`PGSDFile` with ``write_chunk / end_frame / flush / read_chunk /
chunk_exists / find_matching_chunk_names / close``, the same properties, error mapping
(fl.pyx:35-61) and context-manager / pickle behaviour.  What is new:

* ``data`` of :py:meth:`PGSDFile.write_chunk` may live on the GPU (anything exposing
  ``__cuda_array_interface__`` or ``__dlpack__``): contiguous arrays are handed to the C ABI as
  device pointers; strided ones are made contiguous by the K1 pack kernel on the device -- the
  device-side ``numpy.ascontiguousarray`` of fl.pyx:571.
* :py:meth:`PGSDFile.write_chunk_soa` packs + dtype-casts M separate component arrays into one
  (N, M) chunk on the device (the ``ParticleData.validate`` contract, hoomd.py:206-270).
* ``offset='auto'`` lets the library place this rank's rows at the exclusive prefix over ranks
  (K2) instead of the caller passing all ranks' row counts (fl.pyx:596-598).
* :py:meth:`PGSDFile.read_chunk` can read straight into device memory (``device=True``).

The calls release the GIL (ctypes foreign calls do, like the reference's ``with nogil`` blocks).
"""
import ctypes as C
import errno as _errno
import logging
import os
from pickle import PickleError

import numpy

from . import _lib
from .devmem import DeviceArray, as_device_view, is_device_array

logger = logging.getLogger('pgsd.fl')

_NP_TO_PGSD = {
    numpy.dtype(numpy.uint8): _lib.TYPE_UINT8, numpy.dtype(numpy.uint16): _lib.TYPE_UINT16,
    numpy.dtype(numpy.uint32): _lib.TYPE_UINT32, numpy.dtype(numpy.uint64): _lib.TYPE_UINT64,
    numpy.dtype(numpy.int8): _lib.TYPE_INT8, numpy.dtype(numpy.int16): _lib.TYPE_INT16,
    numpy.dtype(numpy.int32): _lib.TYPE_INT32, numpy.dtype(numpy.int64): _lib.TYPE_INT64,
    numpy.dtype(numpy.float32): _lib.TYPE_FLOAT, numpy.dtype(numpy.float64): _lib.TYPE_DOUBLE,
}
_PGSD_TO_NP = {v: k for k, v in _NP_TO_PGSD.items()}


def _raise_on_error(retval, extra):
    """Raise the appropriate error type (ref: fl.pyx:35-61)."""
    if retval == _lib.ERROR_IO:
        err = C.get_errno() or _errno.EIO
        raise IOError(err, os.strerror(err), extra)
    elif retval == _lib.ERROR_NOT_A_PGSD_FILE:
        raise RuntimeError("Not a PGSD file: " + extra)
    elif retval == _lib.ERROR_INVALID_PGSD_FILE_VERSION:
        raise RuntimeError("Unsupported PGSD file version: " + extra)
    elif retval == _lib.ERROR_FILE_CORRUPT:
        raise RuntimeError("Corrupt PGSD file: " + extra)
    elif retval == _lib.ERROR_MEMORY_ALLOCATION_FAILED:
        raise MemoryError("Memory allocation failed: " + extra)
    elif retval == _lib.ERROR_NAMELIST_FULL:
        raise RuntimeError("PGSD namelist is full: " + extra)
    elif retval == _lib.ERROR_FILE_MUST_BE_WRITABLE:
        raise RuntimeError("File must be writable: " + extra)
    elif retval == _lib.ERROR_FILE_MUST_BE_READABLE:
        raise RuntimeError("File must be readable: " + extra)
    elif retval == _lib.ERROR_INVALID_ARGUMENT:
        raise RuntimeError("Invalid pgsd argument: " + extra + " " + _lib.last_error())
    elif retval != 0:
        raise RuntimeError("Unknown error: " + extra + " " + _lib.last_error())


def open(name, mode, application=None, schema=None, schema_version=None):
    """Open a PGSD file and return a :py:class:`PGSDFile` (ref: fl.pyx:149-228).

    Modes: ``'r'`` read, ``'r+'`` read/write existing, ``'w'`` create/overwrite, ``'x'`` create
    exclusively, ``'a'`` read/write, created if missing.
    """
    return PGSDFile(str(name), mode, application, schema, schema_version)


class PGSDFile:
    """PGSD file access interface (ref: fl.pyx:231-380)."""

    def __init__(self, name, mode, application, schema, schema_version):
        self._lib = _lib.load()
        self._handle = _lib.Handle()
        self._is_open = False
        self._mode = mode
        exclusive_create = 0
        overwrite = 0
        if mode == 'w':
            c_flags = _lib.OPEN_READWRITE
            overwrite = 1
        elif mode == 'r':
            c_flags = _lib.OPEN_READONLY
        elif mode == 'r+':
            c_flags = _lib.OPEN_READWRITE
        elif mode == 'x':
            c_flags = _lib.OPEN_READWRITE
            overwrite = 1
            exclusive_create = 1
        elif mode == 'a':
            c_flags = _lib.OPEN_READWRITE
            if not os.path.exists(name):
                overwrite = 1
        else:
            raise ValueError("Invalid mode: " + mode)
        self._name = name

        if overwrite:
            if application is None:
                raise ValueError("Provide application when creating a file")
            if schema is None:
                raise ValueError("Provide schema when creating a file")
            if schema_version is None:
                raise ValueError("Provide schema_version when creating a file")
            logger.info('overwriting file: ' + name + ' with mode: ' + mode + ', application: ' + application
                        + ', schema: ' + schema + ', and schema_version: ' + str(schema_version))
            if exclusive_create and os.path.exists(name) and self._lib.pgsd_b200_comm_size() == 1:
                raise FileExistsError(_errno.EEXIST, os.strerror(_errno.EEXIST), name)
            version = self._lib.pgsd_make_version(int(schema_version[0]), int(schema_version[1]))
            retval = self._lib.pgsd_create_and_open(C.byref(self._handle), name.encode('utf-8'),
                                                    application.encode('utf-8'), schema.encode('utf-8'),
                                                    version, c_flags, exclusive_create)
        else:
            logger.info('opening file: ' + name + ' with mode: ' + mode)
            if not os.path.exists(name):
                raise FileNotFoundError(_errno.ENOENT, os.strerror(_errno.ENOENT), name)
            retval = self._lib.pgsd_open(C.byref(self._handle), name.encode('utf-8'), c_flags)
        _raise_on_error(retval, name)
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
        if self._is_open:
            logger.info('closing file: ' + self._name)
            retval = self._lib.pgsd_close(C.byref(self._handle))
            self._is_open = False
            _raise_on_error(retval, self._name)

    def end_frame(self, write_all=True):
        """Complete the current frame (ref: fl.pyx:460-505).

        With several ranks this is where the frame's chunks get their file offsets: one
        all-gather of the chunk sizes + exclusive scan (K2) instead of per-chunk collectives.
        """
        self._check_open()
        logger.debug('end frame: ' + self._name)
        _raise_on_error(self._lib.pgsd_end_frame(C.byref(self._handle)), self._name)

    def flush(self, write_all=True):
        """Flush all buffered frames to the file and wait for queued device writes (ref: fl.pyx:507-524)."""
        self._check_open()
        logger.debug('flush: ' + self._name)
        _raise_on_error(self._lib.pgsd_flush(C.byref(self._handle)), self._name)

    # ------------------------------------------------------------------ write
    def _offset_args(self, N, M, offset, rank):
        # ref: fl.pyx:594-598 -- `offset` holds the row counts of all ranks
        if offset is None:
            return N, 0
        if isinstance(offset, str):
            if offset != 'auto':
                raise ValueError("offset must be None, 'auto' or an array of per-rank row counts")
            return _lib.N_GLOBAL_AUTO, _lib.OFFSET_AUTO
        offset = numpy.asarray(offset)
        return int(offset.sum()), int(M) * int(offset[0:rank].sum())

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
        data_ptr = data_array.ctypes.data if data_array.size else None
        logger.debug('write chunk: ' + self._name + ' - ' + name)
        gsize = 0 if N_global == _lib.N_GLOBAL_AUTO else N_global * M
        retval = self._lib.pgsd_write_chunk(C.byref(self._handle), name.encode('utf-8'), pgsd_type, N, M,
                                            N_global, M, stride, gsize, bool(write_all), 0, data_ptr)
        _raise_on_error(retval, self._name)

    def _write_chunk_device(self, name, data, offset, rank, write_all):
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
        logger.debug('write chunk (device): ' + self._name + ' - ' + name)
        if contiguous or N == 0:
            gsize = 0 if N_global == _lib.N_GLOBAL_AUTO else N_global * M
            retval = self._lib.pgsd_write_chunk(C.byref(self._handle), name.encode('utf-8'), pgsd_type, N, M,
                                                N_global, M, stride, gsize, bool(write_all), 0,
                                                ptr if N else None)
        else:
            # strided device array: K1 makes it contiguous (device-side ascontiguousarray)
            if M > 8 or any(s % item for s in strides):
                raise ValueError("strided device arrays need M <= 8 and element-aligned strides: " + name)
            logger.warning('implicit device pack when writing chunk: ' + name)
            col_stride = strides[1] if len(shape) == 2 else item
            cols = (_lib.Column * M)(*[_lib.Column(ptr + j * col_stride, strides[0] // item) for j in range(M)])
            retval = self._lib.pgsd_b200_write_chunk_soa(C.byref(self._handle), name.encode('utf-8'), pgsd_type,
                                                         N, M, N_global, M, stride, bool(write_all), pgsd_type, cols)
        del keep
        _raise_on_error(retval, self._name)

    def write_chunk_soa(self, name, columns, dtype=None, offset=None, rank=0, write_all=True):
        """Pack M component arrays into one (N, M) chunk on the device and write it (K1).

        ``columns`` is a sequence of M equally long 1-D arrays of one dtype -- CUDA arrays (hot
        path) or numpy arrays (uploaded first).  ``dtype`` is the chunk's dtype (default: the
        columns' dtype); the cast follows ``numpy.astype``.  This is the device form of
        ``numpy.ascontiguousarray(numpy.stack(columns, 1), dtype)`` -- what the reference's
        callers do on the host before ``write_chunk`` (fl.pyx:571, hoomd.py:206-270).
        """
        self._check_open()
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
        src_type = _NP_TO_PGSD.get(src_dt)
        dst_type = _NP_TO_PGSD.get(numpy.dtype(dtype) if dtype is not None else src_dt)
        if src_type is None or dst_type is None:
            raise ValueError("invalid type for chunk: " + name)
        N_global, stride = self._offset_args(N, M, offset, rank)
        cols = (_lib.Column * M)(*[_lib.Column(v[0] if N else None, v[3]) for v in views])
        retval = self._lib.pgsd_b200_write_chunk_soa(C.byref(self._handle), name.encode('utf-8'), dst_type, N, M,
                                                     N_global, M, stride, bool(write_all), src_type, cols)
        del keep
        _raise_on_error(retval, self._name)

    def _soa_views(self, name, columns):
        """-> (M x (ptr, stride), N, src dtype, keep-alive list) for 1..8 equally long 1-D columns."""
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

    def prepare_frame_soa(self, chunks, rank=0):
        """Build the reusable C descriptor table for :py:meth:`write_frame_soa`.

        ``chunks``: sequence of ``(name, columns, dtype, offset, write_all)`` with the meaning of
        :py:meth:`write_chunk_soa`'s arguments.  The returned object keeps the arrays alive and can
        be written any number of times (one simulation's buffers, one frame per time step).
        """
        n = len(chunks)
        descs = (_lib.ChunkDesc * max(n, 1))()
        keep = []
        for i, (name, columns, dtype, offset, write_all) in enumerate(chunks):
            views, N, src_dt, k = self._soa_views(name, columns)
            M = len(views)
            src_type = _NP_TO_PGSD.get(src_dt)
            dst_type = _NP_TO_PGSD.get(numpy.dtype(dtype) if dtype is not None else src_dt)
            if src_type is None or dst_type is None:
                raise ValueError("invalid type for chunk: " + name)
            N_global, stride = self._offset_args(N, M, offset, rank)
            cols = (_lib.Column * M)(*[_lib.Column(v[0] if N else None, v[3]) for v in views])
            bname = name.encode('utf-8')
            keep.extend([k, cols, bname])
            descs[i] = _lib.ChunkDesc(bname, dst_type, src_type, N, M, N_global, M, stride, bool(write_all),
                                      cols)
        return (descs, n, keep)

    def write_frame_soa(self, prepared):
        """Write all SoA chunks of a frame with ONE K1 launch (``pgsd_b200_write_chunks_soa``);
        ``prepared`` comes from :py:meth:`prepare_frame_soa`.  Equivalent to calling
        :py:meth:`write_chunk_soa` for each chunk in order."""
        self._check_open()
        descs, n, _ = prepared
        _raise_on_error(self._lib.pgsd_b200_write_chunks_soa(C.byref(self._handle), n, descs), self._name)

    # ------------------------------------------------------------------ read
    def chunk_exists(self, frame, name, write_all=True):
        """Test if a chunk exists (ref: fl.pyx:656-715)."""
        self._check_open()
        logger.debug('chunk exists: ' + self._name + ' - ' + name)
        entry = self._lib.pgsd_find_chunk(C.byref(self._handle), int(frame), name.encode('utf-8'))
        return bool(entry)

    def read_chunk(self, frame, name, N=0, M=0, offset=0, r_all=False, device=False):
        """Read a data chunk (ref: fl.pyx:717-874).

        ``r_all=False`` reads the whole (N_global, M) chunk.  ``r_all=True`` reads ``N`` rows of
        ``M`` values starting ``offset`` rows into the chunk and returns exactly those rows
        (the reference returns an (N_global, M) array whose first N rows are filled).
        ``device=True`` returns a :py:class:`~pgsd_sph_b200.devmem.DeviceArray`.
        Raises KeyError if the chunk does not exist.
        """
        self._check_open()
        entry_p = self._lib.pgsd_find_chunk(C.byref(self._handle), int(frame), name.encode('utf-8'))
        if not entry_p:
            raise KeyError("frame " + str(frame) + " / chunk " + name + " not found in: " + self._name)
        entry = _lib.IndexEntry.from_buffer_copy(entry_p.contents)
        dtype = _PGSD_TO_NP.get(entry.type)
        if dtype is None:
            raise ValueError("invalid type for chunk: " + name)
        rows = int(N) if r_all else int(entry.N)
        cols = int(M) if r_all else int(entry.M)
        if r_all and cols != entry.M:
            raise ValueError("M must equal the chunk's M for a partial read: " + name)
        logger.debug('read chunk: ' + self._name + ' - ' + str(frame) + ' - ' + name)
        if device:
            out = DeviceArray((rows, cols), dtype)
            ptr = out.ptr
        else:
            out = numpy.empty(dtype=dtype, shape=[rows, cols])
            ptr = out.ctypes.data
        # only read chunk if we have data
        if rows != 0 and cols != 0:
            retval = self._lib.pgsd_read_chunk(C.byref(self._handle), ptr, C.byref(entry), rows, cols,
                                               int(offset), bool(r_all))
            _raise_on_error(retval, self._name)
        if entry.M == 1:
            return out.reshape([rows])
        return out

    def find_matching_chunk_names(self, match, write_all=True):
        """All chunk names in the file that start with ``match`` (ref: fl.pyx:876-945)."""
        self._check_open()
        c_match = match.encode('utf-8')
        retval = []
        found = self._lib.pgsd_find_matching_chunk_name(C.byref(self._handle), c_match, None)
        while found:
            retval.append(C.string_at(found).decode('utf-8'))
            found = self._lib.pgsd_find_matching_chunk_name(C.byref(self._handle), c_match, found)
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
        v = self._handle.header.pgsd_version
        return (v >> 16, v & 0xffff)

    @property
    def schema_version(self):
        v = self._handle.header.schema_version
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
        return self._lib.pgsd_get_nframes(C.byref(self._handle))

    @property
    def nnames(self):
        self._check_open()
        return self._lib.pgsd_get_nnames(C.byref(self._handle))

    @property
    def maximum_write_buffer_size(self):
        self._check_open()
        return self._lib.pgsd_get_maximum_write_buffer_size(C.byref(self._handle))

    @maximum_write_buffer_size.setter
    def maximum_write_buffer_size(self, size):
        self._check_open()
        _raise_on_error(self._lib.pgsd_set_maximum_write_buffer_size(C.byref(self._handle), int(size)), self._name)

    @property
    def index_entries_to_buffer(self):
        self._check_open()
        return self._lib.pgsd_get_index_entries_to_buffer(C.byref(self._handle))

    @index_entries_to_buffer.setter
    def index_entries_to_buffer(self, number):
        self._check_open()
        _raise_on_error(self._lib.pgsd_set_index_entries_to_buffer(C.byref(self._handle), int(number)), self._name)

    def __del__(self):
        try:
            if self._is_open:
                logger.info('closing file: ' + self._name)
                self._lib.pgsd_close(C.byref(self._handle))
                self._is_open = False
        except Exception:
            pass
