"""Minimal device / pinned arrays over libpgsd_b200's raw CUDA helpers (no PyTorch dependency).

``DeviceArray`` owns (or views) CUDA device memory and exposes ``__cuda_array_interface__`` v3,
so it interoperates zero-copy with torch / cupy / numba and is what ``pgsd.fl`` accepts as chunk
data and hands back from device reads.  ``as_device_view`` parses any object that exposes
``__cuda_array_interface__`` or ``__dlpack__`` into (pointer, shape, dtype, strides).
"""
import ctypes as C

import numpy as np

from . import _lib

H2D, D2H, D2D = 1, 2, 3


class DeviceArray:
    """C-contiguous n-d array in CUDA device memory."""

    def __init__(self, shape, dtype, ptr=None, owner=None):
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape)) if self.shape else 1
        self.nbytes = self.size * self.dtype.itemsize
        self._owner = owner
        self._owned = ptr is None
        if ptr is None:
            p = C.c_void_p()
            _lib.check(_lib.load().pgsd_b200_malloc(C.byref(p), max(self.nbytes, 1)), "pgsd_b200_malloc")
            ptr = p.value
        self.ptr = int(ptr)

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False),
                "version": 3, "strides": None, "stream": 1}

    @classmethod
    def from_numpy(cls, a):
        a = np.ascontiguousarray(a)
        d = cls(a.shape, a.dtype)
        if a.nbytes:
            _lib.check(_lib.load().pgsd_b200_memcpy(d.ptr, a.ctypes.data, a.nbytes, H2D), "H2D copy")
        return d

    def to_numpy(self):
        out = np.empty(self.shape, dtype=self.dtype)
        if self.nbytes:
            _lib.check(_lib.load().pgsd_b200_memcpy(out.ctypes.data, self.ptr, self.nbytes, D2H), "D2H copy")
        return out

    def reshape(self, *shape):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape
        shape = tuple(int(s) for s in shape)
        if -1 in shape:
            known = int(np.prod([s for s in shape if s != -1])) or 1
            shape = tuple(self.size // known if s == -1 else s for s in shape)
        assert int(np.prod(shape)) == self.size
        return DeviceArray(shape, self.dtype, ptr=self.ptr, owner=self)

    def free(self):
        if self._owned and self.ptr:
            _lib.load().pgsd_b200_free(self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, ptr=0x{self.ptr:x})"


class PinnedArray:
    """numpy array backed by page-locked host memory (cudaHostAlloc)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        shape = tuple(shape) if isinstance(shape, (tuple, list)) else (int(shape),)
        nbytes = int(np.prod(shape)) * self.dtype.itemsize
        p = C.c_void_p()
        _lib.check(_lib.load().pgsd_b200_host_alloc(C.byref(p), max(nbytes, 1)), "pgsd_b200_host_alloc")
        self.ptr = p.value
        buf = (C.c_char * max(nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            _lib.load().pgsd_b200_host_free(self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _PinnedBlock:
    """One page-locked allocation; goes back to its pool when the last numpy view of it dies."""

    def __init__(self, pool, ptr, nbytes):
        self.pool, self.ptr, self.nbytes = pool, ptr, nbytes

    def __del__(self):
        try:
            self.pool._give_back(self.ptr, self.nbytes)
        except Exception:
            pass


class PinnedPool:
    """Recycling allocator of page-locked numpy arrays.

    ``cudaHostAlloc`` and first-touch page faults cost more than the D2H copy of a frame, so the
    arrays a reordered frame is returned in come from here: when the caller drops a frame (e.g. the
    loop variable of ``for frame in trajectory`` is rebound) its buffers are reused for the next one.
    """

    def __init__(self, max_cached_bytes=16 << 30):
        self._free = {}
        self._cached = 0
        self._max = int(max_cached_bytes)

    def empty(self, shape, dtype):
        dtype = np.dtype(dtype)
        shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        count = int(np.prod(shape)) if shape else 1
        nbytes = max(count * dtype.itemsize, 1)
        lst = self._free.get(nbytes)
        if lst:
            ptr = lst.pop()
            self._cached -= nbytes
        else:
            p = C.c_void_p()
            _lib.check(_lib.load().pgsd_b200_host_alloc(C.byref(p), nbytes), "pgsd_b200_host_alloc")
            ptr = p.value
        buf = (C.c_char * nbytes).from_address(ptr)
        buf._block = _PinnedBlock(self, ptr, nbytes)  # numpy's memoryview keeps buf (and the block) alive
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def _give_back(self, ptr, nbytes):
        if self._cached + nbytes <= self._max:
            self._free.setdefault(nbytes, []).append(ptr)
            self._cached += nbytes
        else:
            _lib.load().pgsd_b200_host_free(ptr)

    def clear(self):
        lib = _lib.load()
        for lst in self._free.values():
            for ptr in lst:
                lib.pgsd_b200_host_free(ptr)
        self._free, self._cached = {}, 0


_default_pool = None


def pinned_pool():
    global _default_pool
    if _default_pool is None:
        _default_pool = PinnedPool()
    return _default_pool


def download(dev, pool=None):
    """DeviceArray -> numpy array in pooled page-locked memory (one DMA, no staging copy)."""
    out = (pool or pinned_pool()).empty(dev.shape, dev.dtype)
    if dev.nbytes:
        _lib.check(_lib.load().pgsd_b200_memcpy(out.ctypes.data, dev.ptr, dev.nbytes, D2H), "D2H copy")
    return out


# ---- DLPack (v0.x capsule "dltensor") -------------------------------------------------------
class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int), ("device_id", C.c_int)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int), ("dtype", _DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class _DLManagedTensor(C.Structure):
    _fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", C.c_void_p), ("deleter", C.c_void_p)]


_DL_CODES = {0: "i", 1: "u", 2: "f"}
_DL_CUDA, _DL_CUDA_HOST, _DL_CUDA_MANAGED = 2, 3, 13


def _from_dlpack(obj):
    cap = obj.__dlpack__()
    api = C.pythonapi
    api.PyCapsule_GetPointer.restype = C.c_void_p
    api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    p = api.PyCapsule_GetPointer(cap, b"dltensor")
    t = C.cast(p, C.POINTER(_DLManagedTensor)).contents.dl_tensor
    if t.device.device_type not in (_DL_CUDA, _DL_CUDA_MANAGED):
        raise ValueError("DLPack tensor is not in CUDA device memory")
    if t.dtype.lanes != 1 or t.dtype.code not in _DL_CODES:
        raise ValueError("unsupported DLPack dtype")
    dtype = np.dtype(f"{_DL_CODES[t.dtype.code]}{t.dtype.bits // 8}")
    shape = tuple(t.shape[i] for i in range(t.ndim))
    strides = None
    if t.strides:
        strides = tuple(t.strides[i] * dtype.itemsize for i in range(t.ndim))
    # the capsule (kept alive by the caller's object for the duration of the call) owns nothing we free
    return int(t.data or 0) + int(t.byte_offset), shape, dtype, strides, cap


def is_device_array(obj):
    return hasattr(obj, "__cuda_array_interface__") or (
        hasattr(obj, "__dlpack__") and hasattr(obj, "__dlpack_device__")
        and obj.__dlpack_device__()[0] in (_DL_CUDA, _DL_CUDA_MANAGED))


def as_device_view(obj):
    """-> (ptr, shape, dtype, byte_strides or None, keepalive) for a CUDA array-like."""
    if hasattr(obj, "__cuda_array_interface__"):
        cai = obj.__cuda_array_interface__
        dtype = np.dtype(cai["typestr"])
        strides = cai.get("strides")
        return int(cai["data"][0] or 0), tuple(cai["shape"]), dtype, (tuple(strides) if strides else None), obj
    if hasattr(obj, "__dlpack__"):
        return _from_dlpack(obj)
    raise TypeError("object exposes neither __cuda_array_interface__ nor __dlpack__")
