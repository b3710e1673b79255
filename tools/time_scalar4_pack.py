import sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np
from pgsd_sph_b200 import _lib
from pgsd_sph_b200.devmem import DeviceArray
lib = _lib.load(); _lib.check(lib.pgsd_b200_device_init(0), "init")
n = 64 * 1024 * 1024
d = DeviceArray((n, 4), np.float32); out = DeviceArray((n, 3), np.float32)
cols = (_lib.Column * 3)(*[_lib.Column(d.ptr + 4 * j, 4) for j in range(3)])
t = C.c_void_p(); lib.pgsd_b200_timer_create(C.byref(t))
for rep in range(4):
    lib.pgsd_b200_timer_start(t)
    lib.pgsd_b200_pack_soa(out.ptr, _lib.TYPE_FLOAT, n, 3, _lib.TYPE_FLOAT, cols, None)
    ms = C.c_float(); lib.pgsd_b200_timer_stop(t, C.byref(ms))
    print(f"Scalar4 -> (N,3): {ms.value:.3f} ms, algorithmic {n*24/ms.value/1e6:.0f} GB/s (touched {n*28/ms.value/1e6:.0f} GB/s)")
