#!/usr/bin/env python
"""CUDA-event timings of kernel variants bench.py does not cover: K1 on HOOMD Scalar4 records, K1 with
float64 sources (f64 -> f32 cast), and the reorder at 64 Mi particles (config 3 read-back, 26-bit ids)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pgsd_sph_b200 import _lib
from pgsd_sph_b200.devmem import DeviceArray

lib = _lib.load(); _lib.check(lib.pgsd_b200_device_init(0), "init")
PEAK = 6546.6
t = C.c_void_p(); lib.pgsd_b200_timer_create(C.byref(t))


def timed(fn, reps=4):
    best = 1e9
    for _ in range(reps):
        lib.pgsd_b200_timer_start(t); fn(); ms = C.c_float(); lib.pgsd_b200_timer_stop(t, C.byref(ms)); best = min(best, ms.value)
    return best


n = 64 * 1024 * 1024
d = DeviceArray((n, 4), np.float32); out = DeviceArray((n, 3), np.float32)
cols = (_lib.Column * 3)(*[_lib.Column(d.ptr + 4 * j, 4) for j in range(3)])
ms = timed(lambda: lib.pgsd_b200_pack_soa(out.ptr, _lib.TYPE_FLOAT, n, 3, _lib.TYPE_FLOAT, cols, None))
print(f"K1 Scalar4 records -> (N,3) f32, {n} rows: {ms:.3f} ms; 28 B/row touched -> {n*28/ms/1e6:.0f} GB/s = {n*28/ms/1e6/PEAK:.2f} of peak")
d.free()
srcs = [DeviceArray((n,), np.float64) for _ in range(3)]
cols = (_lib.Column * 3)(*[_lib.Column(s.ptr, 1) for s in srcs])
ms = timed(lambda: lib.pgsd_b200_pack_soa(out.ptr, _lib.TYPE_FLOAT, n, 3, _lib.TYPE_DOUBLE, cols, None))
print(f"K1 3 x f64 columns -> (N,3) f32, {n} rows: {ms:.3f} ms; 36 B/row -> {n*36/ms/1e6:.0f} GB/s = {n*36/ms/1e6/PEAK:.2f} of peak")
for s in srcs + [out]:
    s.free()

rng = np.random.default_rng(1)
ids = rng.permutation(n).astype(np.uint32)
shapes = [(n, 3), (n, 3), (n,), (n,), (n,)]
d_in = [DeviceArray(s, np.float32) for s in shapes]; d_out = [DeviceArray(s, np.float32) for s in shapes]
d_ids, d_sorted, d_perm = DeviceArray.from_numpy(ids), DeviceArray((n,), np.uint32), DeviceArray((n,), np.uint32)
fields = (_lib.Field * 5)(*[_lib.Field(i.ptr, o.ptr, 4 * (s[1] if len(s) > 1 else 1)) for i, o, s in zip(d_in, d_out, shapes)])
lib.pgsd_b200_reorder_profiling(1)
best = None
for _ in range(4):
    _lib.check(lib.pgsd_b200_reorder_device(n, d_ids.ptr, d_sorted.ptr, d_perm.ptr, 5, fields, None), "reorder")
    ph = (C.c_float * 4)(); lib.pgsd_b200_reorder_phase_ms(ph)
    ph = [float(x) for x in ph]
    if best is None or sum(ph) < sum(best):
        best = ph
assert np.array_equal(d_sorted.to_numpy()[:1000], np.arange(1000, dtype=np.uint32))
print(f"reorder {n} particles (40 B payload + perm): census {best[0]:.3f}, bucket {best[1]:.3f}, pair passes {best[2]:.3f}, "
      f"gather {best[3]:.3f} ms = {sum(best):.3f} ms -> {n/sum(best)/1e3:.0f} Mparticles/s")
