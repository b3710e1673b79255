#!/usr/bin/env python
"""One reorder of 16 Mi particles (slot path) for ncu:  ncu --kernel-name regex:k6_slot ... python tools/profile_slot.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pgsd_sph_b200 import _lib
from pgsd_sph_b200.devmem import DeviceArray

lib = _lib.load(); _lib.check(lib.pgsd_b200_device_init(0), "init")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ids = np.random.default_rng(1).permutation(n).astype(np.uint32)
shapes = [(n, 3), (n, 3), (n,), (n,), (n,)]
d_in = [DeviceArray(s, np.float32) for s in shapes]
d_out = [DeviceArray(s, np.float32) for s in shapes]
d_ids, d_sorted = DeviceArray.from_numpy(ids), DeviceArray((n,), np.uint32)
fields = (_lib.Field * 5)(*[_lib.Field(i.ptr, o.ptr, 4 * (s[1] if len(s) > 1 else 1)) for i, o, s in zip(d_in, d_out, shapes)])
for _ in range(reps):
    _lib.check(lib.pgsd_b200_reorder_device(n, d_ids.ptr, d_sorted.ptr, None, 5, fields, None), "reorder")
lib.pgsd_b200_synchronize()
assert np.array_equal(d_sorted.to_numpy()[:4096], np.arange(4096, dtype=np.uint32)) or os.environ.get("PGSD_B200_SLOT_DEBUG")
print("ok")
