#!/usr/bin/env python
"""Short driver for ncu captures: runs each hot kernel a few times on device-resident synthetic
data at the bench sizes, through the same C-ABI entry points bench.py uses (no file I/O).

    python tools/profile_kernels.py [--pack-particles N] [--sort-particles N] [--iters K]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgsd_sph_b200 import _lib  # noqa: E402
from pgsd_sph_b200.devmem import DeviceArray  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pack-particles", type=int, default=64 * 1024 * 1024)
    ap.add_argument("--sort-particles", type=int, default=16 * 1024 * 1024)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--frame", action="store_true",
                    help="also write one whole 6-chunk frame through pgsd.fl (the real one-launch-per-frame K1)")
    a = ap.parse_args()
    lib = _lib.load()
    _lib.check(lib.pgsd_b200_device_init(0), "device_init")
    rng = np.random.default_rng(1)

    # K1: position (3 x f32 columns -> (N,3)), the largest chunk of the frame
    n = a.pack_particles
    cols_h = [rng.random(n, dtype=np.float32) for _ in range(3)]
    cols_d = [DeviceArray.from_numpy(c) for c in cols_h]
    out = DeviceArray((n, 3), np.float32)
    cols = (_lib.Column * 3)(*[_lib.Column(c.ptr, 1) for c in cols_d])
    for _ in range(a.iters):
        _lib.check(lib.pgsd_b200_pack_soa(out.ptr, _lib.TYPE_FLOAT, n, 3, _lib.TYPE_FLOAT, cols, None), "pack")
    lib.pgsd_b200_synchronize()
    got = out.to_numpy()
    assert np.array_equal(got[:1000], np.stack([c[:1000] for c in cols_h], axis=1))
    for d in cols_d + [out]:
        d.free()
    del cols_h, got

    if a.frame:
        import tempfile
        from pgsd_sph_b200 import fl
        d = "/dev/shm" if os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
        path = os.path.join(d, "pgsd_profile_frame.gsd")
        cols_h = [rng.random(n, dtype=np.float32) for _ in range(8)] + [rng.integers(0, 3, n, dtype=np.uint32),
                                                                       rng.permutation(n).astype(np.uint32)]
        cols_d = [DeviceArray.from_numpy(c) for c in cols_h]
        chunks = (("particles/position", (0, 1, 2)), ("particles/velocity", (3, 4, 5)), ("particles/typeid", (8,)),
                  ("particles/density", (6,)), ("particles/pressure", (7,)), ("log/particles/id", (9,)))
        with fl.open(path, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
            prep = f.prepare_frame_soa([(nm, [cols_d[j] for j in idx], None, None, True) for nm, idx in chunks])
            for _ in range(2):
                f.write_frame_soa(prep)
                f.end_frame()
        os.unlink(path)
        for d_ in cols_d:
            d_.free()
        del cols_h

    # K4 + K5: 40 B/particle frame
    n = a.sort_particles
    ids = rng.permutation(n).astype(np.uint32)
    fields_h = [rng.random((n, 3), dtype=np.float32), rng.random((n, 3), dtype=np.float32),
                rng.integers(0, 3, n, dtype=np.uint32), rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)]
    d_in = [DeviceArray.from_numpy(f) for f in fields_h]
    d_out = [DeviceArray(f.shape, f.dtype) for f in fields_h]
    d_ids, d_sorted, d_perm = DeviceArray.from_numpy(ids), DeviceArray((n,), np.uint32), DeviceArray((n,), np.uint32)
    fields = (_lib.Field * 5)(*[_lib.Field(i.ptr, o.ptr, f.dtype.itemsize * (f.shape[1] if f.ndim > 1 else 1))
                                for i, o, f in zip(d_in, d_out, fields_h)])
    for _ in range(a.iters):
        _lib.check(lib.pgsd_b200_reorder_device(n, d_ids.ptr, d_sorted.ptr, d_perm.ptr, 5, fields, None), "reorder")
    lib.pgsd_b200_synchronize()
    assert np.array_equal(d_sorted.to_numpy(), np.arange(n, dtype=np.uint32))
    perm = d_perm.to_numpy()
    assert np.array_equal(d_out[3].to_numpy(), fields_h[3][perm])
    print("profile_kernels ok")
    lib.pgsd_b200_shutdown()


if __name__ == "__main__":
    main()
