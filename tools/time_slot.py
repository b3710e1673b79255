#!/usr/bin/env python
"""CUDA-event timings of the reorder (pgsd_b200_reorder_device, perm not requested -- the call
pgsd.hoomd makes) for the slot path variants against the general path, 16 Mi and 64 Mi particles."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pgsd_sph_b200 import _lib
from pgsd_sph_b200.devmem import DeviceArray

lib = _lib.load(); _lib.check(lib.pgsd_b200_device_init(0), "init")
PEAK = 6546.6
t = C.c_void_p(); lib.pgsd_b200_timer_create(C.byref(t))
KNOBS = ["PGSD_B200_SLOT_PER", "PGSD_B200_SLOT_UNIT", "PGSD_B200_SLOT_LAYOUT", "PGSD_B200_SLOT_CSTRIDE", "PGSD_B200_SLOT", "PGSD_B200_SLOT_BITS",
         "PGSD_B200_SLOT_TILE", "PGSD_B200_SLOT_BULK", "PGSD_B200_CLUSTER", "PGSD_B200_CLUSTER_TILE", "PGSD_B200_CLUSTER_THREADS",
         "PGSD_B200_CLUSTER_AGG", "PGSD_B200_CLUSTER_BITS"]
VARIANTS = [
    ("general path", {"PGSD_B200_SLOT": "0"}),
    ("slot path", {"PGSD_B200_CLUSTER": "0"}),
    ("slot per 4", {"PGSD_B200_SLOT_PER": "4"}),
    ("slot per 1", {"PGSD_B200_SLOT_PER": "1"}),
    ("slot per 2 t2048", {"PGSD_B200_SLOT_PER": "2", "PGSD_B200_SLOT_TILE": "2048"}),
    ("slot per 4 t2048", {"PGSD_B200_SLOT_PER": "4", "PGSD_B200_SLOT_TILE": "2048"}),
    ("slot per 2 t512", {"PGSD_B200_SLOT_PER": "2", "PGSD_B200_SLOT_TILE": "512"}),
    ("slot per 4 t512", {"PGSD_B200_SLOT_PER": "4", "PGSD_B200_SLOT_TILE": "512"}),
    ("cluster default", {}),
    ("cluster t4096/512", {"PGSD_B200_CLUSTER_TILE": "4096", "PGSD_B200_CLUSTER_THREADS": "512"}),
    ("cluster t2048/512", {"PGSD_B200_CLUSTER_TILE": "2048", "PGSD_B200_CLUSTER_THREADS": "512"}),
    ("cluster t2048/256", {"PGSD_B200_CLUSTER_TILE": "2048", "PGSD_B200_CLUSTER_THREADS": "256"}),
    ("cluster t1024/256", {"PGSD_B200_CLUSTER_TILE": "1024"}),
    ("cluster no-agg", {"PGSD_B200_CLUSTER_AGG": "0"}),
    ("cluster plain loads", {"PGSD_B200_SLOT_BULK": "0"}),
    ("cluster bits 11", {"PGSD_B200_CLUSTER_BITS": "11"}),
    ("cluster bits 11 t2048", {"PGSD_B200_CLUSTER_BITS": "11", "PGSD_B200_CLUSTER_TILE": "2048"}),
]
if os.environ.get("TIME_SLOT_ONLY"):
    VARIANTS = [v for v in VARIANTS if any(w in v[0] for w in os.environ["TIME_SLOT_ONLY"].split(","))]
sizes = [int(a) for a in sys.argv[1:]] or [16 << 20, 64 << 20]
for n in sizes:
    rng = np.random.default_rng(1)
    ids = rng.permutation(n).astype(np.uint32)
    shapes = [(n, 3), (n, 3), (n,), (n,), (n,)]
    d_in = [DeviceArray(s, np.float32) for s in shapes]
    d_out = [DeviceArray(s, np.float32) for s in shapes]
    # payload that encodes its own id: field 2 holds the id bits
    d_in[2].free(); d_in[2] = DeviceArray.from_numpy(ids.view(np.float32))
    d_ids, d_sorted, d_perm = DeviceArray.from_numpy(ids), DeviceArray((n,), np.uint32), DeviceArray((n,), np.uint32)
    fields = (_lib.Field * 5)(*[_lib.Field(i.ptr, o.ptr, 4 * (s[1] if len(s) > 1 else 1)) for i, o, s in zip(d_in, d_out, shapes)])
    for want_perm in (False, True):
        for name, env in VARIANTS:
            for k in KNOBS:
                os.environ.pop(k, None)
            os.environ.update(env)
            pp = d_perm.ptr if want_perm else None
            for _ in range(2):
                _lib.check(lib.pgsd_b200_reorder_device(n, d_ids.ptr, d_sorted.ptr, pp, 5, fields, None), "reorder")
            lib.pgsd_b200_synchronize()
            lib.pgsd_b200_reorder_profiling(1)
            best, tot = None, 1e9
            for _ in range(5):
                lib.pgsd_b200_timer_start(t)
                _lib.check(lib.pgsd_b200_reorder_device(n, d_ids.ptr, d_sorted.ptr, pp, 5, fields, None), "reorder")
                ms = C.c_float(); lib.pgsd_b200_timer_stop(t, C.byref(ms))
                ph = (C.c_float * 4)(); lib.pgsd_b200_reorder_phase_ms(ph)
                ph = [float(x) for x in ph]
                tot = min(tot, ms.value)
                if best is None or sum(ph) < sum(best):
                    best = ph
            lib.pgsd_b200_reorder_profiling(0)
            ok = np.array_equal(d_sorted.to_numpy(), np.arange(n, dtype=np.uint32)) and \
                np.array_equal(d_out[2].to_numpy().view(np.uint32), np.arange(n, dtype=np.uint32))
            print(f"n={n} perm={int(want_perm)} {name:18s}: census {best[0]:.3f} scatter/bucket {best[1]:.3f} pairs {best[2]:.3f} "
                  f"place/gather {best[3]:.3f} = {sum(best):.3f} ms; call {tot:.3f} ms -> {n/tot/1e3:.0f} Mparticles/s, "
                  f"80 B/particle = {80*n/tot/1e6/PEAK:.3f} of peak; ok={ok}", flush=True)
    for a in d_in + d_out + [d_ids, d_sorted, d_perm]:
        a.free()
