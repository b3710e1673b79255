#!/usr/bin/env python
"""Where the end-to-end reordered read spends its time (host wall clock per phase)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgsd_sph_b200 import _lib, fl, hoomd, synth
from pgsd_sph_b200.devmem import download
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16 * 1024 * 1024
lib = _lib.load(); _lib.check(lib.pgsd_b200_device_init(0), "init")
path = os.path.join(bench.bench_dir(), "prof_read.gsd")
with fl.open(path, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
    for i in range(3):
        cols = bench.make_soa(n, 0, n, 7000 + i)
        for k, a in synth.frame_scalars(n, i):
            f.write_chunk(k, a, write_all=False)
        f.write_frame_soa(f.prepare_frame_soa([(nm, [cols[j] for j in idx], dt, None, True) for nm, idx, dt in bench.SOA_CHUNKS]))
        f.end_frame()
f = fl.open(path, 'r')
names = [c[0] for c in bench.SOA_CHUNKS]
for rep in range(4):
    t0 = time.perf_counter()
    d = {nm: f.read_chunk(1 + rep % 2, nm, device=True) for nm in names}
    lib.pgsd_b200_synchronize(); t1 = time.perf_counter()
    ids = d.pop("log/particles/id")
    sid, out = hoomd.reorder_by_id(ids, d, device=True)
    lib.pgsd_b200_synchronize(); t2 = time.perf_counter()
    host = {k: download(v) for k, v in out.items()}; hs = download(sid)
    t3 = time.perf_counter()
    for v in list(out.values()) + [sid, ids] + list(d.values()): v.free()
    t4 = time.perf_counter()
    print(f"rep {rep}: read->device {1e3*(t1-t0):.1f} ms ({n*40/(t1-t0)/1e9:.1f} GB/s)  reorder {1e3*(t2-t1):.1f} ms  download {1e3*(t3-t2):.1f} ms ({n*40/(t3-t2)/1e9:.1f} GB/s)  free {1e3*(t4-t3):.1f} ms", flush=True)
    del host, hs
f.close()
t = hoomd.open(path, 'r', reorder='id')
for rep in range(4):
    t0 = time.perf_counter(); fr = t[1 + rep % 2]; t1 = time.perf_counter()
    print(f"traj[i] {1e3*(t1-t0):.1f} ms", flush=True)
t.close(); os.unlink(path)
