#!/usr/bin/env python
"""BASELINE config 5 shape: many small frames (N = 4096 particles + 8 log scalars per frame, 18 chunks),
device-resident fields -> file.  Prints frames/s and us/frame for this library (one rank, or N ranks
under torchrun: rows split over the ranks, offsets from one NCCL all-gather + device scan per frame)
and, at one rank, for the unmodified reference (oracle/_ref/ref_driver bench mode, same N) on the same
file system.  Latency measurement tool (offset scan + index/namelist append path); not part of the library.

    python tools/small_frames.py [frames]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/small_frames.py [frames]
"""
import json, os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pgsd_sph_b200 import _lib, comm, fl, synth
from pgsd_sph_b200.devmem import DeviceArray

n_total = 4096
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
dist = bench.Dist(int(os.environ.get("WORLD_SIZE", "1")))
lib = _lib.load(); _lib.check(lib.pgsd_b200_device_init(dist.local), "init")
if dist.world > 1:
    comm.init_nccl(dist.rank, dist.world, dist.bcast_bytes, dist.local)
rows, start = bench.rank_rows(n_total, dist.world, dist.rank)
n = rows[dist.rank]
cols = bench.make_soa(n_total, start, n, 5)
d = [DeviceArray.from_numpy(c) for c in cols]
path = os.path.join(bench.bench_dir(), "small.gsd")
logs = [("log/value/v%d" % k, np.array([k], dtype=np.float32)) for k in range(8)]
for mode in ("device", "host"):
    src = d if mode == "device" else cols
    dist.barrier()
    with fl.open(path, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
        prep = f.prepare_frame_soa([(nm, [src[j] for j in idx], dt, rows, True) for nm, idx, dt in bench.SOA_CHUNKS],
                                   rank=dist.rank)
        dist.barrier()
        t0 = time.perf_counter()
        for i in range(frames):
            for k, a in synth.frame_scalars(n_total, i):
                f.write_chunk(k, a, write_all=False)
            f.write_frame_soa(prep)
            for k, a in logs:
                f.write_chunk(k, a, write_all=False)
            f.end_frame()
        f.flush()
        dt_ = dist.max(time.perf_counter() - t0)
    dist.barrier()
    if dist.rank == 0:
        sz = os.path.getsize(path)
        print(json.dumps({"impl": "b200", "ranks": dist.world, "comm": lib.pgsd_b200_comm_kind().decode(), "source": mode,
                          "frames": frames, "N": n_total, "frames_per_s": frames / dt_, "us_per_frame": 1e6 * dt_ / frames,
                          "file_bytes": sz}), flush=True)
        os.unlink(path)
drv = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_driver")
if dist.world == 1 and os.path.exists(drv):
    blob = os.path.join(bench.bench_dir(), "small_blob.bin")
    with open(blob, "wb") as fh:
        for c in cols:
            fh.write(c.tobytes())
    for P in (1, 8):
        r = subprocess.run([drv, "bench", path, str(n_total), str(frames), blob], env=dict(os.environ, PGSD_SHIM_NP=str(P)),
                           capture_output=True, text=True)
        t = json.loads(r.stdout.strip().splitlines()[-1])["frame_s"]
        print(json.dumps({"impl": "reference pgsd.c (10 chunks/frame, no log scalars)", "ranks": P, "frames": frames,
                          "frames_per_s": len(t) / sum(t), "us_per_frame": 1e6 * sum(t) / len(t)}), flush=True)
        os.unlink(path)
    os.unlink(blob)
if dist.world > 1:
    comm.finalize()
dist.close()
lib.pgsd_b200_shutdown()
