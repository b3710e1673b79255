#!/usr/bin/env python
"""bench.py -- PGSD hot-path benchmark on B200 (contract: see DESIGN.md "Measurement").

Metric (BASELINE.json): frame write GB/s [primary: `metric`/`value`/`e2e`] and ID-reordered read
Mparticles/s [`read_reorder` object of the same JSON line], at 1/2/4/8 GPUs, beside the reference's
CPU path timed on the box's host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] ...                # the reference CPU path

One process per GPU (torchrun for N > 1).  A "step" of the write leg is one frame of the
BASELINE config-3 workload -- a 64 Mi-particle HOOMD-schema frame (40 B/particle, 2.68 GB), row-
partitioned over the N ranks (strong scaling) -- from 10 SoA field arrays to the .gsd file:
K1 pack, K2 offset scan, K3 pinned staging + file stage.  A step of the read leg is one 16 Mi-particle
frame (config 4) put into particle-ID order: K4 + K5; frames are independent, every rank processes
its own (weak).  `value` legs start with inputs resident in HBM; `e2e` legs start from host buffers /
the file and end in the file / host arrays.

Extra objects of the same line (each with its own cpu_baseline at N = 1):
  parity             one config-3 frame written at P = N under the run's communicator and by the unmodified
                     reference at PGSD_SHIM_NP = N from the same inputs: sha256 + size compared; the run FAILS
                     when they differ
  split              where a frame's wall time goes: K1 (events), D2H (events on the copy streams), file stage
                     (writer-thread time), and the host-only ceiling of the file stage on the same target
  write_disk         the same write leg on a disk-backed target (PGSD_BENCH_DIR2, default /tmp when it is not tmpfs)
  trajectory_write   config 2: 100 frames x 1 Mi particles
  small_frames       config 5: 10 k frames x 4096 particles + 8 log scalars (18 chunks): latency per frame
  read_reorder.vtu   config 4's converter leg: reordered frame -> pgsd2vtu arrays -> .vtu (container parity unpinned)
  distributed_reorder, benchmark_write, benchmark_read (the reference's own benchmark programs' workloads)

Only this file's `cpu_baseline` legs and `--impl reference` execute anything under oracle/.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WRITE_PARTICLES = 64 * 1024 * 1024   # BASELINE.json configs[2]
READ_PARTICLES = 16 * 1024 * 1024    # BASELINE.json configs[3]
TRAJ_PARTICLES, TRAJ_FRAMES = 1024 * 1024, 100   # BASELINE.json configs[1]
SMALL_PARTICLES, SMALL_FRAMES, SMALL_LOGS = 4096, 10000, 8   # BASELINE.json configs[4]
BPP = 40                             # bytes per particle
REF = os.path.join(REPO, "oracle", "_ref")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.t = threading.Thread(target=self._pump, daemon=True)
        self.t.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, windows):
        sm, mx, reasons = [], 0.0, set()
        for t, line in self.rows:
            if not any(a <= t <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ data
def rank_rows(n, nprocs, rank):
    from pgsd_sph_b200 import synth
    rows = synth.split_rows(n, nprocs)
    return rows, synth.row_starts(rows)[rank]


GEN_BLOCK = 1 << 20


def make_soa(n_total, start, n, seed):
    """Rows [start, start+n) of a synthetic n_total-particle frame as 10 SoA columns (pos xyz, vel xyz, density,
    pressure: f32; typeid, id: u32).  ids = slice of one seeded permutation of 0..n_total-1 (dense, unique,
    unsorted).  Every value is a function of (seed, global row) only -- generated in blocks of 1 Mi rows -- so any
    partition of the rows over ranks sees the same frame (the parity leg compares files written at different P)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ids = rng.permutation(n_total).astype(np.uint32)[start:start + n].copy()
    cols = [np.empty(n, dtype=np.float32) for _ in range(8)] + [np.empty(n, dtype=np.uint32)]
    for b in range(start // GEN_BLOCK, (start + n + GEN_BLOCK - 1) // GEN_BLOCK if n else 0):
        lo, hi = b * GEN_BLOCK, min((b + 1) * GEN_BLOCK, n_total)
        m = hi - lo
        rng = np.random.Generator(np.random.PCG64(seed * 1000003 + b + 1))
        blk = [(rng.random(m, dtype=np.float32) * np.float32(10.0)) for _ in range(3)]
        blk += [(rng.random(m, dtype=np.float32) - np.float32(0.5)) for _ in range(3)]
        dens = np.float32(1000.0) * (np.float32(1.0) + np.float32(0.01) * (rng.random(m, dtype=np.float32) - np.float32(0.5)))
        blk.append(dens.astype(np.float32))
        blk.append((np.float32(2.25) * (dens - np.float32(1000.0))).astype(np.float32))
        blk.append(rng.integers(0, 3, size=m, dtype=np.uint32))
        a, e = max(lo, start), min(hi, start + n)
        for c, src in zip(cols, blk):
            c[a - start:e - start] = src[a - lo:e - lo]
    cols.append(ids)
    return cols


SOA_CHUNKS = (  # name, column indices, dtype
    ("particles/position", (0, 1, 2), np.float32),
    ("particles/velocity", (3, 4, 5), np.float32),
    ("particles/typeid", (8,), np.uint32),
    ("particles/density", (6,), np.float32),
    ("particles/pressure", (7,), np.float32),
    ("log/particles/id", (9,), np.uint32),
)


def bench_dir(base=None):
    d = base or os.environ.get("PGSD_BENCH_DIR")
    if not d:
        d = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    d = os.path.join(d, "pgsd_bench_%d" % os.getuid())
    os.makedirs(d, exist_ok=True)
    return d


def fs_kind(path):
    try:
        best = ("", "?")
        for line in open("/proc/mounts"):
            f = line.split()
            if path.startswith(f[1]) and len(f[1]) > len(best[0]):
                best = (f[1], f[2])
        return best[1]
    except OSError:
        return "?"


def disk_dir():
    """A disk-backed second target for the write leg: PGSD_BENCH_DIR2, else /tmp when it is not tmpfs and has room."""
    d = os.environ.get("PGSD_BENCH_DIR2")
    if d == "":
        return None
    if not d:
        d = "/tmp"
        if fs_kind(d) in ("tmpfs", "ramfs", "?") or fs_kind(d) == fs_kind(bench_dir()) == "tmpfs":
            return None
    try:
        st = os.statvfs(d)
        if st.f_bavail * st.f_frsize < 24 << 30:
            return None
    except OSError:
        return None
    return d


def sha256_file(path):
    h = hashlib.sha256()
    with open(path, "rb", buffering=0) as f:
        while True:
            b = f.read(16 << 20)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


# ------------------------------------------------------------------------------------ this repo's arm
class Dist:
    def __init__(self, want):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.td = None
        if self.world > 1:
            import torch
            import torch.distributed as td
            torch.cuda.set_device(self.local)
            td.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.td, self.torch = td, torch
        if want != self.world:
            log(f"bench.py: --gpus {want} but WORLD_SIZE={self.world}; launch with torchrun for N > 1")

    def barrier(self):
        if self.td:
            self.td.barrier()

    def max(self, x):
        if not self.td:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        if not self.td:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.td.all_reduce(t, op=self.td.ReduceOp.SUM)
        return float(t.item())

    def bcast_bytes(self, b, n):
        if not self.td:
            return b
        t = self.torch.zeros(n, dtype=self.torch.uint8, device="cuda")
        if self.rank == 0:
            t.copy_(self.torch.frombuffer(bytearray(b), dtype=self.torch.uint8))
        self.td.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    def close(self):
        if self.td:
            self.td.barrier()
            self.td.destroy_process_group()


class Timer:
    def __init__(self, lib):
        self.lib, self.t = lib, C.c_void_p()
        lib.pgsd_b200_timer_create(C.byref(self.t))

    def start(self):
        self.lib.pgsd_b200_timer_start(self.t)

    def stop(self):
        ms = C.c_float()
        self.lib.pgsd_b200_timer_stop(self.t, C.byref(ms))
        return float(ms.value)


def get_stats(lib):
    from pgsd_sph_b200 import _lib
    st = _lib.Stats()
    lib.pgsd_b200_get_stats(st)
    return st


def committed_traffic(name, n):
    """DRAM bytes of one launch / one call from the committed ncu capture (profiles/<name>), scaled to this
    particle count; None if the capture is missing."""
    p = os.path.join(REPO, "profiles", name)
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return int(d["traffic_bytes"] * (n / d["particles"]))


def run_read_leg(lib, dist, args, peaks, windows):
    """ID-reordered read, frame-parallel: this rank's own frames of a 16 Mi-particle trajectory."""
    from pgsd_sph_b200 import _lib, fl, hoomd, synth, vtu
    from pgsd_sph_b200.devmem import DeviceArray
    n = args.read_particles
    path = os.path.join(bench_dir(), f"read_r{dist.rank}.gsd")
    nframes = 7   # frame 0 + 6 frames read round-robin (sequential access: the reader prefetches i+1)
    t0 = time.perf_counter()
    frames = []
    with fl.open(path, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
        for i in range(nframes):
            cols = make_soa(n, 0, n, 7000 + 10 * dist.rank + i)
            for k, a in synth.frame_scalars(n, i):
                f.write_chunk(k, a, write_all=False)
            prep = f.prepare_frame_soa([(nm, [cols[j] for j in idx], dt, None, True) for nm, idx, dt in SOA_CHUNKS])
            f.write_frame_soa(prep)
            f.end_frame()
            if i == 1:
                frames = cols
    log(f"[rank {dist.rank}] read leg: wrote {nframes} x {n} particles in {time.perf_counter() - t0:.1f}s -> {path}")

    # ---- value: decoded frame resident in HBM -> K4 sort + K5 gather (CUDA events on the stream)
    cols = frames
    pos = np.ascontiguousarray(np.stack(cols[0:3], axis=1))
    vel = np.ascontiguousarray(np.stack(cols[3:6], axis=1))
    host_fields = [pos, vel, cols[8], cols[6], cols[7]]
    d_in = [DeviceArray.from_numpy(a) for a in host_fields]
    d_out = [DeviceArray(a.shape, a.dtype) for a in host_fields]
    d_ids = DeviceArray.from_numpy(cols[9])
    d_sorted = DeviceArray((n,), np.uint32)
    d_perm = DeviceArray((n,), np.uint32)
    fields = (_lib.Field * 5)(*[_lib.Field(i.ptr, o.ptr, a.dtype.itemsize * (a.shape[1] if a.ndim > 1 else 1))
                                for i, o, a in zip(d_in, d_out, host_fields)])

    def step_dev(perm=None):
        # the call pgsd.hoomd makes (hoomd.py reorder_by_id): ids + 5 fields, no permutation output
        _lib.check(lib.pgsd_b200_reorder_device(n, d_ids.ptr, d_sorted.ptr, perm, 5, fields, None), "reorder")

    for _ in range(args.warmup):
        step_dev()
    lib.pgsd_b200_synchronize()
    # per-phase device times of the real path (CUDA events recorded by the library on the stream)
    tm = Timer(lib)
    lib.pgsd_b200_reorder_profiling(1)
    phases = []
    for _ in range(3):
        step_dev()
        ms = (C.c_float * 4)()
        lib.pgsd_b200_reorder_phase_ms(ms)
        phases.append([float(x) for x in ms])
    lib.pgsd_b200_reorder_profiling(0)
    phase_ms = [min(p[i] for p in phases) for i in range(4)]
    lib.pgsd_b200_reset_stats()
    dist.barrier()
    lib.pgsd_b200_synchronize()
    w0 = time.perf_counter()
    tm.start()
    for _ in range(args.steps):
        step_dev()
    dev_ms = tm.stop()
    lib.pgsd_b200_synchronize()
    dist.barrier()
    windows.append((w0, time.perf_counter()))
    launches = get_stats(lib).kernel_launches
    dev_ms = dist.max(dev_ms)
    value = dist.world * n * args.steps / (dev_ms * 1e-3) / 1e6

    # parity spot check of the timed configuration: sortedness + id->row consistency (size independent)
    got_ids = d_sorted.to_numpy()
    assert np.array_equal(got_ids, np.arange(n, dtype=np.uint32)), "reordered ids are not 0..N-1"
    got_pos = d_out[0].to_numpy()
    step_dev(d_perm.ptr)   # untimed: same reorder, this time also returning the permutation for the checks below
    lib.pgsd_b200_synchronize()
    assert np.array_equal(d_sorted.to_numpy(), got_ids) and np.array_equal(d_out[0].to_numpy(), got_pos)
    del got_pos
    perm = d_perm.to_numpy()
    assert np.array_equal(cols[9][perm], got_ids), "perm does not sort the ids"
    chk = np.random.default_rng(1).integers(0, n, size=4096)
    assert np.array_equal(d_out[0].to_numpy()[chk], pos[perm[chk]]), "gathered positions differ"
    del got_ids, perm

    # ---- e2e: file -> HOOMDTrajectory(reorder='id') -> host numpy arrays, every step
    for a in d_in + d_out + [d_ids, d_sorted, d_perm]:
        a.free()
    traj = hoomd.open(path, 'r', reorder='id')
    fr = traj[0]

    def step_e2e(i):
        fr = traj[1 + (i % (nframes - 1))]
        p = fr.particles
        return p.position, p.velocity, p.typeid, p.density, p.pressure, fr.log['particles/id']

    out = None
    for i in range(args.warmup):
        out = step_e2e(i)  # held like in the timed loop: two generations of pooled pinned arrays
    lib.pgsd_b200_reset_stats()
    dist.barrier()
    lib.pgsd_b200_synchronize()
    w0 = time.perf_counter()
    step_t = []
    for i in range(args.steps):
        ts = time.perf_counter()
        out = step_e2e(i)
        step_t.append(time.perf_counter() - ts)
    lib.pgsd_b200_synchronize()
    t_e2e = time.perf_counter() - w0
    log(f"[rank {dist.rank}] read e2e per-step ms: " + " ".join(f"{1e3 * x:.1f}" for x in step_t))
    dist.barrier()
    windows.append((w0, time.perf_counter()))
    assert out[5][0] == 0 and out[5][-1] == n - 1
    st = get_stats(lib)
    t_e2e = dist.max(t_e2e)
    e2e = dist.world * n * args.steps / t_e2e / 1e6

    # ---- pgsd2vtu leg (config 4's consumer): file -> reordered frame on the device -> column split + f64 cast
    # (K1's strided path) -> .vtu.  The VTU container has no reference implementation here (pyevtk is absent and
    # unpinned): container parity unpinned; the array preparation is pinned in tests/test_python_layer.py.
    vtu_obj = None
    if not args.no_vtu:
        vdir = bench_dir()
        traj.close()   # joins its frame-prefetch thread: the device-mode trajectory below reads the same file
        traj = None
        traj_dev = hoomd.open(path, 'r', reorder='id', device=True)
        tv = []
        vsteps = max(2, min(args.steps, 3))
        for i in range(vsteps + 1):
            t1 = time.perf_counter()
            fr = traj_dev[1 + (i % (nframes - 1))]
            x, y, z, pd = vtu.point_arrays(fr)
            t2 = time.perf_counter()
            vp = vtu.write_vtu(os.path.join(vdir, f"vtu_r{dist.rank}"), x, y, z, pd)
            t3 = time.perf_counter()
            if i > 0:
                tv.append((t2 - t1, t3 - t2))
            vbytes = os.path.getsize(vp)
            os.unlink(vp)
        traj_dev.close()
        prep_s, enc_s = float(np.mean([a for a, _ in tv])), float(np.mean([b for _, b in tv]))
        vtu_obj = {"metric": "pgsd2vtu_Mparticles_per_s", "value": dist.world * n / (dist.max(prep_s + enc_s)) / 1e6,
                   "unit": "Mparticles/s", "prepare_s_per_frame": prep_s, "encode_write_s_per_frame": enc_s,
                   "vtu_bytes": vbytes,
                   "path": "file -> reordered frame in HBM -> K1 column split / f64 cast / xyz interleave -> D2H into a "
                           "page-locked file image -> file stage (8 threads) -> .vtu",
                   "parity": "array preparation pinned vs numpy; container laid out as pyevtk's pointsToVTK writes it "
                             "(oracle/vtu_oracle.py, restated from its source) but UNPINNED: pyevtk is absent, the "
                             "reference names no version and ships no output"}
    if traj is not None:
        traj.close()
    os.unlink(path)

    peak = peaks["hbm_gbs"]

    def kern(ms, bytes_pp, note):
        t = max(ms, 1e-6) * 1e-3
        return {"ms": ms, "algorithmic_bytes": bytes_pp * n, "GBps": bytes_pp * n / t / 1e9,
                "frac": bytes_pp * n / t / 1e9 / peak, "note": note}

    slot = os.environ.get("PGSD_B200_SLOT", "1") != "0" and phase_ms[2] < 0.02   # unique ids: no pair passes ran
    cluster = os.environ.get("PGSD_B200_CLUSTER", "0") == "1"
    if slot and cluster:
        kernels = {
            "k7_coarse_scatter": kern(phase_ms[1], 2 * 40, "rows moved once into coarse buckets as contiguous runs"),
            "k7_cluster_place": kern(phase_ms[3], 2 * 40, "clusters of 8 CTAs: records to their owner, slot order, fields out"),
        }
    elif slot:
        kernels = {
            "k6_slot_hist+scan+scatter": kern(phase_ms[1], 4 + 2 * 40, "bucket histogram (4 B) + rows moved once into "
                                              "interleaved records: 40 B read + 40 B written (one cursor atomic per row)"),
            "k6_slot_place": kern(phase_ms[3], 2 * 40, "one CTA per bucket: 40 B records read, 40 B of fields written"),
        }
        if phase_ms[0] > 0.005:   # the key range guessed from n was wrong: the census measured it
            kernels["k4_digit_census"] = kern(phase_ms[0], 4, "keys read once (OR/AND); includes the 8-byte D2H + host sync")
    else:
        kernels = {
            "k4_digit_census": kern(phase_ms[0], 4, "keys read once; includes the 8 KB D2H + host sync"),
            "k4_bucket_rows": kern(phase_ms[1], 4 + 2 * 40, "tile histogram (4 B) + rows moved once: 40 B read + 40 B written"),
            "k4_pair_passes": kern(phase_ms[2], 2 * 4 + (4 + 8) + 16, "2 segmented LSD passes inside buckets: histogram 4 B each; (key,idx) 12 + 16 B"),
            "k5_gather": kern(phase_ms[3], 4 + 2 * 36, "perm 4 B + 36 B payload read + 36 B written"),
        }
    reorder_s = dev_ms * 1e-3 / args.steps
    out = {
        "metric": "id_reordered_read_Mparticles_per_s", "value": value, "unit": "Mparticles/s",
        "ms_per_step": dev_ms / args.steps, "scaling": "weak", "dtype": "u32",
        "config": {"workload": f"config 4: {n}-particle unsorted frames, 40 B/particle, one frame per step per GPU "
                               "(frame-parallel, no collective)", "l2": "inputs (671 MB) larger than L2"},
        "e2e": {"value": e2e, "unit": "Mparticles/s", "h2d_bytes_per_step": st.h2d_bytes // args.steps,
                "d2h_bytes_per_step": st.d2h_bytes // args.steps,
                "path": "file -> pgsd.hoomd.HOOMDTrajectory(reorder='id')[i] -> host numpy arrays"},
        "roofline": {"bound": "hbm", "achieved": 80 * n / reorder_s / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": 80 * n / reorder_s / 1e9 / peak,
                     "traffic": committed_traffic("reorder_traffic.json", n) if slot and not cluster else None,
                     "peak_source": peaks["source"],
                     "what": "K4+K5 reorder as one operation: 80 B/particle algorithmic"},
        "kernels": kernels,
        "reorder_path": ("cluster (unique ids)" if cluster else "slot (unique ids)") if slot else "general (stable LSD)",
        "gpu_launches": int(launches),
    }
    if vtu_obj:
        out["vtu"] = vtu_obj
    return out


def run_dist_reorder_leg(lib, dist, args, peaks, windows):
    """Config 3 read-back: ONE args.particles-particle frame whose rows are partitioned over the ranks (rank r holds
    file partition r, as after pgsd_read_chunk(all=true)), put into particle-id order across the GPUs by
    pgsd_b200_reorder_distributed: records go straight into the owner's HBM over NVLink, owners finish locally."""
    from pgsd_sph_b200 import _lib
    from pgsd_sph_b200.devmem import DeviceArray
    n_total = args.particles
    all_rows, start = rank_rows(n_total, dist.world, dist.rank)
    rows = all_rows[dist.rank]
    cols = make_soa(n_total, start, rows, 4242)
    ids = cols[9]
    cols[8] = (ids ^ np.uint32(0x9e3779b9)).astype(np.uint32)      # typeid column carries a function of the id: checkable
    pos = np.ascontiguousarray(np.stack(cols[0:3], axis=1))
    vel = np.ascontiguousarray(np.stack(cols[3:6], axis=1))
    host_fields = [pos, vel, cols[8], cols[6], cols[7]]
    first, cap = C.c_uint64(), C.c_uint64()
    _lib.check(lib.pgsd_b200_reorder_distributed_plan(n_total, dist.world, dist.rank, C.byref(first), C.byref(cap)), "plan")
    cap = int(cap.value)
    d_in = [DeviceArray.from_numpy(a) for a in host_fields]
    d_out = [DeviceArray((cap,) + a.shape[1:], a.dtype) for a in host_fields]
    d_ids = DeviceArray.from_numpy(ids)
    d_sorted = DeviceArray((cap,), np.uint32)
    fields = (_lib.Field * 5)(*[_lib.Field(i.ptr, o.ptr, a.dtype.itemsize * (a.shape[1] if a.ndim > 1 else 1))
                                for i, o, a in zip(d_in, d_out, host_fields)])
    n_out, id_first = C.c_uint64(), C.c_uint64()

    def step():
        rc = lib.pgsd_b200_reorder_distributed(rows, d_ids.ptr, cap, C.byref(n_out), C.byref(id_first), d_sorted.ptr, 5, fields, None)
        if rc != 0:
            raise RuntimeError(f"reorder_distributed rc={rc}: {lib.pgsd_b200_last_error().decode()}")

    for _ in range(args.warmup):
        step()
    lib.pgsd_b200_reset_stats()
    tm = Timer(lib)
    dist.barrier()
    lib.pgsd_b200_synchronize()
    w0 = time.perf_counter()
    tm.start()
    for _ in range(args.steps):
        step()
    dev_ms = tm.stop()
    wall = time.perf_counter() - w0
    dist.barrier()
    windows.append((w0, time.perf_counter()))
    st = get_stats(lib)
    dev_ms = dist.max(dev_ms)
    wall = dist.max(wall)
    k, f0 = int(n_out.value), int(id_first.value)
    got = d_sorted.to_numpy()[:k]
    assert np.array_equal(got, np.arange(f0, f0 + k, dtype=np.uint32)), "owned ids are not consecutive"
    assert np.array_equal(d_out[2].to_numpy()[:k], got ^ np.uint32(0x9e3779b9)), "payload does not follow its id"
    assert int(dist.sum(float(k))) == n_total
    for a in d_in + d_out + [d_ids, d_sorted]:
        a.free()
    value = n_total * args.steps / (dev_ms * 1e-3) / 1e6
    coll = st.collectives // max(args.steps, 1)
    return {
        "metric": "distributed_id_reorder_Mparticles_per_s", "value": value, "unit": "Mparticles/s",
        "ms_per_step": dev_ms / args.steps, "wall_ms_per_step": 1e3 * wall / args.steps, "scaling": "strong",
        "config": {"workload": f"config 3 read-back: one {n_total}-particle frame, rows partitioned over {dist.world} rank(s), "
                               "40 B/particle, ids dense and unsorted; fields resident in HBM on both sides",
                   "comm": lib.pgsd_b200_comm_kind().decode(),
                   "timed": f"whole collective call ({coll} host-visible collective(s) per call, histogram, records over "
                            "NVLink, placement), CUDA events on the stream, max over ranks"},
        "rows_owned_rank0": k, "gpu_launches": int(dist.sum(float(st.kernel_launches))),
        "collectives_per_step": coll,
    }


def file_ceiling(lib, dist, target_dir, total_bytes):
    """Host-only ceiling of the file stage on `target_dir`: every rank writes its share of `total_bytes` of host
    memory into ONE new file with the library's own piece size, thread counts and mode, all ranks at once."""
    path = os.path.join(target_dir, "ceiling.bin")
    share = total_bytes // dist.world
    s, thr, mapped = C.c_double(), C.c_int(), C.c_int()
    dist.barrier()
    rc = lib.pgsd_b200_file_stage_ceiling(path.encode(), 4096 + 1234 + dist.rank * share, share, C.byref(s), C.byref(thr), C.byref(mapped))
    t = dist.max(s.value if rc == 0 else 1e9)
    dist.barrier()
    if dist.rank == 0 and os.path.exists(path):
        os.unlink(path)
    return {"GBps": share * dist.world / t / 1e9, "threads_per_rank": int(thr.value),
            "mode": "shared mappings (tmpfs)" if mapped.value else "pwrite", "bytes": share * dist.world}


def run_write_leg(lib, dist, args, peaks, windows, target_dir=None, steps=None, warmup=None, legs=("device", "e2e")):
    from pgsd_sph_b200 import _lib, fl, synth
    from pgsd_sph_b200.devmem import DeviceArray, PinnedArray
    steps = steps or args.steps
    warmup = warmup or args.warmup
    tdir = bench_dir(target_dir)
    n_total = args.particles
    rows, start = rank_rows(n_total, dist.world, dist.rank)
    n = rows[dist.rank]
    t0 = time.perf_counter()
    cols = make_soa(n_total, start, n, 20261018)
    log(f"[rank {dist.rank}] write leg ({tdir}): generated {n} of {n_total} particles in {time.perf_counter() - t0:.1f}s")
    d_cols = [DeviceArray.from_numpy(a) for a in cols] if "device" in legs else None
    h_cols = None
    if "e2e" in legs:
        pinned = [PinnedArray(a.shape, a.dtype) for a in cols]
        for p, a in zip(pinned, cols):
            p.array[...] = a
        h_cols = [p.array for p in pinned]
    payload = BPP * n_total
    path = os.path.join(tdir, "write_A.gsd")
    result = {}
    for leg in legs:
        src = d_cols if leg == "device" else h_cols
        dist.barrier()
        f = fl.open(path, 'w', 'pgsd-b200', 'hoomd', [1, 4])
        prep = f.prepare_frame_soa([(nm, [src[j] for j in idx], dt, rows, True) for nm, idx, dt in SOA_CHUNKS],
                                   rank=dist.rank)
        k1_ms = []
        lib.pgsd_b200_pack_profiling(1 if leg == "device" else 0)

        def step(i, timed=False):
            for k, a in synth.frame_scalars(n_total, i):
                f.write_chunk(k, a, write_all=False)
            f.write_frame_soa(prep)
            if timed and leg == "device":
                ms = C.c_float()
                if lib.pgsd_b200_pack_last_ms(C.byref(ms)) == 0:  # CUDA events around the K1 launch
                    k1_ms.append(float(ms.value))
            f.end_frame()

        # the trajectory file must fit the target: start a new file when it would pass the budget
        # (only matters for very long runs; closing + reopening is inside the timed region then)
        st_fs = os.statvfs(tdir)
        budget = min(64 << 30, int(0.35 * st_fs.f_bavail * st_fs.f_frsize))
        frames_per_file = max(1, budget // max(payload, 1))
        in_file = 0

        def roll():
            nonlocal f, prep, in_file, path
            if in_file < frames_per_file:
                return
            f.close()
            dist.barrier()
            old = path
            path = old[:-5] + ("B.gsd" if old.endswith("A.gsd") else "A.gsd")
            if dist.rank == 0:  # freeing tens of GB of page cache takes seconds: off the critical path
                threading.Thread(target=os.unlink, args=(old,), daemon=False).start()
            f = fl.open(path, 'w', 'pgsd-b200', 'hoomd', [1, 4])
            prep = f.prepare_frame_soa([(nm, [src[j] for j in idx], dt, rows, True) for nm, idx, dt in SOA_CHUNKS],
                                       rank=dist.rank)
            in_file = 0

        for i in range(warmup):
            roll()
            step(i)
            in_file += 1
        f.flush()
        lib.pgsd_b200_reset_stats()
        dist.barrier()
        lib.pgsd_b200_synchronize()
        w0 = time.perf_counter()
        for i in range(steps):
            roll()
            step(warmup + i, timed=True)
            in_file += 1
        f.flush()  # drains every queued D2H + file write of this rank
        lib.pgsd_b200_synchronize()
        dt = time.perf_counter() - w0
        dist.barrier()
        windows.append((w0, time.perf_counter()))
        st = get_stats(lib)
        dt = dist.max(dt)
        lib.pgsd_b200_pack_profiling(0)
        f.close()
        size = os.path.getsize(path) if dist.rank == 0 else 0
        result[leg] = {"s": dt, "GBps": payload * steps / dt / 1e9, "stats": st, "k1_ms": k1_ms,
                       "file_bytes": size}
        dist.barrier()
        if dist.rank == 0:
            os.unlink(path)
    if d_cols:
        for a in d_cols:
            a.free()
    ceil = file_ceiling(lib, dist, tdir, 2 * payload if payload < (8 << 30) else payload)

    first = result[legs[0]]
    out = {"metric": "frame_write_GBps", "value": first["GBps"], "unit": "GB/s", "ms_per_step": first["s"] / steps * 1e3,
           "file_target": f"{tdir} ({fs_kind(tdir)}; no fsync, as the reference)"}
    peak = peaks["hbm_gbs"]
    if "device" in result:
        dev = result["device"]
        k1_s = dist.max(float(np.median(dev["k1_ms"])) * 1e-3)   # K1 alone, CUDA events around its launch on the stream
        k1_bytes = 80 * n  # read 40 + write 40 B/particle, this rank's launch
        st = dev["stats"]
        writers = max(1, ceil["threads_per_rank"])
        out["roofline"] = {"bound": "hbm", "achieved": k1_bytes / k1_s / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": k1_bytes / k1_s / 1e9 / peak, "traffic": committed_traffic("k1_traffic.json", n),
                           "peak_source": peaks["source"], "kernel": "k1_pack_frame", "ms": k1_s * 1e3,
                           "algorithmic_bytes": k1_bytes}
        out["gpu_launches"] = int(dist.sum(float(st.kernel_launches)))
        out["split"] = {
            "device_k1_ms_per_frame": k1_s * 1e3,
            "d2h_s_per_frame": dist.max(st.d2h_busy_s / steps / 2.0),
            "d2h_note": "sum over the frame's pieces of their copy time (CUDA events on the 2 copy streams) / 2 streams; "
                        "max over ranks",
            "d2h_GBps_per_rank": (st.d2h_bytes / max(st.d2h_busy_s / 2.0, 1e-9)) / 1e9,
            "file_s_per_frame": dist.max(st.file_busy_s / steps / writers),
            "file_note": f"writer-thread time inside the file write, summed over pieces / {writers} threads at work; max over ranks",
            "file_busy_thread_s_per_frame": st.file_busy_s / steps,
            "pieces_per_frame_rank0": st.pieces // steps,
            "commit_wait_s_per_frame": st.commit_wait_s / steps,
            "d2h_bytes_per_step": int(dist.sum(float(st.d2h_bytes))) // steps,
            "file_bytes": dev["file_bytes"],
            "file_ceiling_GBps": ceil["GBps"], "file_ceiling": ceil,
            "value_over_file_ceiling": dev["GBps"] / ceil["GBps"],
            "note": "wall per frame = max(K1, D2H over PCIe, file stage); the file stage is the ceiling: host memory -> "
                    "page cache of ONE file on this host, measured in this run without any GPU work (file_ceiling)",
        }
    if "e2e" in result:
        e2e = result["e2e"]
        out["e2e"] = {"value": e2e["GBps"], "unit": "GB/s",
                      "h2d_bytes_per_step": int(dist.sum(float(e2e["stats"].h2d_bytes))) // steps,
                      "d2h_bytes_per_step": int(dist.sum(float(e2e["stats"].d2h_bytes))) // steps,
                      "path": "pinned host SoA columns -> pgsd_b200_write_chunks_soa (H2D, K1, D2H) -> file stage -> file"}
    return out


def write_one_frame(dist, path, n_total, cols, rows, frame=0, nlogs=0):
    from pgsd_sph_b200 import fl, synth
    from pgsd_sph_b200.devmem import DeviceArray
    d = [DeviceArray.from_numpy(a) for a in cols]
    with fl.open(path, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
        for k, a in synth.frame_scalars(n_total, frame):
            f.write_chunk(k, a, write_all=False)
        prep = f.prepare_frame_soa([(nm, [d[j] for j in idx], dt, rows, True) for nm, idx, dt in SOA_CHUNKS], rank=dist.rank)
        f.write_frame_soa(prep)
        for k in range(nlogs):
            f.write_chunk("log/value/v%d" % k, np.array([k], dtype=np.float32), write_all=False)
        f.end_frame()
    for a in d:
        a.free()


def ref_driver_bench(out, n, frames, blob, nranks, nlogs=0):
    """oracle/_ref/ref_driver: the UNMODIFIED reference pgsd.c writing `frames` frames of the blob's particles."""
    drv = os.path.join(REF, "ref_driver")
    if not os.path.exists(drv):
        return None
    r = subprocess.run([drv, "bench", out, str(n), str(frames), blob, "nofsync", str(nlogs)],
                       env=dict(os.environ, PGSD_SHIM_NP=str(nranks)), capture_output=True, text=True)
    if r.returncode != 0:
        log("ref_driver failed:", r.stderr[-400:])
        return None
    return json.loads(r.stdout.strip().splitlines()[-1])["frame_s"]


def run_parity_leg(lib, dist, args):
    """north_star's correctness target inside the driver-run record: one config-3 frame written by this library at
    P = N ranks under the run's communicator (NCCL over NVLink for N > 1; device-resident columns -> K1 -> K2 -> K3)
    and by the unmodified reference pgsd.c at PGSD_SHIM_NP = N from the same inputs.  Reference rule for the bytes:
    pgsd.c:2225-2247 (per-rank MPI_File_write_at at file_size + offset, SUM accounting), :1150-1154 (small chunks
    replicated per rank)."""
    n_total = args.particles
    rows, start = rank_rows(n_total, dist.world, dist.rank)
    n = rows[dist.rank]
    d = bench_dir()
    mine, ref, blob = os.path.join(d, "parity_b200.gsd"), os.path.join(d, "parity_ref.gsd"), os.path.join(d, "parity_blob.bin")
    cols = make_soa(n_total, start, n, 20261018)
    dist.barrier()
    write_one_frame(dist, mine, n_total, cols, rows)
    # every rank drops its rows into the reference's input blob (SoA columns of the whole frame)
    if dist.rank == 0:
        with open(blob, "wb") as fh:
            fh.truncate(BPP * n_total)
    dist.barrier()
    fd = os.open(blob, os.O_RDWR)
    for j, c in enumerate(cols):
        os.pwrite(fd, c.tobytes(), (j * n_total + start) * 4)
    os.close(fd)
    del cols
    dist.barrier()
    out = None
    if dist.rank == 0:
        t = ref_driver_bench(ref, n_total, 1, blob, dist.world)
        if t is None:
            out = {"P": dist.world, "sha256_equal": None, "note": "oracle/_ref/ref_driver is not built"}
        else:
            res = {}
            th = [threading.Thread(target=lambda k, p: res.__setitem__(k, sha256_file(p)), args=a) for a in (("b200", mine), ("ref", ref))]
            for x in th:
                x.start()
            for x in th:
                x.join()
            out = {"P": dist.world, "comm": lib.pgsd_b200_comm_kind().decode(), "particles": n_total,
                   "sha256_equal": res["b200"] == res["ref"], "bytes": os.path.getsize(mine),
                   "bytes_reference": os.path.getsize(ref), "sha256": res["b200"],
                   "reference": f"unmodified pgsd.c + MPI shim at {dist.world} rank(s), same inputs"}
        for p in (mine, ref, blob):
            if os.path.exists(p):
                os.unlink(p)
    dist.barrier()
    return out


def run_frames_leg(lib, dist, args, windows, n_total, frames, nlogs, tag):
    """config 2 (trajectory: 100 frames x 1 Mi particles) and config 5 (10 k frames x 4096 particles + 8 log
    scalars): device-resident SoA fields -> file, `frames` frames back to back, then host-resident ones."""
    from pgsd_sph_b200 import fl, synth
    from pgsd_sph_b200.devmem import DeviceArray
    rows, start = rank_rows(n_total, dist.world, dist.rank)
    n = rows[dist.rank]
    cols = make_soa(n_total, start, n, 5)
    d = [DeviceArray.from_numpy(c) for c in cols]
    path = os.path.join(bench_dir(), tag + ".gsd")
    logs = [("log/value/v%d" % k, np.array([k], dtype=np.float32)) for k in range(nlogs)]
    res = {}
    for mode in ("device", "host"):
        src = d if mode == "device" else cols
        dist.barrier()
        with fl.open(path, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
            prep = f.prepare_frame_soa([(nm, [src[j] for j in idx], dt, rows, True) for nm, idx, dt in SOA_CHUNKS],
                                       rank=dist.rank)
            # the per-frame scalars live in buffers that are updated in place (what a simulation loop does): their
            # pgsd_write_chunk arguments are validated once (PGSDFile.prepare_chunks)
            scal = synth.frame_scalars(n_total, 0)
            step = scal[0][1]
            head = f.prepare_chunks([(k, a, None, False) for k, a in scal])
            tail = f.prepare_chunks([(k, a, None, False) for k, a in logs])
            lib.pgsd_b200_reset_stats()
            dist.barrier()
            w0 = time.perf_counter()
            for i in range(frames):
                step[0] = 10 * i
                f.write_prepared(head)
                f.write_frame_soa(prep)
                f.write_prepared(tail)
                f.end_frame()
            f.flush()
            dt_ = time.perf_counter() - w0
            dist.barrier()
            windows.append((w0, time.perf_counter()))
            st = get_stats(lib)
            dt_ = dist.max(dt_)
        dist.barrier()
        size = 0
        if dist.rank == 0:
            size = os.path.getsize(path)
            os.unlink(path)
        res[mode] = {"s": dt_, "us_per_frame": 1e6 * dt_ / frames, "GBps": BPP * n_total * frames / dt_ / 1e9,
                     "file_bytes": size, "launches": int(dist.sum(float(st.kernel_launches))),
                     "collectives_per_frame": st.collectives / frames}
    for a in d:
        a.free()
    return res


def cpu_frames_reference(n_total, frames, nranks, nlogs):
    d = bench_dir()
    blob, out = os.path.join(d, "ref_blob_small.bin"), os.path.join(d, "ref_frames.gsd")
    cols = make_soa(n_total, 0, n_total, 5)
    with open(blob, "wb") as fh:
        for c in cols:
            fh.write(c.tobytes())
    t = ref_driver_bench(out, n_total, frames, blob, nranks, nlogs)
    for p in (blob, out):
        if os.path.exists(p):
            os.unlink(p)
    if t is None:
        return None
    return {"us_per_frame": 1e6 * sum(t) / len(t), "GBps": BPP * n_total * len(t) / sum(t) / 1e9, "frames": len(t), "ranks": nranks}


def run_benchmark_write_leg(lib, dist, args, keep=False):
    """The reference's own published benchmark (scripts/benchmark-write.cc:30-45,85-160; CHANGELOG.md:172-194):
    17 keys x 100 frames x 1 Mi float64 per key, rows split over the ranks, every key written with
    all=true at the caller-computed offset; throughput = MiB written in the SECOND 50 frames / their time.
    Here the keys are device-resident arrays handed to pgsd_write_chunk as device pointers."""
    from pgsd_sph_b200 import fl
    from pgsd_sph_b200.devmem import DeviceArray
    nkeys, nframes, n_total = 17, 100, 1024 * 1024
    rows, start = rank_rows(n_total, dist.world, dist.rank)
    n = rows[dist.rank]
    rng = np.random.default_rng(11 + dist.rank)
    keys = [DeviceArray.from_numpy(rng.standard_normal(n)) for _ in range(nkeys)]
    names = ["quantity/%d" % k for k in range(nkeys)]
    path = os.path.join(bench_dir(), "benchmark_write.gsd")
    dist.barrier()
    f = fl.open(path, 'w', 'pgsd-b200', 'benchmark', [1, 0])
    t1 = None
    for i in range(nframes):
        if i == nframes // 2:
            f.flush()
            lib.pgsd_b200_synchronize()
            dist.barrier()
            t1 = time.perf_counter()
        for nm, a in zip(names, keys):
            f.write_chunk(nm, a, offset=rows, rank=dist.rank)
        f.end_frame()
    f.flush()
    lib.pgsd_b200_synchronize()
    t2 = time.perf_counter() - t1
    dist.barrier()
    t2 = dist.max(t2)
    f.close()
    dist.barrier()
    if dist.rank == 0 and not keep:
        os.unlink(path)
    for a in keys:
        a.free()
    mib = (nframes - nframes // 2) * nkeys * n_total * 8 / 1048576.0
    return {"metric": "benchmark_write_MiBps", "value": mib / t2, "unit": "MiB/s", "seconds": t2,
            "workload": "17 keys x 100 frames x 1 Mi float64 (14.26 GB), second 50 frames timed, as benchmark-write.cc",
            "vs_baseline": mib / t2 / 167.0,
            "vs_baseline_note": "published 167.0 MiB/s at 1 rank on NVMe (CHANGELOG.md:186); this run writes to "
                                "tmpfs, so the ratio mixes implementation and storage"}, path


def run_benchmark_read_leg(lib, dist, args, path):
    """The reference's benchmark-read workload (scripts/benchmark-read.cc:46-120): every rank reads its row slice of
    every key of every frame (pgsd_read_chunk(all=true), pgsd.c:2497-2508) -- here straight into device memory
    (file -> pinned pieces -> H2D on reader threads).  SURVEY.md section 8(f) row 1."""
    from pgsd_sph_b200 import fl
    nkeys, nframes, n_total = 17, 100, 1024 * 1024
    rows, start = rank_rows(n_total, dist.world, dist.rank)
    n = rows[dist.rank]
    f = fl.open(path, 'r')
    names = ["quantity/%d" % k for k in range(nkeys)]
    assert f.nframes == nframes
    for nm in names[:2]:   # warm the reader threads and pinned buffers
        f.read_chunk(0, nm, N=n, M=1, offset=start, r_all=True, device=True).free()
    lib.pgsd_b200_reset_stats()
    import ctypes as C
    ah0 = [C.c_uint64() for _ in range(3)]
    lib.pgsd_b200_read_ahead_stats(*[C.byref(x) for x in ah0])
    ah0 = [x.value for x in ah0]
    dist.barrier()
    lib.pgsd_b200_synchronize()
    t0 = time.perf_counter()
    for i in range(nframes):
        for nm in names:
            a = f.read_chunk(i, nm, N=n, M=1, offset=start, r_all=True, device=True)
            a.free()
    lib.pgsd_b200_synchronize()
    t = time.perf_counter() - t0
    dist.barrier()
    st = get_stats(lib)
    t = dist.max(t)
    f.close()
    dist.barrier()
    if dist.rank == 0:
        os.unlink(path)
    gb = nkeys * nframes * n_total * 8 / 1e9
    import ctypes as C
    ah = [C.c_uint64() for _ in range(3)]
    lib.pgsd_b200_read_ahead_stats(*[C.byref(x) for x in ah])
    return {"metric": "partitioned_read_GBps", "value": gb / t, "unit": "GB/s", "seconds": t,
            "h2d_bytes": int(dist.sum(float(st.h2d_bytes))),
            "read_ahead": {"served_from_staging": ah[0].value - ah0[0], "ranges_fetched_ahead": ah[1].value - ah0[1],
                           "fetched_never_used": ah[2].value - ah0[2],
                           "note": "rank 0; with PGSD_B200_READ_AHEAD=1 (opt-in, off in this run unless set) equally sized reads "
                                   "at a constant file stride are fetched ahead into device staging: +5-8 % here"},
            "workload": "17 keys x 100 frames x 1 Mi float64 (14.26 GB), every rank reads its row slice of every key into "
                        "device memory: read_chunk(r_all=True, device=True), as benchmark-read.cc"}


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------ the reference's CPU path
def cpu_write_reference(n, frames, nranks, warm=1, target_dir=None):
    """The UNMODIFIED reference pgsd.c (oracle/_ref/ref_driver, MPI shim) writing the same frame."""
    if not os.path.exists(os.path.join(REF, "ref_driver")):
        return None
    d = bench_dir(target_dir)
    blob, out = os.path.join(bench_dir(), "ref_blob.bin"), os.path.join(d, "ref_write.gsd")
    cols = make_soa(n, 0, n, 20261018)
    with open(blob, "wb") as fh:
        for c in cols:
            fh.write(c.tobytes())
    del cols
    t = ref_driver_bench(out, n, frames + warm, blob, nranks)
    for p in (blob, out):
        if os.path.exists(p):
            os.unlink(p)
    if t is None:
        return None
    t = t[warm:]
    return {"s_per_frame": float(np.mean(t)), "GBps": BPP * n / float(np.mean(t)) / 1e9, "frames": len(t)}


def cpu_benchmark_reference(nranks, read=True):
    """The reference's OWN benchmark binaries (scripts/benchmark-write.cc / benchmark-read.cc compiled unmodified
    against the MPI shim): 17 keys x 100 frames x 1 Mi float64; benchmark-write prints "MB/s:" (MiB/s of the second
    50 frames), benchmark-read the seconds it took to read the file benchmark-write left behind."""
    exe = os.path.join(REF, "benchmark-write")
    if not os.path.exists(exe):
        return None, None
    d = bench_dir()
    env = dict(os.environ, PGSD_SHIM_NP=str(nranks))
    r = subprocess.run([exe], cwd=d, env=env, capture_output=True, text=True)
    out = os.path.join(d, "test%d.gsd" % nranks)
    bw = br = None
    for line in r.stdout.splitlines():
        if line.startswith("MB/s:"):
            bw = {"metric": "benchmark_write_MiBps", "value": float(line.split()[1]), "unit": "MiB/s", "ranks": nranks,
                  "workload": "17 keys x 100 frames x 1 Mi float64, unmodified benchmark-write.cc + pgsd.c + MPI shim",
                  "vs_baseline": float(line.split()[1]) / 167.0}
    if bw is None:
        log("benchmark-write failed:", r.stderr[-300:])
    rexe = os.path.join(REF, "benchmark-read")
    if read and bw and os.path.exists(rexe) and os.path.exists(out):
        r = subprocess.run([rexe], cwd=d, env=env, capture_output=True, text=True)
        for line in r.stdout.splitlines():
            if line.startswith("Total time required:"):
                secs = float(line.split()[3])
                gb = 17 * 100 * 1024 * 1024 * 8 / 1e9
                br = {"metric": "partitioned_read_GBps", "value": gb / secs, "unit": "GB/s", "seconds": secs, "ranks": nranks,
                      "workload": "unmodified benchmark-read.cc + pgsd.c + MPI shim reading benchmark-write's file into host memory"}
        if br is None:
            log("benchmark-read failed:", r.stderr[-300:], r.stdout[-300:])
    if os.path.exists(out):
        os.unlink(out)
    return bw, br


def cpu_read_reference(n, steps, warm=1):
    """The reference's reader timed on the host: file written by the unmodified reference pgsd.c (ref_driver), decoded
    by the reference's own pure-Python reader (pypgsd.py + hoomd.py staged unmodified in oracle/_ref/pyref by
    oracle/build_ref.sh; kind "reference"), then numpy stable argsort + gather -- the oracle-defined reorder.  Falls
    back to the oracle port of the reader (kind "port") when the staged modules are missing."""
    d = bench_dir()
    path, blob = os.path.join(d, "ref_read.gsd"), os.path.join(d, "ref_read_blob.bin")
    cols = make_soa(n, 0, n, 7000)
    with open(blob, "wb") as fh:
        for c in cols:
            fh.write(c.tobytes())
    del cols
    ok = ref_driver_bench(path, n, 2, blob, 1)
    os.unlink(blob)
    if ok is None:
        return None
    pyref = os.path.join(REF, "pyref")
    kind = "port"
    if os.path.isdir(os.path.join(pyref, "pgsd")):
        sys.dont_write_bytecode = True
        sys.path.insert(0, pyref)
        try:
            import pgsd.hoomd as ref_hoomd
            import pgsd.pypgsd as ref_pypgsd
            kind = "reference"
        finally:
            sys.path.remove(pyref)
    ts = []
    for i in range(steps + warm):
        t0 = time.perf_counter()
        if kind == "reference":
            with open(path, "rb") as fh:
                traj = ref_hoomd.HOOMDTrajectory(ref_pypgsd.PGSDFile(fh))
                fr = traj[1]
                ids = fr.log['particles/id']
                o = np.argsort(ids, kind='stable')
                p = fr.particles
                res = [ids[o]] + [np.asarray(getattr(p, k))[o] for k in ("position", "velocity", "typeid", "density", "pressure")]
        else:
            from oracle import reader_oracle, reorder_oracle
            orc = reader_oracle.OracleFile(path)
            r = reorder_oracle.reorder_frame(reader_oracle.decode_particles(orc, 1))
            orc.close()
            res = [r['log/particles/id']]
        ts.append(time.perf_counter() - t0)
        assert res[0][0] == 0 and res[0][-1] == n - 1
    os.unlink(path)
    t = float(np.mean(ts[warm:]))
    what = ("the reference's own pypgsd.py + hoomd.py (unmodified, staged by oracle/build_ref.sh) on a file written by the "
            "unmodified pgsd.c" if kind == "reference" else "oracle port of pypgsd + hoomd decode")
    return {"s_per_frame": t, "Mpps": n / t / 1e6, "kind": kind, "what": what}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=WRITE_PARTICLES)
    ap.add_argument("--read-particles", type=int, default=READ_PARTICLES)
    ap.add_argument("--small-frames", type=int, default=SMALL_FRAMES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vtu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the 14 GB benchmark-write/read legs, the disk target and the "
                                                         "config 2 / 5 legs (contract tests, development)")
    ap.add_argument("--legs", default="", help="development: comma list of legs to run (read,dist,write,parity,disk,traj,small,bw); "
                                                  "a partial run prints only the objects it measured")
    ap.add_argument("--only-distributed", action="store_true",
                    help="development: run only the distributed-reorder leg and print its object")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # A run of this file takes 2-4 minutes.  Should a leg ever stall (a rank waiting for another one), do not sit on
    # the box until somebody else's limit strikes: after PGSD_BENCH_WATCHDOG_S seconds (default 1500) every thread's
    # Python stack goes to stderr and the process exits with status 1 (torchrun then ends the other ranks).
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ.get("PGSD_BENCH_WATCHDOG_S", "1500")), exit=True)
    ncores = os.cpu_count() or 1
    bdir = bench_dir()
    common_cfg = {"workload": f"config 3: one {args.particles}-particle HOOMD-schema frame per step "
                              f"(position/velocity/typeid/density/pressure/id, 40 B/particle, "
                              f"{BPP * args.particles / 1e9:.2f} GB), row-partitioned over the ranks",
                  "file_target": f"{bdir} ({fs_kind(bdir)}; no fsync, as the reference)",
                  "l2": "inputs larger than L2 (no flush needed)"}

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return 0
        nr = min(ncores, 16)
        n_w = args.particles  # the full frame: a few seconds per step on the host cores
        st_fs = os.statvfs(bdir)
        budget = min(64 << 30, int(0.35 * st_fs.f_bavail * st_fs.f_frsize)) - BPP * n_w  # minus the input blob
        warm = max(1, min(args.warmup, 2))
        steps_ref = max(1, min(args.steps, budget // (BPP * n_w) - warm))  # the file must fit the target
        w = cpu_write_reference(n_w, steps_ref, nr, warm=warm)
        if w is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver is not built"}))
            return 0
        n_r = min(args.read_particles, 2 * 1024 * 1024)
        r = cpu_read_reference(n_r, min(args.steps, 4), warm=1)
        bw, br = (None, None) if args.quick else cpu_benchmark_reference(min(ncores, 8))
        line = {
            "impl": "reference", "metric": "frame_write_GBps", "value": w["GBps"], "unit": "GB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": w["s_per_frame"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": common_cfg,
            "cpu_baseline": {"value": w["GBps"], "unit": "GB/s", "cores": nr, "kind": "reference",
                             "sample": f"{w['frames']} frames of {n_w} particles, unmodified reference pgsd.c + MPI shim, "
                                       f"{nr} ranks (processes)"},
            "e2e": {"value": w["GBps"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "benchmark_write": bw, "benchmark_read": br,
            "read_reorder": {"metric": "id_reordered_read_Mparticles_per_s", "value": r["Mpps"], "unit": "Mparticles/s",
                             "e2e": {"value": r["Mpps"], "unit": "Mparticles/s", "h2d_bytes_per_step": 0,
                                     "d2h_bytes_per_step": 0},
                             "cpu_baseline": {"value": r["Mpps"], "unit": "Mparticles/s", "cores": 1, "kind": r["kind"],
                                              "sample": f"{n_r}-particle frames: {r['what']}, numpy stable argsort + gather "
                                                        "(the reference reader is single-process)"}},
        }
        if not args.quick:
            dd = disk_dir()
            if dd:
                # the same volume as the repo's disk leg (3 warm-up + up to 4 timed full-size frames in one file): on a
                # disk-backed target the rate depends on how far into write-back the run gets
                dsteps = max(2, min(args.steps, 4))
                wd = cpu_write_reference(n_w, dsteps, nr, warm=3, target_dir=dd)
                if wd:
                    line["write_disk"] = {"metric": "frame_write_GBps", "value": wd["GBps"], "unit": "GB/s",
                                          "file_target": f"{bench_dir(dd)} ({fs_kind(dd)}; no fsync)", "ranks": nr,
                                          "sample": f"{wd['frames']} timed frames of {n_w} particles after 3 warm-up frames, one file"}
            t2 = cpu_frames_reference(TRAJ_PARTICLES, TRAJ_FRAMES, 2, 0)
            if t2:
                line["trajectory_write"] = {"metric": "trajectory_write_GBps", "value": t2["GBps"], "unit": "GB/s", "ranks": 2,
                                            "workload": "config 2: 100 frames x 1 Mi particles, 2 ranks"}
            t5 = cpu_frames_reference(SMALL_PARTICLES, min(args.small_frames, SMALL_FRAMES), min(8, ncores), SMALL_LOGS)
            if t5:
                line["small_frames"] = {"metric": "small_frame_us", "value": t5["us_per_frame"], "unit": "us/frame",
                                        "higher_is_better": False, "ranks": t5["ranks"],
                                        "workload": "config 5: 4096 particles + 8 log scalars per frame (18 chunks)"}
        print(json.dumps(line))
        return 0

    # ---- this repo's CUDA path
    from pgsd_sph_b200 import _lib
    lib = _lib.load()
    if not lib.pgsd_b200_cuda_available():
        raise SystemExit("bench.py: no CUDA device; pgsd_sph_b200 has no CPU fallback")
    dist = Dist(args.gpus)
    _lib.check(lib.pgsd_b200_device_init(dist.local), "device_init")
    peaks = load_peaks()
    sampler = ClockSampler(dist.local)
    if dist.rank == 0:
        sampler.start()
    windows = []
    if args.only_distributed:
        if dist.world > 1:
            from pgsd_sph_b200 import comm
            comm.init_nccl(dist.rank, dist.world, dist.bcast_bytes, dist.local)
        dr = run_dist_reorder_leg(lib, dist, args, peaks, windows)
        sampler.stop()
        if dist.rank == 0:
            dr["n_gpus"] = dist.world
            dr["clocks"] = sampler.summary(windows)
            print(json.dumps(dr), flush=True)
        if dist.world > 1:
            lib.pgsd_b200_comm_finalize()
        dist.close()
        lib.pgsd_b200_shutdown()
        return 0
    legs = set(x for x in args.legs.split(",") if x)
    if legs:   # development: a subset of the legs, printed as they are
        if dist.world > 1:
            from pgsd_sph_b200 import comm
            comm.init_nccl(dist.rank, dist.world, dist.bcast_bytes, dist.local)
        out = {"n_gpus": dist.world, "partial": sorted(legs)}
        if "read" in legs and dist.world == 1:
            out["read_reorder"] = run_read_leg(lib, dist, args, peaks, windows)
        if "dist" in legs:
            out["distributed_reorder"] = run_dist_reorder_leg(lib, dist, args, peaks, windows)
        if "write" in legs:
            out["write"] = run_write_leg(lib, dist, args, peaks, windows)
        if "parity" in legs:
            out["parity"] = run_parity_leg(lib, dist, args)
        if "disk" in legs and disk_dir():
            out["write_disk"] = run_write_leg(lib, dist, args, peaks, windows, target_dir=disk_dir(), steps=max(2, min(args.steps, 4)),
                                              warmup=3, legs=("device",))
        if "traj" in legs:
            out["trajectory_write"] = run_frames_leg(lib, dist, args, windows, TRAJ_PARTICLES, TRAJ_FRAMES, 0, "traj")
        if "small" in legs:
            out["small_frames"] = run_frames_leg(lib, dist, args, windows, SMALL_PARTICLES, args.small_frames, SMALL_LOGS, "small")
        if "bw" in legs:
            bw, bw_path = run_benchmark_write_leg(lib, dist, args, keep=True)
            out["benchmark_write"], out["benchmark_read"] = bw, run_benchmark_read_leg(lib, dist, args, bw_path)
        sampler.stop()
        if dist.rank == 0:
            out["clocks"] = sampler.summary(windows)
            print(json.dumps(out, default=str), flush=True)
        if dist.world > 1:
            lib.pgsd_b200_comm_finalize()
        dist.close()
        lib.pgsd_b200_shutdown()
        return 0
    rd = run_read_leg(lib, dist, args, peaks, windows)   # communicator: "single" (frames are independent)
    if dist.world > 1:
        from pgsd_sph_b200 import comm
        comm.init_nccl(dist.rank, dist.world, dist.bcast_bytes, dist.local)
    dr = run_dist_reorder_leg(lib, dist, args, peaks, windows)
    wr = run_write_leg(lib, dist, args, peaks, windows)
    parity = run_parity_leg(lib, dist, args)
    extra = {}
    if not args.quick:
        tr = run_frames_leg(lib, dist, args, windows, TRAJ_PARTICLES, TRAJ_FRAMES, 0, "traj")
        extra["trajectory_write"] = {
            "metric": "trajectory_write_GBps", "value": tr["device"]["GBps"], "unit": "GB/s",
            "e2e": {"value": tr["host"]["GBps"], "unit": "GB/s", "path": "host SoA columns -> file"},
            "workload": f"config 2: {TRAJ_FRAMES} frames x {TRAJ_PARTICLES} particles (41.9 MB/frame), rows over {dist.world} rank(s)",
            "file_bytes": tr["device"]["file_bytes"], "gpu_launches": tr["device"]["launches"]}
        # a 0.6-s leg of 60-us frames is the one most easily disturbed by whatever else the host is doing: median of 3
        sms = [run_frames_leg(lib, dist, args, windows, SMALL_PARTICLES, args.small_frames, SMALL_LOGS, "small") for _ in range(3)]
        sm = {m: sorted(sms, key=lambda r: r[m]["us_per_frame"])[1][m] for m in ("device", "host")}
        extra["small_frames"] = {
            "metric": "small_frame_us", "value": sm["device"]["us_per_frame"], "unit": "us/frame", "higher_is_better": False,
            "e2e": {"value": sm["host"]["us_per_frame"], "unit": "us/frame", "path": "host SoA columns -> file"},
            "workload": f"config 5: {args.small_frames} frames x {SMALL_PARTICLES} particles + {SMALL_LOGS} log scalars "
                        f"(18 chunks/frame), rows over {dist.world} rank(s); index/namelist append + offset scan latency",
            "collectives_per_frame": sm["device"]["collectives_per_frame"], "file_bytes": sm["device"]["file_bytes"],
            "gpu_launches": sm["device"]["launches"],
            "runs_us_per_frame": {m: [round(r[m]["us_per_frame"], 2) for r in sms] for m in ("device", "host")},
            "runs_note": "value / e2e are the median of these 3 runs of the whole leg"}
        bw, bw_path = run_benchmark_write_leg(lib, dist, args, keep=True)
        extra["benchmark_write"] = bw
        extra["benchmark_read"] = run_benchmark_read_leg(lib, dist, args, bw_path)
        # the disk-backed target last: it leaves ~15 GB of dirty pages whose write-back would disturb whatever followed
        dd = disk_dir()
        if dd:
            wd = run_write_leg(lib, dist, args, peaks, windows, target_dir=dd, steps=max(2, min(args.steps, 4)), warmup=3,
                               legs=("device",))
            extra["write_disk"] = {k: wd[k] for k in ("metric", "value", "unit", "ms_per_step", "file_target")}
            extra["write_disk"]["file_ceiling_GBps"] = wd["split"]["file_ceiling_GBps"]
            extra["write_disk"]["file_ceiling"] = wd["split"]["file_ceiling"]
    sampler.stop()

    line = None
    if dist.rank == 0:
        line = {
            "metric": wr["metric"], "value": wr["value"], "unit": wr["unit"], "n_gpus": dist.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": wr["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(common_cfg, comm=lib.pgsd_b200_comm_kind().decode(), parallelism=f"rows/{dist.world}"),
            "e2e": wr["e2e"], "roofline": wr["roofline"], "gpu_launches": wr["gpu_launches"] + rd["gpu_launches"] + dr["gpu_launches"],
            "split": wr["split"], "parity": parity, "read_reorder": rd, "distributed_reorder": dr,
            "clocks": sampler.summary(windows),
            "vs_baseline_note": "BASELINE.md's only published number (0.175 GB/s benchmark-write, f64 keys, NVMe) is "
                                "for another workload and storage; not used",
        }
        line.update(extra)
        if dist.world == 1 and not args.no_cpu_baseline:
            nr = min(ncores, 16)
            n_w = min(args.particles, 16 * 1024 * 1024)
            w = cpu_write_reference(n_w, 3, nr, warm=1)
            n_r = min(args.read_particles, 2 * 1024 * 1024)
            r = cpu_read_reference(n_r, 2, warm=1)
            if w:
                line["cpu_baseline"] = {"value": w["GBps"], "unit": "GB/s", "cores": nr, "kind": "reference",
                                        "sample": f"3 frames of {n_w} particles, unmodified reference pgsd.c + MPI "
                                                  f"shim at {nr} ranks, same file target"}
            if r:
                line["read_reorder"]["cpu_baseline"] = {
                    "value": r["Mpps"], "unit": "Mparticles/s", "cores": 1, "kind": r["kind"],
                    "sample": f"2 frames of {n_r} particles: {r['what']}, numpy stable argsort + gather"}
            if not args.quick:
                if "write_disk" in line:
                    # same frame size, warm-up and frame count as the leg above (the rate on a disk-backed target
                    # depends on how far into write-back a run gets)
                    dsteps = max(2, min(args.steps, 4))
                    wd = cpu_write_reference(args.particles, dsteps, nr, warm=3, target_dir=disk_dir())
                    if wd:
                        line["write_disk"]["cpu_baseline"] = {
                            "value": wd["GBps"], "unit": "GB/s", "cores": nr, "kind": "reference",
                            "sample": f"{wd['frames']} timed frames of {args.particles} particles after 3 warm-up frames at {nr} "
                                      f"ranks, same target, one file"}
                t2 = cpu_frames_reference(TRAJ_PARTICLES, 30, 2, 0)
                if t2:
                    line["trajectory_write"]["cpu_baseline"] = {
                        "value": t2["GBps"], "unit": "GB/s", "cores": 2, "kind": "reference",
                        "sample": "30 frames x 1 Mi particles, unmodified pgsd.c at 2 ranks (the configuration's rank count)"}
                for P in (1, min(8, ncores)):
                    t5 = cpu_frames_reference(SMALL_PARTICLES, min(args.small_frames, 5000), P, SMALL_LOGS)
                    if t5:
                        line["small_frames"]["cpu_baseline" if P == 1 else "cpu_baseline_8_ranks"] = {
                            "value": t5["us_per_frame"], "unit": "us/frame", "cores": P, "kind": "reference",
                            "sample": f"{t5['frames']} frames, same 18 chunks per frame, unmodified pgsd.c at {P} rank(s)"}
        if parity and parity.get("sha256_equal") is False:
            line["parity_failed"] = True
    if dist.world > 1:
        lib.pgsd_b200_comm_finalize()
    dist.close()
    if line:
        print(json.dumps(line), flush=True)
    lib.pgsd_b200_shutdown()
    if line and line.get("parity_failed"):
        log("bench.py: PARITY FAILURE -- the file written at P = %d differs from the reference's" % dist.world)
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
