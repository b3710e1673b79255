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
K1 pack, K2 offset scan, K3 pinned staging + pwrite.  A step of the read leg is one 16 Mi-particle
frame (config 4) put into particle-ID order: K4 radix sort + K5 gather; frames are independent,
every rank processes its own (weak).  `value` legs start with inputs resident in HBM; `e2e` legs
start from host buffers / the file and end in the file / host arrays.

Only this file's `cpu_baseline` leg and `--impl reference` execute anything under oracle/.
"""
import argparse
import ctypes as C
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
BPP = 40                             # bytes per particle


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


def make_soa(n_total, start, n, seed):
    """This rank's rows [start, start+n) of a synthetic n_total-particle frame as 10 SoA columns
    (pos xyz, vel xyz, density, pressure: f32; typeid, id: u32).  ids = slice of one seeded
    permutation of 0..n_total-1 (dense, unique, unsorted)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ids = rng.permutation(n_total).astype(np.uint32)[start:start + n].copy()
    rng = np.random.Generator(np.random.PCG64(seed * 1000003 + start))
    cols = [(rng.random(n, dtype=np.float32) * np.float32(10.0)) for _ in range(3)]
    cols += [(rng.random(n, dtype=np.float32) - np.float32(0.5)) for _ in range(3)]
    dens = np.float32(1000.0) * (np.float32(1.0) + np.float32(0.01) * (rng.random(n, dtype=np.float32) - np.float32(0.5)))
    cols.append(dens.astype(np.float32))
    cols.append((np.float32(2.25) * (dens - np.float32(1000.0))).astype(np.float32))
    cols.append(rng.integers(0, 3, size=n, dtype=np.uint32))
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


def bench_dir():
    d = os.environ.get("PGSD_BENCH_DIR")
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


def run_read_leg(lib, dist, args, peaks, windows):
    """ID-reordered read, frame-parallel: this rank's own frames of a 16 Mi-particle trajectory."""
    from pgsd_sph_b200 import _lib, fl, hoomd, synth
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
    traj.close()
    os.unlink(path)

    peak = peaks["hbm_gbs"]
    def kern(ms, bytes_pp, note):
        t = max(ms, 1e-6) * 1e-3
        return {"ms": ms, "algorithmic_bytes": bytes_pp * n, "GBps": bytes_pp * n / t / 1e9,
                "frac": bytes_pp * n / t / 1e9 / peak, "note": note}

    slot = os.environ.get("PGSD_B200_SLOT", "1") != "0" and phase_ms[2] < 0.02   # unique ids: no pair passes ran
    if slot:
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
    return {
        "metric": "id_reordered_read_Mparticles_per_s", "value": value, "unit": "Mparticles/s",
        "ms_per_step": dev_ms / args.steps, "scaling": "weak", "dtype": "u32",
        "config": {"workload": f"config 4: {n}-particle unsorted frames, 40 B/particle, one frame per step per GPU "
                               "(frame-parallel, no collective)", "l2": "inputs (671 MB) larger than L2"},
        "e2e": {"value": e2e, "unit": "Mparticles/s", "h2d_bytes_per_step": st.h2d_bytes // args.steps,
                "d2h_bytes_per_step": st.d2h_bytes // args.steps,
                "path": "file -> pgsd.hoomd.HOOMDTrajectory(reorder='id')[i] -> host numpy arrays"},
        "roofline": {"bound": "hbm", "achieved": 80 * n / reorder_s / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": 80 * n / reorder_s / 1e9 / peak, "traffic": None, "peak_source": peaks["source"],
                     "what": "K4+K5 reorder as one operation: 80 B/particle algorithmic"},
        "kernels": kernels, "reorder_path": "slot (unique ids)" if slot else "general (stable LSD)",
        "gpu_launches": int(launches),
    }


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
    return {
        "metric": "distributed_id_reorder_Mparticles_per_s", "value": value, "unit": "Mparticles/s",
        "ms_per_step": dev_ms / args.steps, "wall_ms_per_step": 1e3 * wall / args.steps, "scaling": "strong",
        "config": {"workload": f"config 3 read-back: one {n_total}-particle frame, rows partitioned over {dist.world} rank(s), "
                               "40 B/particle, ids dense and unsorted; fields resident in HBM on both sides",
                   "comm": lib.pgsd_b200_comm_kind().decode(),
                   "timed": "whole collective call (5 host all-gathers of sizes / IPC handles / bucket counts / flags, "
                            "histogram, scatter over NVLink, placement), CUDA events on the stream, max over ranks"},
        "rows_owned_rank0": k, "gpu_launches": int(dist.sum(float(st.kernel_launches))),
        "collectives_per_step": st.collectives // max(args.steps, 1),
    }


def run_write_leg(lib, dist, args, peaks, windows):
    from pgsd_sph_b200 import _lib, fl, synth
    from pgsd_sph_b200.devmem import DeviceArray, PinnedArray
    n_total = args.particles
    rows, start = rank_rows(n_total, dist.world, dist.rank)
    n = rows[dist.rank]
    t0 = time.perf_counter()
    cols = make_soa(n_total, start, n, 20261018)
    log(f"[rank {dist.rank}] write leg: generated {n} of {n_total} particles in {time.perf_counter() - t0:.1f}s")
    d_cols = [DeviceArray.from_numpy(a) for a in cols]
    pinned = [PinnedArray(a.shape, a.dtype) for a in cols]
    for p, a in zip(pinned, cols):
        p.array[...] = a
    h_cols = [p.array for p in pinned]
    payload = BPP * n_total
    path = os.path.join(bench_dir(), "write_A.gsd")
    result = {}
    for leg, src in (("device", d_cols), ("e2e", h_cols)):
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
        st_fs = os.statvfs(bench_dir())
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

        for i in range(args.warmup):
            roll()
            step(i)
            in_file += 1
        f.flush()
        lib.pgsd_b200_reset_stats()
        dist.barrier()
        lib.pgsd_b200_synchronize()
        w0 = time.perf_counter()
        for i in range(args.steps):
            roll()
            step(args.warmup + i, timed=True)
            in_file += 1
        f.flush()  # drains every queued D2H + pwrite of this rank
        lib.pgsd_b200_synchronize()
        dt = time.perf_counter() - w0
        dist.barrier()
        windows.append((w0, time.perf_counter()))
        st = get_stats(lib)
        dt = dist.max(dt)
        lib.pgsd_b200_pack_profiling(0)
        f.close()
        size = os.path.getsize(path) if dist.rank == 0 else 0
        result[leg] = {"s": dt, "GBps": payload * args.steps / dt / 1e9, "stats": st, "k1_ms": k1_ms,
                       "file_bytes": size}
        dist.barrier()
        if dist.rank == 0:
            os.unlink(path)
    for a in d_cols:
        a.free()

    # ---- K1 alone, device to device, CUDA events (the kernel's roofline point)
    dev, e2e = result["device"], result["e2e"]
    k1_s = float(np.median(dev["k1_ms"])) * 1e-3
    k1_s = dist.max(k1_s)
    peak = peaks["hbm_gbs"]
    k1_bytes = 80 * n  # read 40 + write 40 B/particle, this rank's launch
    out = {
        "metric": "frame_write_GBps", "value": dev["GBps"], "unit": "GB/s", "ms_per_step": dev["s"] / args.steps * 1e3,
        "e2e": {"value": e2e["GBps"], "unit": "GB/s",
                "h2d_bytes_per_step": int(dist.sum(float(e2e["stats"].h2d_bytes))) // args.steps,
                "d2h_bytes_per_step": int(dist.sum(float(e2e["stats"].d2h_bytes))) // args.steps,
                "path": "pinned host SoA columns -> pgsd_b200_write_chunks_soa (H2D, K1, D2H) -> pwrite -> file"},
        "roofline": {"bound": "hbm", "achieved": k1_bytes / k1_s / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": k1_bytes / k1_s / 1e9 / peak, "traffic": k1_traffic(n), "peak_source": peaks["source"],
                     "kernel": "k1_pack_frame", "ms": k1_s * 1e3, "algorithmic_bytes": k1_bytes},
        "gpu_launches": int(dist.sum(float(dev["stats"].kernel_launches))),
        "split": {"device_k1_ms_per_frame": k1_s * 1e3,
                  "commit_wait_s_per_frame": dev["stats"].commit_wait_s / args.steps,
                  "d2h_bytes_per_step": int(dist.sum(float(dev["stats"].d2h_bytes))) // args.steps,
                  "file_bytes": dev["file_bytes"],
                  "note": "wall = max(K1, D2H over PCIe, pwrite into the page cache); K1 is <1% of it"},
    }
    return out


def k1_traffic(n):
    """DRAM bytes of one K1 launch from the committed ncu capture (profiles/k1_traffic.json), scaled
    to this launch's particle count; None if the capture is missing."""
    p = os.path.join(REPO, "profiles", "k1_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return int(d["traffic_bytes"] * (n / d["particles"]))


def run_benchmark_write_leg(lib, dist, args):
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
    names = ["key%d" % k for k in range(nkeys)]
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
    if dist.rank == 0:
        os.unlink(path)
    for a in keys:
        a.free()
    mib = (nframes - nframes // 2) * nkeys * n_total * 8 / 1048576.0
    return {"metric": "benchmark_write_MiBps", "value": mib / t2, "unit": "MiB/s", "seconds": t2,
            "workload": "17 keys x 100 frames x 1 Mi float64 (14.26 GB), second 50 frames timed, as benchmark-write.cc",
            "vs_baseline": mib / t2 / 167.0,
            "vs_baseline_note": "published 167.0 MiB/s at 1 rank on NVMe (CHANGELOG.md:186); this run writes to "
                                "tmpfs, so the ratio mixes implementation and storage"}


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def cpu_write_reference(n, frames, nranks, warm=1):
    """The UNMODIFIED reference pgsd.c (oracle/_ref/ref_driver, MPI shim) writing the same frame."""
    drv = os.path.join(REPO, "oracle", "_ref", "ref_driver")
    if not os.path.exists(drv):
        return None
    d = bench_dir()
    blob, out = os.path.join(d, "ref_blob.bin"), os.path.join(d, "ref_write.gsd")
    cols = make_soa(n, 0, n, 20261018)
    order = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]
    with open(blob, "wb") as fh:
        for j in order:
            fh.write(cols[j].tobytes())
    del cols
    env = dict(os.environ, PGSD_SHIM_NP=str(nranks))
    r = subprocess.run([drv, "bench", out, str(n), str(frames + warm), blob], env=env, capture_output=True, text=True)
    for p in (blob, out):
        if os.path.exists(p):
            os.unlink(p)
    if r.returncode != 0:
        log("ref_driver failed:", r.stderr[-400:])
        return None
    t = json.loads(r.stdout.strip().splitlines()[-1])["frame_s"][warm:]
    return {"s_per_frame": float(np.mean(t)), "GBps": BPP * n / float(np.mean(t)) / 1e9, "frames": len(t)}


def cpu_benchmark_write_reference(nranks):
    """The reference's OWN benchmark binary (scripts/benchmark-write.cc compiled unmodified against the
    MPI shim): 17 keys x 100 frames x 1 Mi float64, prints "MB/s:" (MiB/s of the second 50 frames)."""
    exe = os.path.join(REPO, "oracle", "_ref", "benchmark-write")
    if not os.path.exists(exe):
        return None
    d = bench_dir()
    r = subprocess.run([exe], cwd=d, env=dict(os.environ, PGSD_SHIM_NP=str(nranks)), capture_output=True, text=True)
    out = os.path.join(d, "test%d.gsd" % nranks)
    if os.path.exists(out):
        os.unlink(out)
    for line in r.stdout.splitlines():
        if line.startswith("MB/s:"):
            return {"metric": "benchmark_write_MiBps", "value": float(line.split()[1]), "unit": "MiB/s", "ranks": nranks,
                    "workload": "17 keys x 100 frames x 1 Mi float64, unmodified benchmark-write.cc + pgsd.c + MPI shim",
                    "vs_baseline": float(line.split()[1]) / 167.0}
    log("benchmark-write failed:", r.stderr[-300:])
    return None


def cpu_read_reference(n, steps, warm=1):
    """Oracle port of the reference reader (pypgsd + hoomd decode) + numpy stable argsort + gather."""
    from oracle import reader_oracle, reorder_oracle
    from pgsd_sph_b200 import fl, synth
    path = os.path.join(bench_dir(), "ref_read.gsd")
    with fl.open(path, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
        for i in range(2):
            cols = make_soa(n, 0, n, 7000 + i)
            for k, a in synth.frame_scalars(n, i):
                f.write_chunk(k, a, write_all=False)
            for nm, idx, dt in SOA_CHUNKS:
                a = np.stack([cols[j] for j in idx], axis=1) if len(idx) > 1 else cols[idx[0]]
                f.write_chunk(nm, np.ascontiguousarray(a))
            f.end_frame()
    ts = []
    for i in range(steps + warm):
        t0 = time.perf_counter()
        orc = reader_oracle.OracleFile(path)
        res = reorder_oracle.reorder_frame(reader_oracle.decode_particles(orc, 1))
        orc.close()
        ts.append(time.perf_counter() - t0)
        assert res['log/particles/id'][0] == 0
    os.unlink(path)
    t = float(np.mean(ts[warm:]))
    return {"s_per_frame": t, "Mpps": n / t / 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=WRITE_PARTICLES)
    ap.add_argument("--read-particles", type=int, default=READ_PARTICLES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the 14 GB benchmark-write leg (contract tests)")
    ap.add_argument("--only-distributed", action="store_true",
                    help="development: run only the distributed-reorder leg and print its object")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
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
        n_r = min(args.read_particles, 2 * 1024 * 1024)
        r = cpu_read_reference(n_r, min(args.steps, 4), warm=1)
        bw = None if args.quick else cpu_benchmark_write_reference(min(ncores, 8))
        if w is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver is not built"}))
            return 0
        line = {
            "impl": "reference", "metric": "frame_write_GBps", "value": w["GBps"], "unit": "GB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": w["s_per_frame"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": common_cfg,
            "cpu_baseline": {"value": w["GBps"], "unit": "GB/s", "cores": nr, "kind": "reference",
                             "sample": f"{w['frames']} frames of {n_w} particles, unmodified reference pgsd.c + MPI shim, "
                                       f"{nr} ranks (processes)"},
            "e2e": {"value": w["GBps"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "benchmark_write": bw,
            "read_reorder": {"metric": "id_reordered_read_Mparticles_per_s", "value": r["Mpps"], "unit": "Mparticles/s",
                             "e2e": {"value": r["Mpps"], "unit": "Mparticles/s", "h2d_bytes_per_step": 0,
                                     "d2h_bytes_per_step": 0},
                             "cpu_baseline": {"value": r["Mpps"], "unit": "Mparticles/s", "cores": 1, "kind": "port",
                                              "sample": f"{n_r}-particle frames: oracle port of pypgsd + hoomd decode, "
                                                        "numpy stable argsort + gather (the reference reader is "
                                                        "single-process)"}},
        }
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
    rd = run_read_leg(lib, dist, args, peaks, windows)   # communicator: "single" (frames are independent)
    if dist.world > 1:
        from pgsd_sph_b200 import comm
        comm.init_nccl(dist.rank, dist.world, dist.bcast_bytes, dist.local)
    dr = run_dist_reorder_leg(lib, dist, args, peaks, windows)
    wr = run_write_leg(lib, dist, args, peaks, windows)
    bw = None if args.quick else run_benchmark_write_leg(lib, dist, args)
    sampler.stop()

    line = None
    if dist.rank == 0:
        line = {
            "metric": wr["metric"], "value": wr["value"], "unit": wr["unit"], "n_gpus": dist.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": wr["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(common_cfg, comm=lib.pgsd_b200_comm_kind().decode(), parallelism=f"rows/{dist.world}"),
            "e2e": wr["e2e"], "roofline": wr["roofline"], "gpu_launches": wr["gpu_launches"] + rd["gpu_launches"] + dr["gpu_launches"],
            "split": wr["split"], "read_reorder": rd, "distributed_reorder": dr, "benchmark_write": bw, "clocks": sampler.summary(windows),
            "vs_baseline_note": "BASELINE.md's only published number (0.175 GB/s benchmark-write, f64 keys, NVMe) is "
                                "for another workload and storage; not used",
        }
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
            line["read_reorder"]["cpu_baseline"] = {
                "value": r["Mpps"], "unit": "Mparticles/s", "cores": 1, "kind": "port",
                "sample": f"2 frames of {n_r} particles: oracle reader + numpy stable argsort + gather"}
    if dist.world > 1:
        lib.pgsd_b200_comm_finalize()
    dist.close()
    if line:
        print(json.dumps(line), flush=True)
    lib.pgsd_b200_shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
