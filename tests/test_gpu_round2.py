"""GPU tests added in round 2: frame arenas owned per file handle, frame-0 fallback fields under reorder='id',
validation of device arrays handed to the reorder, a slow stress of the device write path in both file modes."""
import hashlib
import json
import os

import numpy as np
import pytest

import opscript
from golden.make_golden import seed_nprocs
from pgsd_sph_b200 import fl, hoomd, synth
from pgsd_sph_b200.devmem import DeviceArray
from randscript import random_script

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def cuda(lib):
    assert lib.pgsd_b200_cuda_available() == 1, "no CUDA device: the device path has no CPU fallback"


def _frame_chunks(n, i, seed):
    rng = np.random.default_rng(seed * 100 + i)
    return [("particles/position", rng.standard_normal((n, 3)).astype(np.float32)),
            ("particles/density", rng.standard_normal(n).astype(np.float32)),
            ("log/particles/id", rng.permutation(n).astype(np.uint32))]


def _write(path, n, frames, seed, device):
    with fl.open(path, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
        for i in range(frames):
            for k, a in synth.frame_scalars(n, i):
                f.write_chunk(k, a, write_all=False)
            for k, a in _frame_chunks(n, i, seed):
                f.write_chunk(k, DeviceArray.from_numpy(a) if device else a)
            f.end_frame()


def test_two_files_interleave_device_frames(tmp_path):
    """Two writable handles pack device chunks alternately and commit in the other order: every handle owns its
    frame arena (one process-global arena let B's frame_submit take A's chunks)."""
    n, frames = 50000, 4
    ra, rb = str(tmp_path / "ra.gsd"), str(tmp_path / "rb.gsd")
    _write(ra, n, frames, 1, device=False)
    _write(rb, n + 17, frames, 2, device=False)
    pa, pb = str(tmp_path / "a.gsd"), str(tmp_path / "b.gsd")
    fa = fl.open(pa, 'w', 'pgsd-b200', 'hoomd', [1, 4])
    fb = fl.open(pb, 'w', 'pgsd-b200', 'hoomd', [1, 4])
    for i in range(frames):
        ca, cb = _frame_chunks(n, i, 1), _frame_chunks(n + 17, i, 2)
        for k, a in synth.frame_scalars(n, i):
            fa.write_chunk(k, a, write_all=False)
        for k, a in synth.frame_scalars(n + 17, i):
            fb.write_chunk(k, a, write_all=False)
        for (ka, a), (kb, b) in zip(ca, cb):          # A.write, B.write, A.write, B.write ...
            fa.write_chunk(ka, DeviceArray.from_numpy(a))
            fb.write_chunk(kb, DeviceArray.from_numpy(b))
        if i % 2 == 0:                                 # ... and the commits in alternating order
            fb.end_frame()
            fa.end_frame()
        else:
            fa.end_frame()
            fb.end_frame()
    fa.close()
    fb.close()
    assert opscript.read_bytes(pa) == opscript.read_bytes(ra)
    assert opscript.read_bytes(pb) == opscript.read_bytes(rb)


def test_handle_closed_with_an_uncommitted_device_frame_releases_its_arena(tmp_path):
    n = 20000
    for rep in range(5):   # more often than there are arenas: a leaked arena would block the next pack
        f = fl.open(str(tmp_path / f"x{rep}.gsd"), 'w', 'pgsd-b200', 'hoomd', [1, 4])
        f.write_chunk("particles/density", DeviceArray.from_numpy(np.ones(n, dtype=np.float32)))
        f.close()          # close commits what was written (ref: pgsd_close flushes, pgsd.c:1814-1914)
    p = str(tmp_path / "y.gsd")
    _write(p, n, 2, 3, device=True)
    q = str(tmp_path / "yr.gsd")
    _write(q, n, 2, 3, device=False)
    assert opscript.read_bytes(p) == opscript.read_bytes(q)


def test_reorder_uses_frame0_order_for_fields_that_fall_back_to_frame0(tmp_path):
    """particles/mass is only written in frame 0 and the storage order changes in frame 1 (particles migrated between
    ranks): with reorder='id' the fallback values must follow frame 0's ids, not frame 1's permutation.
    Oracle: the reference reader's fallback rule (hoomd.py:866-870) + stable argsort by each frame's own ids."""
    n = 30011
    rng = np.random.default_rng(5)
    ids0, ids1 = rng.permutation(n).astype(np.uint32), rng.permutation(n).astype(np.uint32)
    mass0 = rng.standard_normal(n).astype(np.float32)
    tag0 = rng.integers(0, 1000, size=n).astype(np.int32)
    pos = [rng.standard_normal((n, 3)).astype(np.float32) for _ in range(2)]
    p = str(tmp_path / "fb.gsd")
    with fl.open(p, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
        for i, ids in enumerate((ids0, ids1)):
            for k, a in synth.frame_scalars(n, i):
                f.write_chunk(k, a, write_all=False)
            f.write_chunk("particles/position", pos[i])
            f.write_chunk("log/particles/id", ids)
            if i == 0:
                f.write_chunk("particles/mass", mass0)
                f.write_chunk("log/particles/tag", tag0)
            f.end_frame()
    o0, o1 = np.argsort(ids0, kind='stable'), np.argsort(ids1, kind='stable')
    for device in (False, True):
        with hoomd.open(p, 'r', reorder='id', device=device) as t:
            f1 = t[1]
            get = (lambda a: a.to_numpy()) if device else (lambda a: np.asarray(a))
            assert get(f1.particles.position).tobytes() == pos[1][o1].tobytes()
            assert get(f1.particles.mass).tobytes() == mass0[o0].tobytes()          # mass of particle id k at row k
            assert get(f1.log['particles/tag']).tobytes() == tag0[o0].tobytes()
            f0 = t[0]
            assert get(f0.particles.mass).tobytes() == mass0[o0].tobytes()


def test_reorder_rejects_short_and_strided_device_fields():
    n = 5000
    ids = DeviceArray.from_numpy(np.random.default_rng(1).permutation(n).astype(np.uint32))
    short = DeviceArray.from_numpy(np.zeros(n - 1, dtype=np.float32))
    with pytest.raises(ValueError, match="rows"):
        hoomd.reorder_by_id(ids, {"short": short}, device=True)
    buf = DeviceArray.from_numpy(np.zeros((n, 4), dtype=np.float32))

    class View:  # (n, 3) view of the (n, 4) buffer: strided rows
        __cuda_array_interface__ = {"shape": (n, 3), "typestr": "<f4", "data": (buf.ptr, False), "version": 3,
                                    "strides": (16, 4)}

    with pytest.raises(ValueError, match="contiguous"):
        hoomd.reorder_by_id(ids, {"view": View()}, device=True)


@pytest.mark.slow
@pytest.mark.parametrize("mode", ["auto", "pwrite", "mmap"])
def test_device_write_stress_both_file_modes(golden, tmp_path, mode, monkeypatch):
    """seeds covering P = 1/2/3/8, every file mode, each script repeated; every file hashed
    against the reference-made golden.  PGSD_STRESS_REPEAT raises the repeat count (profiles/r3_device_write_stress.txt
    records a 200-iteration run)."""
    monkeypatch.setenv("PGSD_B200_FILE_MODE", mode)
    sums = json.load(open(os.path.join(golden, "script_sha256.json")))
    repeat = int(os.environ.get("PGSD_STRESS_REPEAT", "2"))
    bad = []
    for rep in range(repeat):
        for seed in (1, 2, 3, 5, 7, 10):
            P = seed_nprocs(seed)
            gsd, prefix = opscript.run_replay(random_script(seed, P, lookups=(seed % 3 != 0)), str(tmp_path),
                                              f"s{seed}_{rep}", P, device=True, soa=(seed % 2 == 0), timeout=300)
            if hashlib.sha256(opscript.read_bytes(gsd)).hexdigest() != sums[str(seed)]["gsd"]:
                bad.append((seed, P, rep))
            os.unlink(gsd)
    assert bad == [], bad
