"""Write-path parity on the host-pointer path: files written by libpgsd_b200 are byte-identical to
the reference's for the same call sequence and rank partitioning (no GPU needed).

Three sources of truth, all produced by the UNMODIFIED reference pgsd.c:
  * committed golden files / hashes (tests/golden, made by make_golden.py),
  * the compiled reference run live (oracle/_ref/ref_driver) when it is present.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import opscript
from golden.make_golden import SCRIPT_SEEDS, hoomd_script, seed_nprocs
from opscript import READONLY, READWRITE, Script
from randscript import random_script

pytestmark = pytest.mark.skipif(not opscript.have_replay(), reason="tools/pgsd_replay not built")
needs_ref = pytest.mark.skipif(not opscript.have_reference(), reason="oracle/_ref/ref_driver not built")


@pytest.mark.parametrize("P", [1, 2, 3, 8])
def test_hoomd_frames_match_golden(golden, tmp_path, P):
    gsd, prefix = opscript.run_replay(hoomd_script(), str(tmp_path), f"h{P}", P)
    assert opscript.read_bytes(gsd) == opscript.read_bytes(os.path.join(golden, f"hoomd_p{P}.gsd"))
    assert opscript.read_bytes(prefix + ".log") == opscript.read_bytes(os.path.join(golden, f"hoomd_p{P}.log"))


@pytest.mark.parametrize("P", [2, 8])
def test_auto_offset_matches_golden(golden, tmp_path, P):
    """PGSD_B200_OFFSET_AUTO (library-side K2 prefix) lays the file out like caller-side offsets."""
    gsd, _ = opscript.run_replay(hoomd_script(), str(tmp_path), f"a{P}", P, auto_offset=True)
    assert opscript.read_bytes(gsd) == opscript.read_bytes(os.path.join(golden, f"hoomd_p{P}.gsd"))


@pytest.mark.parametrize("seed", SCRIPT_SEEDS)
def test_random_scripts_match_golden_hashes(golden, tmp_path, seed):
    sums = json.load(open(os.path.join(golden, "script_sha256.json")))[str(seed)]
    P = seed_nprocs(seed)
    assert sums["nprocs"] == P
    gsd, prefix = opscript.run_replay(random_script(seed, P, lookups=(seed % 3 != 0)), str(tmp_path), f"s{seed}", P)
    assert os.path.getsize(gsd) == sums["bytes"]
    assert hashlib.sha256(opscript.read_bytes(gsd)).hexdigest() == sums["gsd"]
    assert hashlib.sha256(opscript.read_bytes(prefix + ".log")).hexdigest() == sums["log"]


@needs_ref
@pytest.mark.parametrize("seed", range(1000, 1032))
def test_random_scripts_match_live_reference(tmp_path, seed):
    P = seed_nprocs(seed)
    s = random_script(seed, P, lookups=(seed % 3 != 0))
    ref = opscript.run_reference(s, str(tmp_path), "r", P)
    new = opscript.run_replay(s, str(tmp_path), "r", P)
    opscript.compare_runs(ref, new, P)


def _dtype_script():
    # ref case list: test_fl.py:29-88 (dtype matrix), :399-429 (zero-size chunk)
    s = Script()
    s.create(application="test_dtype", schema="none", schema_version=(1 << 16) | 2)
    rng = np.random.default_rng(7)
    for dt in (np.uint8, np.uint16, np.uint32, np.uint64, np.int8, np.int16, np.int32, np.int64, np.float32,
               np.float64):
        s.chunk(f"data1d/{np.dtype(dt).name}", rng.integers(0, 100, size=9).astype(dt), True, "S")
        s.chunk(f"data2d/{np.dtype(dt).name}", rng.integers(0, 100, size=(5, 2)).astype(dt), True, "S")
        s.chunk(f"zero/{np.dtype(dt).name}", np.zeros((0, 3), dtype=dt), True, "S")
    s.end_frame()
    s.close()
    s.open(READONLY)
    for dt in ("uint8", "int64", "float32", "float64"):
        s.read(0, f"data1d/{dt}", 0)
        s.read(0, f"data2d/{dt}", 1)
        s.find(0, f"zero/{dt}")
        s.read(0, f"zero/{dt}", 0)
    s.close()
    return s


def _many_names_script(P):
    # ref case list: test_fl.py:574-610 (63-char names), :863-893 (1000 names x 5 frames shuffled)
    s = Script()
    s.create()
    rng = np.random.default_rng(3)
    names = [f"name_{i:04d}_" + "y" * (i % 50) for i in range(700)]
    names.append("z" * 63)
    names.append("w" * 100)  # longer than PGSD_NAME_SIZE: v2 stores it whole
    for frame in range(3):
        order = rng.permutation(len(names))
        for k in order[:400]:
            s.chunk(names[k], np.array([frame * 1000 + k], dtype=np.int32), False, "R")
        s.end_frame()
    s.close()
    s.open(READONLY)
    s.nnames()
    s.find(2, names[int(order[0])])
    s.read(1, "z" * 63, 0)
    s.match("name_01")
    s.close()
    return s


def _index_growth_script(P, frames=120):
    # many small frames: the index outgrows 128 entries several times and relocates (Q7)
    s = Script()
    s.create()
    s.setidx(50)
    for f in range(frames):
        s.chunk("configuration/step", np.array([f], dtype=np.uint64), False, "R")
        s.chunk("particles/position", np.full((7, 3), f, dtype=np.float32), True, "S")
        s.chunk("log/value/a", np.array([f * 0.5], dtype=np.float32), False, "R")
        if f % 3 == 0:
            s.chunk("log/value/b", np.array([f], dtype=np.float64), False, "R")
        s.end_frame()
    s.close()
    s.open(READONLY)
    s.nframes()
    s.find(frames - 1, "particles/position")
    s.read(frames - 1, "particles/position", 0)
    s.read(frames // 2, "log/value/a", 0)
    s.close()
    return s


def _buffered_only_script(P):
    # frames that hold only buffered chunks are committed late (index_entries_to_buffer, pgsd.c:1942)
    s = Script()
    s.create()
    s.setidx(10)
    for f in range(30):
        s.chunk("log/value/a", np.array([f], dtype=np.float32), False, "R")
        s.chunk("log/value/b", np.arange(5, dtype=np.int16) + f, False, "R")
        s.end_frame()
    s.nframes()
    s.close()
    return s


def _big_unbuffered_script(P):
    # all=false chunk at least maximum_write_buffer_size big: only rank 0 writes, file grows by the SUM (Q4)
    s = Script()
    s.create()
    s.setbuf(256)
    s.chunk("big/replicated", np.arange(200, dtype=np.float64), False, "R")
    s.chunk("small", np.arange(3, dtype=np.int32), False, "R")
    s.chunk("big/split", np.arange(300, dtype=np.float32).reshape(100, 3), True, "S")
    s.end_frame()
    s.chunk("small", np.arange(3, dtype=np.int32) + 9, False, "R")
    s.end_frame()
    s.close()
    return s


def _append_reopen_script(P):
    s = Script()
    s.create()
    s.chunk("a", np.arange(10, dtype=np.int32), True, "S")
    s.end_frame()
    s.close()
    s.open(READWRITE)
    s.nframes()
    s.chunk("a", np.arange(10, dtype=np.int32) + 100, True, "S")
    s.chunk("b", np.arange(4, dtype=np.float32), False, "R")
    s.end_frame()
    s.close()
    s.open(opscript.APPEND)
    s.chunk("c", np.arange(6, dtype=np.uint8), True, "S")
    s.end_frame()
    s.nframes()
    s.close()
    s.open(READONLY)
    s.read(1, "a", 0)
    s.read(2, "c", 0)
    s.find(0, "b")
    s.close()
    return s


SCENARIOS = {
    "dtypes": lambda P: _dtype_script(),
    "many_names": _many_names_script,
    "index_growth": _index_growth_script,
    "buffered_only": _buffered_only_script,
    "big_unbuffered": _big_unbuffered_script,
    "append_reopen": _append_reopen_script,
}


@needs_ref
@pytest.mark.parametrize("P", [1, 2, 8])
@pytest.mark.parametrize("scenario", sorted(SCENARIOS))
def test_scenarios_match_live_reference(tmp_path, scenario, P):
    s = SCENARIOS[scenario](P)
    ref = opscript.run_reference(s, str(tmp_path), "r", P)
    new = opscript.run_replay(s, str(tmp_path), "r", P)
    opscript.compare_runs(ref, new, P)


@needs_ref
def test_config5_shape_many_small_frames(tmp_path):
    """BASELINE config 5 in miniature: N=4096 frames with 8 log scalars, index relocations."""
    from pgsd_sph_b200 import synth
    P = 8
    s = Script()
    s.create()
    fr = synth.make_frame(4096, 0)
    for f in range(60):
        logs = [(f"log/value/q{k}", np.array([f + 0.25 * k], dtype=np.float32)) for k in range(8)]
        s.hoomd_frame(fr, synth.frame_scalars(4096, f), logs)
    s.close()
    ref = opscript.run_reference(s, str(tmp_path), "r", P)
    new = opscript.run_replay(s, str(tmp_path), "r", P)
    size = opscript.compare_runs(ref, new, P)
    assert size > 60 * 4096 * 40
