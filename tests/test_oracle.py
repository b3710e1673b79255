"""Pin the oracles: against the reference's own known-answer file and against outputs of the
reference itself (tests/golden/make_golden.py)."""
import os
import subprocess

import numpy as np
import pytest

import opscript
from oracle import reader_oracle, reorder_oracle


def test_reader_oracle_known_answer_v1(golden):
    # ref: test_fl.py:613-651 -- 127 names '0'..'126' x 5 frames, int32 value = 13 * int(name)
    f = reader_oracle.OracleFile(os.path.join(golden, "v1_known_answer.gsd"))
    assert f.pgsd_version == (1, 0)
    assert f.nframes == 5
    names = sorted(f.find_matching_chunk_names(""), key=int)
    assert names == [str(i) for i in range(127)]
    for frame in range(5):
        for n in names:
            d = f.read_chunk(frame, n)
            assert d.dtype == np.int32 and d.shape == (1,) and d[0] == 13 * int(n)
    assert not f.chunk_exists(5, "0") and not f.chunk_exists(0, "127")


def test_reader_and_reorder_oracle_match_reference_python(golden):
    """oracle decode + argsort == the reference's pypgsd + hoomd + argsort (golden npz)."""
    ref = np.load(os.path.join(golden, "reorder_p2.npz"))
    f = reader_oracle.OracleFile(os.path.join(golden, "hoomd_p2.gsd"))
    assert f.schema == "hoomd" and f.schema_version == (1, 4) and f.application == "pgsd-b200"
    assert f.nframes == 3
    for i in range(3):
        dec = reorder_oracle.reorder_frame(reader_oracle.decode_particles(f, i))
        assert dec["N"] == int(ref[f"f{i}/N"][0])
        assert (dec["log/particles/id"] == ref[f"f{i}/id"]).all()
        assert (dec["log/particles/id"] == np.arange(dec["N"], dtype=np.uint32)).all()
        for name in reader_oracle.PARTICLE_DEFAULTS:
            a, b = dec[name], ref[f"f{i}/{name}"]
            assert a.dtype == b.dtype and a.shape == b.shape, name
            assert a.tobytes() == b.tobytes(), name
        assert f.read_chunk(i, "configuration/step")[0] == ref[f"f{i}/step"][0] == 10 * i


@pytest.mark.skipif(not opscript.have_reference(), reason="oracle/_ref/ref_driver not built")
def test_compiled_reference_reproduces_committed_goldens(golden, tmp_path):
    """The travelling oracle/_ref binary is the same reference that made tests/golden."""
    from golden.make_golden import hoomd_script
    for P in (1, 2, 8):
        gsd, prefix = opscript.run_reference(hoomd_script(), str(tmp_path), f"g{P}", P)
        assert opscript.read_bytes(gsd) == opscript.read_bytes(os.path.join(golden, f"hoomd_p{P}.gsd"))
        assert opscript.read_bytes(prefix + ".log") == opscript.read_bytes(os.path.join(golden, f"hoomd_p{P}.log"))


def test_known_layout_appendix_b(golden):
    """Layout facts recorded from the reference in SURVEY.md Appendix B hold for the goldens."""
    f = reader_oracle.OracleFile(os.path.join(golden, "hoomd_p1.gsd"))
    h = f.header
    assert int(h["index_location"]) == 256 and int(h["index_allocated_entries"]) == 128
    assert int(h["namelist_location"]) == 4352 and int(h["namelist_allocated_entries"]) == 16
    first = f.index[f.index["id"] == f.names["particles/position"]][0]
    assert int(first["location"]) == 5376 and int(first["N"]) == 512 and int(first["M"]) == 3
    # replicated small chunks: +1 copy of the buffered bytes per extra rank and frame (quirk Q1)
    sizes = {P: os.path.getsize(os.path.join(golden, f"hoomd_p{P}.gsd")) for P in (1, 2, 3, 8)}
    per_rank = sizes[2] - sizes[1]
    assert per_rank > 0 and sizes[3] - sizes[2] == per_rank and sizes[8] - sizes[1] == 7 * per_rank


def test_distributed_ownership_oracle_matches_the_library_plan():
    """The numpy restatement of the distributed reorder's ownership rule (oracle/reorder_oracle.py) and the
    library's host-only plan function agree for every frame size / rank count tried, and the oracle's shares
    tile the sorted frame."""
    import ctypes as C
    import numpy as np
    from oracle import reorder_oracle
    from pgsd_sph_b200 import _lib
    lib = _lib.load()
    for n in (1, 2, 1023, 1024, 1025, 5000, 123457, 300001, 1 << 20, (1 << 25) - 1, 1 << 25, (1 << 25) + 1,
              64 << 20, (64 << 20) + 1, 100 << 20, (1 << 27)):
        for ranks in (1, 2, 3, 4, 5, 8):
            S, cap = reorder_oracle.distributed_ownership(n, ranks)
            for r in range(ranks):
                first, rows = C.c_uint64(), C.c_uint64()
                assert lib.pgsd_b200_reorder_distributed_plan(n, ranks, r, C.byref(first), C.byref(rows)) == 0
                assert (first.value, rows.value) == (r * S, S), (n, ranks, r)
    rng = np.random.default_rng(5)
    ids = rng.permutation(70001).astype(np.uint32)
    pos = rng.standard_normal((70001, 3)).astype(np.float32)
    shares = reorder_oracle.reorder_distributed(ids, {"pos": pos}, 3)
    assert np.array_equal(np.concatenate([s[1] for s in shares]), np.arange(70001, dtype=np.uint32))
    assert np.concatenate([s[2]["pos"] for s in shares]).tobytes() == pos[np.argsort(ids, kind='stable')].tobytes()
    assert all(len(s[1]) == 0 or (s[1][0] >= s[0] and s[1][-1] < s[0] + 23552) for s in shares)   # S = ceil(69 / 3) * 1024
