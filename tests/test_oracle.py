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
