"""Host-path tests of the drop-in Python layer (pgsd_sph_b200.fl / .hoomd) on CPU.

Case list re-targeted from the reference's (upstream-GSD) tests, which never exercise PGSD itself
(SURVEY.md section 4): dtype round trip, metadata, chunk_exists, read-only errors, bad dtypes, name
length, open modes, name matching, zero-size chunks (/root/reference/pgsd/pgsd/test/test_fl.py:29-88,
399-429,574-610,863-893) and the frame-0 / default fallback rules, slicing and log reading
(test_hoomd.py:57-166,297-529).  Decoding is checked against the oracle reader on reference-made goldens."""
import os
import pickle

import numpy as np
import pytest

from oracle import reader_oracle
from pgsd_sph_b200 import fl, hoomd

DTYPES = [np.uint8, np.uint16, np.uint32, np.uint64, np.int8, np.int16, np.int32, np.int64, np.float32, np.float64]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(1,), (7,), (5, 3), (4, 1)])
def test_dtype_round_trip(tmp_path, dt, shape):
    rng = np.random.default_rng(3)
    a = (rng.standard_normal(shape) * 100).astype(dt)
    p = str(tmp_path / "t.gsd")
    with fl.open(p, 'w', 'app', 'schema', [1, 2]) as f:
        f.write_chunk('chunk', a)
        f.end_frame()
    with fl.open(p, 'r') as f:
        b = f.read_chunk(0, 'chunk')
        assert b.dtype == np.dtype(dt)
        assert b.tobytes() == a.tobytes()
        assert b.shape == (a.shape if not (a.ndim == 2 and a.shape[1] == 1) else (a.shape[0],))


def test_metadata_modes_and_errors(tmp_path):
    p = str(tmp_path / "m.gsd")
    with fl.open(p, 'x', 'my app', 'my schema', [3, 7]) as f:
        assert (f.mode, f.name, f.application, f.schema, f.schema_version) == ('x', p, 'my app', 'my schema', (3, 7))
        assert f.pgsd_version == (2, 0) and f.nframes == 0
        for i in range(5):
            f.write_chunk('a', np.array([i], dtype=np.int32))
            f.write_chunk('b/c', np.arange(3, dtype=np.float64) + i)
            f.end_frame()
        assert f.nframes == 5 and f.nnames == 2
        assert f.chunk_exists(4, 'b/c') and not f.chunk_exists(5, 'a') and not f.chunk_exists(0, 'zz')
        assert f.find_matching_chunk_names('') == ['a', 'b/c'] and f.find_matching_chunk_names('b') == ['b/c']
        with pytest.raises(KeyError):
            f.read_chunk(0, 'missing')
        with pytest.raises(ValueError):
            f.write_chunk('bad', np.array(['x', 'y']))           # unsupported dtype
        with pytest.raises(ValueError):
            f.write_chunk('bad', np.zeros((2, 2, 2), dtype=np.float32))  # > 2 dimensions
    with pytest.raises(FileExistsError):
        fl.open(p, 'x', 'a', 's', [1, 0])
    with pytest.raises(FileNotFoundError):
        fl.open(str(tmp_path / "nope.gsd"), 'r')
    with pytest.raises(ValueError):
        fl.open(p, 'q')
    with pytest.raises(RuntimeError):
        fl.open(p, 'r', schema='other schema')
    f = fl.open(p, 'r')
    with pytest.raises(RuntimeError):
        f.write_chunk('a', np.array([1], dtype=np.int32))        # file must be writable
    g = pickle.loads(pickle.dumps(f))                              # read-only files pickle
    assert g.read_chunk(3, 'a')[0] == 3
    g.close()
    f.close()
    with pytest.raises(ValueError):
        f.nframes
    with fl.open(p, 'a') as f:                                     # append continues at frame 5
        f.write_chunk('a', np.array([99], dtype=np.int32))
        f.end_frame()
        assert f.nframes == 6 and f.read_chunk(5, 'a')[0] == 99


def test_long_names_and_zero_size_chunk(tmp_path):
    p = str(tmp_path / "n.gsd")
    long_name = 'n' * 80
    with fl.open(p, 'w', 'app', 'schema', [1, 0]) as f:
        f.write_chunk(long_name, np.arange(4, dtype=np.uint16))
        f.write_chunk('empty', np.zeros((0, 3), dtype=np.float32))
        f.end_frame()
    with fl.open(p, 'r') as f:
        # v2 namelists store NUL-separated names of any length (pgsd.c:1340-1404); only v1 files cut at 63
        assert f.find_matching_chunk_names('n') == [long_name]
        assert f.read_chunk(0, long_name).tolist() == [0, 1, 2, 3]
        e = f.read_chunk(0, 'empty')
        assert e.shape == (0, 3) and e.dtype == np.float32


@pytest.mark.parametrize("P", [1, 2, 8])
def test_hoomd_decode_matches_oracle_reader(golden, P):
    path = os.path.join(golden, f"hoomd_p{P}.gsd")
    orc = reader_oracle.OracleFile(path)
    with hoomd.open(path, 'r') as t:
        assert len(t) == orc.nframes
        frames = [t[i] for i in range(len(t))]
        for i, fr in enumerate(frames):
            ref = reader_oracle.decode_particles(orc, i)
            assert int(fr.particles.N) == ref['N']
            assert fr.configuration.step == 10 * i
            for name in ('position', 'velocity', 'typeid', 'density', 'pressure'):
                assert getattr(fr.particles, name).tobytes() == ref[name].tobytes(), (i, name)
            assert fr.log['particles/id'].tobytes() == ref['log/particles/id'].tobytes()
            assert fr.log['value/kinetic_energy'][0] == np.float32(0.5 * i + 1.25)
        # views and iteration
        assert [f.configuration.step for f in t[1:]] == [10 * i for i in range(1, len(t))]
        assert [f.configuration.step for f in t[::-1]] == [10 * i for i in reversed(range(len(t)))]
        assert t[-1].configuration.step == 10 * (len(t) - 1)
        with pytest.raises(IndexError):
            t[len(t)]
    logs = hoomd.read_log(path, scalar_only=True)
    assert logs['configuration/step'].tolist() == [10 * i for i in range(len(frames))]
    assert np.allclose(logs['log/value/potential_energy'], [-3.0 * i for i in range(len(frames))])


def test_hoomd_append_defaults_and_frame0_fallback(tmp_path):
    """Fields absent from a frame come from frame 0 when N matches, else from the schema default
    (ref: hoomd.py:852-881)."""
    p = str(tmp_path / "h.gsd")
    n = 50
    rng = np.random.default_rng(0)
    pos0 = rng.random((n, 3)).astype(np.float32)
    with hoomd.open(p, 'w') as t:
        f0 = hoomd.Frame()
        f0.configuration.step = 5
        f0.configuration.box = [3, 4, 5, 0, 0, 0]
        f0.particles.N = n
        f0.particles.types = ['fluid', 'wall']
        f0.particles.position = pos0.astype(np.float64)     # validate() casts to float32
        f0.particles.typeid = np.arange(n) % 2
        f0.log['value/e'] = np.array([1.5], dtype=np.float32)
        t.append(f0)
        f1 = hoomd.Frame()
        f1.configuration.step = 6
        f1.particles.N = n
        f1.particles.velocity = np.ones((n, 3), dtype=np.float32)
        t.append(f1)
        f2 = hoomd.Frame()
        f2.configuration.step = 7
        f2.particles.N = n + 1                              # N differs: defaults, not frame 0
        t.append(f2)
    with hoomd.open(p, 'r') as t:
        a, b, c = t[0], t[1], t[2]
        assert a.particles.types == ['fluid', 'wall'] and b.particles.types == ['fluid', 'wall']
        assert a.particles.position.dtype == np.float32 and a.particles.position.tobytes() == pos0.tobytes()
        assert (a.particles.velocity == 0).all() and a.particles.velocity.shape == (n, 3)      # default
        assert b.particles.position.tobytes() == pos0.tobytes()                                 # frame-0 fallback
        assert (b.particles.typeid == np.arange(n) % 2).all()
        assert (b.particles.velocity == 1).all()
        assert b.configuration.box.tolist() == [3, 4, 5, 0, 0, 0]
        assert c.particles.position.shape == (n + 1, 3) and (c.particles.position == 0).all()   # default, N changed
        assert (c.particles.typeid == 0).all()
        assert a.log['value/e'][0] == np.float32(1.5) and b.log['value/e'][0] == np.float32(1.5)


def test_vtu_point_arrays_and_file(golden, tmp_path):
    """pgsd2vtu input preparation == the manual's numpy.ascontiguousarray(col, float64) (pgsd.tex:1249-1259); the
    .vtu container == the sequential restatement of pyevtk's writer in oracle/vtu_oracle.py (parity with pyevtk
    itself UNPINNED: pyevtk is absent and the reference names no version)."""
    from oracle import vtu_oracle
    from pgsd_sph_b200 import vtu
    with hoomd.open(os.path.join(golden, "hoomd_p2.gsd"), 'r') as t:
        fr = t[1]
        x, y, z, pd = vtu.point_arrays(fr)
        assert x.dtype == np.float64 and x.flags.c_contiguous
        assert x.tobytes() == np.ascontiguousarray(fr.particles.position[:, 0], dtype=np.float64).tobytes()
        assert pd['velocity'][2].tobytes() == np.ascontiguousarray(fr.particles.velocity[:, 2], dtype=np.float64).tobytes()
        assert pd['slength'].shape == (int(fr.particles.N),) and (pd['slength'] == 1).all()  # schema default (hoomd.py:178)
        name = vtu.write_vtu(str(tmp_path / "f_00001"), x, y, z, pd)
    assert name.endswith("f_00001.vtu")
    raw = open(name, 'rb').read()
    assert raw == vtu_oracle.points_to_vtk_bytes(x, y, z, pd)
    head, tail = raw.split(b'<AppendedData encoding="raw">\n_', 1)
    assert b'NumberOfPoints="512"' in head and b'Name="velocity" NumberOfComponents="3"' in head
    n = 512
    first = int.from_bytes(tail[:8], 'little')
    assert first == n * 3 * 8
    assert np.frombuffer(tail[8:8 + first], dtype=np.float64).reshape(n, 3)[:, 1].tobytes() == y.tobytes()
    back = vtu_oracle.parse_vtu(raw)
    assert back['points'][:, 2].tobytes() == z.tobytes() and back['density'].tobytes() == pd['density'].tobytes()
    assert (back['connectivity'] == np.arange(n)).all() and (back['offsets'] == np.arange(1, n + 1)).all()
    assert (back['types'] == 1).all() and back['velocity'][:, 0].tobytes() == pd['velocity'][0].tobytes()


@pytest.mark.parametrize("n", [0, 1, 5, 4097, 1 << 20])
def test_vtu_container_sizes_and_dtypes(tmp_path, n):
    """Empty, tiny, page-straddling and multi-piece (> 16 MiB: several writer threads) files; float32 inputs, a
    file without point data, and a second frame reusing the cached image."""
    from oracle import vtu_oracle
    from pgsd_sph_b200 import vtu
    rng = np.random.default_rng(n)
    for rep, dt in enumerate((np.float64, np.float32, np.float64)):
        x, y, z = (rng.standard_normal(n).astype(dt) for _ in range(3))
        pd = {'vel': tuple(rng.standard_normal(n).astype(dt) for _ in range(3)), 'rho': rng.random(n).astype(dt),
              'tag': np.arange(n, dtype=np.int32)[::-1].copy()}
        name = vtu.write_vtu(str(tmp_path / f"a{rep}"), x, y, z, pd)
        assert open(name, 'rb').read() == vtu_oracle.points_to_vtk_bytes(x, y, z, pd)
    name = vtu.write_vtu(str(tmp_path / "bare.vtu"), x, y, z)
    assert name.endswith("bare.vtu") and open(name, 'rb').read() == vtu_oracle.points_to_vtk_bytes(x, y, z, None)
    with pytest.raises(ValueError):
        vtu.write_vtu(str(tmp_path / "bad"), x, y, z[:-1] if n else np.zeros(1))


def test_reorder_distributed_plan_tiles_the_id_space():
    """Host-side ownership of pgsd_b200_reorder_distributed (include/pgsd_b200.h): the ranks' id ranges are
    consecutive, equally sized, cover every id of a dense frame, and rank order is id order."""
    import ctypes as C
    from pgsd_sph_b200 import _lib
    lib = _lib.load()
    for n in (1, 2, 1023, 1024, 1025, 5000, 300001, 1 << 20, (1 << 24) + 5, 64 << 20, 100 << 20):
        for ranks in (1, 2, 3, 4, 8):
            spans = []
            for r in range(ranks):
                first, rows = C.c_uint64(), C.c_uint64()
                assert lib.pgsd_b200_reorder_distributed_plan(n, ranks, r, C.byref(first), C.byref(rows)) == 0
                spans.append((first.value, rows.value))
            size = spans[0][1]
            cap = 1024 if n <= (32 << 20) else (2048 if n <= (64 << 20) else 4096)
            assert size % cap == 0 and all(s == (r * size, size) for r, s in enumerate(spans))
            assert ranks * size >= n                      # every id 0..n-1 has an owner
            assert (ranks * size - n) < ranks * cap       # balanced to one bucket per rank
    bad = C.c_uint64()
    assert lib.pgsd_b200_reorder_distributed_plan(0, 2, 0, C.byref(bad), C.byref(bad)) < 0
    assert lib.pgsd_b200_reorder_distributed_plan(1 << 32, 2, 0, C.byref(bad), C.byref(bad)) < 0
    assert lib.pgsd_b200_reorder_distributed_plan(1000, 9, 0, C.byref(bad), C.byref(bad)) < 0
    assert lib.pgsd_b200_reorder_distributed_plan(1000, 2, 2, C.byref(bad), C.byref(bad)) < 0


def test_communicator_cannot_be_replaced_while_files_are_open(tmp_path):
    """Open handles keep a pointer to the communicator they were opened under: replacing it is refused
    until they are closed."""
    from pgsd_sph_b200 import _lib
    lib = _lib.load()
    f = fl.open(str(tmp_path / "c.gsd"), 'w', 'pgsd-b200', 'hoomd', [1, 4])
    assert lib.pgsd_b200_comm_finalize() == -2
    assert b"still open" in lib.pgsd_b200_last_error()
    f.write_chunk("a", np.arange(4, dtype=np.uint32), write_all=False)
    f.end_frame()
    f.close()
    assert lib.pgsd_b200_comm_finalize() == 0


def test_prepared_host_chunks_write_the_same_file(tmp_path):
    """PGSDFile.prepare_chunks / write_prepared == write_chunk per chunk, with buffers updated in place."""
    n, frames = 100, 40
    step, box = np.zeros(1, np.uint64), np.array([10, 10, 10, 0, 0, 0], np.float32)
    pos = np.zeros((n, 3), np.float32)
    logs = [("log/value/v%d" % k, np.array([k], dtype=np.float32)) for k in range(3)]
    a, b = str(tmp_path / "a.gsd"), str(tmp_path / "b.gsd")
    with fl.open(a, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
        prep = f.prepare_chunks([("configuration/step", step, None, False), ("configuration/box", box, None, False),
                                 ("particles/position", pos, None, True)] + [(k, v, None, False) for k, v in logs])
        for i in range(frames):
            step[0] = 10 * i
            pos[:] = i
            f.write_prepared(prep)
            f.end_frame()
    with fl.open(b, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
        for i in range(frames):
            f.write_chunk("configuration/step", np.array([10 * i], np.uint64), write_all=False)
            f.write_chunk("configuration/box", box, write_all=False)
            f.write_chunk("particles/position", np.full((n, 3), i, np.float32))
            for k, v in logs:
                f.write_chunk(k, v, write_all=False)
            f.end_frame()
    assert open(a, "rb").read() == open(b, "rb").read()
    with pytest.raises(ValueError):
        with fl.open(a, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
            f.prepare_chunks([("x", np.zeros((4, 4))[:, ::2], None, False)])
