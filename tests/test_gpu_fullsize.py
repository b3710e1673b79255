"""BASELINE.json full sizes on the GPU (config 3: one 64 Mi-particle frame), checked through
size-independent properties and direct byte comparison with the numpy restatement of the layout
(the reference-made goldens pin the same layout at small N)."""
import os

import numpy as np
import pytest

from oracle import cast_oracle
from pgsd_sph_b200 import fl, hoomd, synth
from pgsd_sph_b200.devmem import DeviceArray

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

N = 64 * 1024 * 1024
CHUNKS = (("particles/position", (0, 1, 2), np.float32), ("particles/velocity", (3, 4, 5), np.float32),
          ("particles/typeid", (8,), np.uint32), ("particles/density", (6,), np.float32),
          ("particles/pressure", (7,), np.float32), ("log/particles/id", (9,), np.uint32))


def _columns(n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    ids = rng.permutation(n).astype(np.uint32)
    cols = [rng.random(n, dtype=np.float32) for _ in range(8)]
    cols.append(rng.integers(0, 3, size=n, dtype=np.uint32))
    cols.append(ids)
    return cols


@pytest.fixture(scope="module")
def big_file(lib):
    assert lib.pgsd_b200_cuda_available() == 1, "no CUDA device: the device path has no CPU fallback"
    d = "/dev/shm" if os.access("/dev/shm", os.W_OK) else "/tmp"
    path = os.path.join(d, "pgsd_fullsize_%d.gsd" % os.getpid())
    cols = _columns(N, 42)
    dcols = [DeviceArray.from_numpy(c) for c in cols]
    with fl.open(path, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
        for k, a in synth.frame_scalars(N, 0):
            f.write_chunk(k, a, write_all=False)
        f.write_frame_soa(f.prepare_frame_soa([(nm, [dcols[j] for j in idx], dt, None, True) for nm, idx, dt in CHUNKS]))
        f.end_frame()
    for d_ in dcols:
        d_.free()
    yield path, cols
    os.unlink(path)


def test_64M_frame_file_bytes(big_file):
    """Direct chunks start at 5376 in call order; every chunk's bytes equal the
    host-side pack of the same columns; the buffered scalars follow."""
    path, cols = big_file
    m = np.memmap(path, dtype=np.uint8, mode='r')
    off = 5376
    for nm, idx, dt in CHUNKS:
        want = cast_oracle.pack_soa([cols[j] for j in idx], dt)
        got = m[off:off + want.nbytes]
        assert got.tobytes() == want.view(np.uint8).reshape(-1).tobytes(), nm
        off += want.nbytes
    assert off == 5376 + 40 * N
    # buffered small chunks: step u64, dimensions u8, box f32[6], N u32 = 37 bytes
    assert m.shape[0] == off + 37
    assert int(np.frombuffer(m[off + 33:off + 37].tobytes(), dtype=np.uint32)[0]) == N
    del m


def test_64M_frame_reads_back_and_reorders(big_file):
    path, cols = big_file
    with fl.open(path, 'r') as f:
        assert f.nframes == 1
        # partitioned read (r_all) of a slice in the middle
        got = f.read_chunk(0, "particles/velocity", N=1000, M=3, offset=N // 2, r_all=True)
        want = np.stack([cols[3][N // 2:N // 2 + 1000], cols[4][N // 2:N // 2 + 1000], cols[5][N // 2:N // 2 + 1000]], 1)
        assert got.tobytes() == want.tobytes()
    with hoomd.open(path, 'r', reorder='id') as t:
        fr = t[0]
        ids = fr.log['particles/id']
        assert ids[0] == 0 and ids[-1] == N - 1 and (np.diff(ids.astype(np.int64)) == 1).all()
        inv = np.empty(N, dtype=np.int64)
        inv[cols[9]] = np.arange(N)          # row of particle id in file order
        chk = np.random.default_rng(0).integers(0, N, size=200000)
        assert (fr.particles.position[chk, 0] == cols[0][inv[chk]]).all()
        assert (fr.particles.position[chk, 2] == cols[2][inv[chk]]).all()
        assert (fr.particles.velocity[chk, 1] == cols[4][inv[chk]]).all()
        assert (fr.particles.typeid[chk] == cols[8][inv[chk]]).all()
        assert (fr.particles.density[chk] == cols[6][inv[chk]]).all()
        # a permutation preserves every column's multiset: checksum of checksums
        assert int(fr.particles.typeid.astype(np.uint64).sum()) == int(cols[8].astype(np.uint64).sum())
        assert fr.particles.pressure.view(np.uint32).astype(np.uint64).sum() == cols[7].view(np.uint32).astype(np.uint64).sum()


def test_100M_reorder_widest_slot_geometry(lib):
    """More than 64 Mi particles: the slot path needs 4096-id buckets (160 KB of shared memory per bucket, one
    CTA per SM) and 25600 of them.  Size-independent checks: sorted ids are 0..N-1 and a payload that is a
    function of its id lands at row id."""
    from pgsd_sph_b200 import _lib
    assert lib.pgsd_b200_cuda_available() == 1, "no CUDA device: the device path has no CPU fallback"
    n = 100 * 1024 * 1024
    rng = np.random.Generator(np.random.PCG64(7))
    ids = rng.permutation(n).astype(np.uint32)
    tag = ids ^ np.uint32(0x5bd1e995)
    pos = np.empty((n, 3), dtype=np.float32)
    pos[:, 0] = ids
    pos[:, 1] = ids * np.float32(0.25)
    pos[:, 2] = -pos[:, 0]
    d_ids, d_tag, d_pos = DeviceArray.from_numpy(ids), DeviceArray.from_numpy(tag), DeviceArray.from_numpy(pos)
    o_ids, o_tag, o_pos = DeviceArray((n,), np.uint32), DeviceArray((n,), np.uint32), DeviceArray((n, 3), np.float32)
    fields = (_lib.Field * 2)(_lib.Field(d_tag.ptr, o_tag.ptr, 4), _lib.Field(d_pos.ptr, o_pos.ptr, 12))
    lib.pgsd_b200_reset_stats()
    _lib.check(lib.pgsd_b200_reorder_device(n, d_ids.ptr, o_ids.ptr, None, 2, fields, None), "reorder_device")
    _lib.check(lib.pgsd_b200_synchronize(), "sync")
    st = _lib.Stats()
    lib.pgsd_b200_get_stats(st)
    assert st.kernel_launches == 4, st.kernel_launches          # histogram, scan, scatter, place: the slot path
    ar = np.arange(n, dtype=np.uint32)
    assert np.array_equal(o_ids.to_numpy(), ar)
    assert np.array_equal(o_tag.to_numpy(), ar ^ np.uint32(0x5bd1e995))
    got = o_pos.to_numpy()
    assert np.array_equal(got[:, 0], ar.astype(np.float32)) and np.array_equal(got[:, 2], -ar.astype(np.float32))
    for a in (d_ids, d_tag, d_pos, o_ids, o_tag, o_pos):
        a.free()
