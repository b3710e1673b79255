"""GPU parity tests of the cluster path of the reorder (pgsd_sph_b200/csrc/kernels_cluster.cu: coarse partition into
write-combined runs + placement by thread-block clusters), called through the C ABI, against numpy's stable argsort
+ gather (oracle/reorder_oracle.py states the same rule).  Bit-exact.  The path is opt-in (PGSD_B200_CLUSTER=1: it
measured slower than the slot path) and takes frames from 1 Mi rows on;
PGSD_B200_CLUSTER_MIN_ROWS=0 sends small frames down it as well."""
import numpy as np
import pytest

from pgsd_sph_b200 import _lib
from pgsd_sph_b200.devmem import DeviceArray
from test_gpu_kernels import _launches, _reorder_device_full, slot_key_cases

pytestmark = pytest.mark.gpu

# coarse scatter + cluster placement with the key range guessed from n; a wrong guess (ids with an offset or gaps)
# costs that attempt + census + the two kernels again (5), or + the four kernels of the slot path when the measured
# range is too sparse for the cluster geometry (7); anything more = the general path ran
CLUSTER_LAUNCHES = (2, 5, 7)


@pytest.fixture(scope="module")
def cuda(lib):
    assert lib.pgsd_b200_cuda_available() == 1, "no CUDA device: the device path has no CPU fallback"
    _lib.check(lib.pgsd_b200_device_init(0), "device_init")
    return lib


@pytest.fixture(autouse=True)
def cluster_on(monkeypatch):
    monkeypatch.setenv("PGSD_B200_CLUSTER", "1")


def _fields(n, rng):
    return [rng.standard_normal((n, 3)).astype(np.float32), rng.standard_normal((n, 3)).astype(np.float32),
            rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32),
            rng.integers(0, 3, size=n).astype(np.uint32)]


def _check(cuda, keys, fields, want_perm):
    o = np.argsort(keys, kind='stable')
    cuda.pgsd_b200_reset_stats()
    s, p, outs = _reorder_device_full(cuda, keys, fields, want_perm)
    launches = _launches(cuda)
    assert (s == keys[o]).all()
    if want_perm:
        assert (p == o.astype(np.uint32)).all()
    for f, g in zip(fields, outs):
        assert g.tobytes() == f[o].tobytes()
    return launches


@pytest.mark.parametrize("name,keys,slot", list(slot_key_cases()), ids=[k for k, _, _ in slot_key_cases()])
def test_reorder_cluster_path_equals_stable_argsort(cuda, monkeypatch, name, keys, slot):
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    monkeypatch.setenv("PGSD_B200_CLUSTER_MIN_ROWS", "0")
    n = len(keys)
    fields = _fields(n, np.random.default_rng(n))
    for want_perm in (True, False):
        launches = _check(cuda, keys, fields, want_perm)
        if slot is True:
            assert launches in CLUSTER_LAUNCHES, launches
        elif slot is False:
            assert launches > 9, launches


@pytest.mark.parametrize("bulk", ["1", "0"])
@pytest.mark.parametrize("agg", ["1", "0"])
@pytest.mark.parametrize("tile,threads", [("1024", "256"), ("2048", "512"), ("2048", "256"), ("4096", "1024"), ("4096", "512")])
@pytest.mark.parametrize("bits", ["10", "11", "12"])
def test_reorder_cluster_path_variants(cuda, monkeypatch, bits, tile, threads, agg, bulk):
    """Slots per CTA of the placement cluster (1024 / 2048 / 4096), scatter tile and CTA sizes, plain shared-memory
    atomics instead of warp-aggregated ones, plain-load staging instead of bulk copies."""
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    monkeypatch.setenv("PGSD_B200_CLUSTER_MIN_ROWS", "0")
    monkeypatch.setenv("PGSD_B200_CLUSTER_BITS", bits)
    monkeypatch.setenv("PGSD_B200_CLUSTER_TILE", tile)
    monkeypatch.setenv("PGSD_B200_CLUSTER_THREADS", threads)
    monkeypatch.setenv("PGSD_B200_CLUSTER_AGG", agg)
    monkeypatch.setenv("PGSD_B200_SLOT_BULK", bulk)
    rng = np.random.default_rng(int(bits) * 7 + int(tile) + int(threads))
    n = 150001
    keys = (rng.permutation(n) + rng.integers(0, 2)).astype(np.uint32)
    fields = [rng.integers(0, 2 ** 32, size=(n, 3), dtype=np.uint64).astype(np.uint32), rng.standard_normal(n),
              rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)]
    assert _check(cuda, keys, fields, True) in CLUSTER_LAUNCHES


@pytest.mark.parametrize("widths", [(1,), (2, 2), (3, 3, 1, 1, 1), (4, 4, 4, 4), (5, 7), (16, 14), (16, 16)])
def test_reorder_cluster_path_row_widths(cuda, monkeypatch, widths):
    """Even and odd record widths (8-byte and 4-byte lanes), records up to 32 words (fewer slots per CTA)."""
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    monkeypatch.setenv("PGSD_B200_CLUSTER_MIN_ROWS", "0")
    rng = np.random.default_rng(sum(widths))
    n = 70001
    keys = rng.permutation(n).astype(np.uint32)
    fields = [rng.integers(0, 2 ** 32, size=(n, w), dtype=np.uint64).astype(np.uint32) for w in widths]
    for want_perm in (True, False):
        launches = _check(cuda, keys, fields, want_perm)
        if 1 + sum(widths) + (1 if want_perm else 0) <= 32:
            assert launches in CLUSTER_LAUNCHES, launches


@pytest.mark.parametrize("seed", range(24))
def test_reorder_cluster_path_random_geometry(cuda, monkeypatch, seed):
    """Seeded random cases: n, id range (10..27 bits), density, offset, record shape, permutation wanted or not."""
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    monkeypatch.setenv("PGSD_B200_CLUSTER_MIN_ROWS", "0")
    rng = np.random.default_rng(7000 + seed)
    bits = int(rng.integers(10, 28))
    n = int(min(rng.integers(1, 600000), 1 << bits))
    keys = rng.choice(1 << bits, size=n, replace=False).astype(np.uint32)
    if rng.random() < 0.3:
        keys = (keys + np.uint32(rng.integers(1, 1 << 30))).astype(np.uint32)
    if rng.random() < 0.2 and n > 10:
        keys[int(rng.integers(0, n))] = keys[int(rng.integers(0, n))]
    widths = [int(w) for w in rng.integers(1, 6, size=int(rng.integers(1, 6)))]
    fields = [rng.integers(0, 2 ** 32, size=(n, w), dtype=np.uint64).astype(np.uint32) for w in widths]
    _check(cuda, keys, fields, bool(rng.integers(0, 2)))


def big_cases():
    rng = np.random.default_rng(99)
    n = (1 << 21) + 12345
    yield "perm_2Mi_ragged", rng.permutation(n).astype(np.uint32), True
    yield "sorted_2Mi", np.arange(1 << 21, dtype=np.uint32), True                       # every tile is one run
    yield "reversed_blocks", np.arange(1 << 21, dtype=np.uint32).reshape(-1, 4096)[::-1].ravel().copy(), True
    yield "even_ids_1Mi", (rng.permutation(1 << 20) * 2).astype(np.uint32), True        # half-empty slots: compaction in every CTA
    yield "offset_ids", (rng.permutation(1 << 20) + 1).astype(np.uint32), True          # guess misses by one key
    yield "one_dup", np.concatenate([rng.permutation((1 << 20) + 7), [31337]]).astype(np.uint32), False
    yield "many_dups", rng.integers(0, 1 << 20, size=1 << 20).astype(np.uint32), False
    k = rng.permutation(1 << 20).astype(np.uint32)
    k[:40000] = k[0] >> 15 << 15                                                       # one coarse bucket overflows its region
    yield "bucket_overflow", k, False


@pytest.mark.parametrize("name,keys,unique", list(big_cases()), ids=[k for k, _, _ in big_cases()])
def test_reorder_cluster_path_production_sizes(cuda, name, keys, unique):
    """Frames of >= 1 Mi rows take the cluster path once it is switched on (no other environment switches)."""
    n = len(keys)
    fields = _fields(n, np.random.default_rng(n))
    for want_perm in (False, True):
        launches = _check(cuda, keys, fields, want_perm)
        if unique:
            assert launches in CLUSTER_LAUNCHES, launches
        else:
            assert launches > 9, launches


def test_reorder_cluster_path_unaligned_inputs(cuda, monkeypatch):
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    monkeypatch.setenv("PGSD_B200_CLUSTER_MIN_ROWS", "0")
    rng = np.random.default_rng(78)
    n = 50000
    keys = rng.permutation(n).astype(np.uint32)
    pos = rng.standard_normal((n, 3)).astype(np.float32)
    dens = rng.standard_normal(n).astype(np.float32)
    dk = DeviceArray.from_numpy(keys)
    dpos = DeviceArray.from_numpy(np.concatenate([np.zeros(1, np.float32), pos.ravel()]))
    ddens = DeviceArray.from_numpy(np.concatenate([np.zeros(1, np.float32), dens]))
    opos, odens = DeviceArray((n * 3 + 1,), np.float32), DeviceArray((n + 1,), np.float32)
    ds = DeviceArray((n,), np.uint32)
    fl_ = (_lib.Field * 2)(_lib.Field(dpos.ptr + 4, opos.ptr + 4, 12), _lib.Field(ddens.ptr + 4, odens.ptr + 4, 4))
    cuda.pgsd_b200_reset_stats()
    _lib.check(cuda.pgsd_b200_reorder_device(n, dk.ptr, ds.ptr, None, 2, fl_, None), "reorder_device")
    _lib.check(cuda.pgsd_b200_synchronize(), "sync")
    assert _launches(cuda) in CLUSTER_LAUNCHES
    o = np.argsort(keys, kind='stable')
    assert (ds.to_numpy() == keys[o]).all()
    assert opos.to_numpy()[1:].tobytes() == pos[o].tobytes()
    assert odens.to_numpy()[1:].tobytes() == dens[o].tobytes()


def test_bounded_mbarrier_wait_reports_a_lost_bulk_copy(cuda):
    """The reorder kernels wait for their TMA bulk copies with a cycle bound and report flag 3 instead of hanging
    the GPU; the self-test kernel waits on a barrier whose bytes never arrive."""
    assert cuda.pgsd_b200_selftest(0) == 0
