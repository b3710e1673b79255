"""GPU parity tests of the CUDA kernels, called through the C ABI (include/pgsd_b200.h), against the
CPU oracles (oracle/cast_oracle.py, oracle/reorder_oracle.py).  Bit-exact: integer / byte / index
work and IEEE casts whose result numpy defines exactly."""
import ctypes as C

import numpy as np
import pytest

from oracle import cast_oracle, reorder_oracle
from pgsd_sph_b200 import _lib
from pgsd_sph_b200.devmem import DeviceArray
from pgsd_sph_b200.fl import _NP_TO_PGSD
from pgsd_sph_b200.hoomd import reorder_by_id

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda(lib):
    assert lib.pgsd_b200_cuda_available() == 1, "no CUDA device: the device path has no CPU fallback"
    _lib.check(lib.pgsd_b200_device_init(0), "device_init")
    return lib


def gpu_pack(lib, columns, dst_dtype):
    """columns: list of 1-D numpy arrays (possibly strided views) of one dtype."""
    M, N = len(columns), len(columns[0])
    src_dt = columns[0].dtype
    keep, cols = [], (_lib.Column * M)()
    for j, c in enumerate(columns):
        # upload the strided view's base buffer span, keep the stride
        stride = c.strides[0] // src_dt.itemsize if N > 1 else 1
        span = np.lib.stride_tricks.as_strided(c, shape=((N - 1) * stride + 1,), strides=(src_dt.itemsize,)) if N else c
        d = DeviceArray.from_numpy(np.ascontiguousarray(span))
        keep.append(d)
        cols[j] = _lib.Column(d.ptr, stride)
    out = DeviceArray((N, M), dst_dtype)
    _lib.check(lib.pgsd_b200_pack_soa(out.ptr, _NP_TO_PGSD[np.dtype(dst_dtype)], N, M, _NP_TO_PGSD[src_dt], cols, None), "pack_soa")
    _lib.check(lib.pgsd_b200_synchronize(), "sync")
    return out.to_numpy()


INT_TYPES = [np.uint8, np.uint16, np.uint32, np.uint64, np.int8, np.int16, np.int32, np.int64]


def rand_values(rng, dt, n):
    dt = np.dtype(dt)
    if dt.kind == 'f':
        a = rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, size=n)
        return a.astype(dt)
    info = np.iinfo(dt)
    a = rng.integers(info.min, info.max, size=n, dtype=dt, endpoint=True)
    return a


@pytest.mark.parametrize("M", [1, 2, 3, 4, 7])
@pytest.mark.parametrize("N", [0, 1, 3, 4, 5, 1000, 2049, 100003])
def test_pack_f32_bitcopy(cuda, N, M):
    rng = np.random.default_rng(N * 10 + M)
    cols = [rng.standard_normal(N).astype(np.float32) for _ in range(M)]
    got = gpu_pack(cuda, cols, np.float32)
    assert got.tobytes() == cast_oracle.pack_soa(cols, np.float32).tobytes()


@pytest.mark.parametrize("M", [1, 3, 4])
@pytest.mark.parametrize("N", [1, 7, 4096, 50001])
def test_pack_f64_to_f32_special_values(cuda, N, M):
    rng = np.random.default_rng(N + M)
    cols = []
    for _ in range(M):
        a = rand_values(rng, np.float64, N)
        special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-46, 3.4028235677973366e38, 3.5e38,
                            1.0000000596046448, 1.0000001788139343, 5e-324, 1.401298464324817e-45 * 0.5])
        k = min(N, len(special))
        a[:k] = special[:k]
        if N > 20:
            # NaNs with payloads and sign
            bits = np.array([0x7ff8000000000001, 0xfff0000000000123, 0x7ff4000012345678], dtype=np.uint64)
            a[14:17] = bits.view(np.float64)
        cols.append(a)
    got = gpu_pack(cuda, cols, np.float32)
    assert got.tobytes() == cast_oracle.pack_soa(cols, np.float32).tobytes()


@pytest.mark.parametrize("src", INT_TYPES + [np.float32])
@pytest.mark.parametrize("dst", INT_TYPES + [np.float32, np.float64])
def test_pack_cast_matrix(cuda, src, dst):
    rng = np.random.default_rng(np.dtype(src).num * 100 + np.dtype(dst).num)
    N, M = 3001, 3
    cols = [rand_values(rng, src, N) for _ in range(M)]
    if np.dtype(src).kind == 'f' and np.dtype(dst).kind != 'f':
        # float -> integer: numpy defines the cast only where the truncated value fits the destination
        cols = [in_range_floats(rng, src, dst, N) for _ in range(M)]
    got = gpu_pack(cuda, cols, dst)
    assert got.dtype == np.dtype(dst)
    assert got.tobytes() == cast_oracle.pack_soa(cols, dst).tobytes()


def in_range_floats(rng, src, dst, n):
    """Floats (fractions, both signs, the extremes) whose truncation toward zero fits dst."""
    info = np.iinfo(dst)
    # float32/float64 cannot hold the 64-bit extremes exactly: stay strictly inside what the source type represents
    hi = min(float(info.max), np.nextafter(np.dtype(src).type(float(info.max) + 1.0), np.dtype(src).type(0)))
    lo = float(info.min)
    a = rng.uniform(max(lo, -1e18) - 0.0, min(hi, 1e18), size=n)
    a[: n // 4] = rng.uniform(max(lo, -300.0), min(hi, 300.0), size=n // 4)  # small magnitudes with fractions
    edge = [0.0, -0.0, 0.5, -0.5, 0.999999, -0.999999, 1.5, float(hi), float(lo), float(lo) + 0.25 if lo < 0 else 0.25]
    a[n // 4: n // 4 + len(edge)] = edge
    a = a.astype(src)
    a = np.clip(a, np.dtype(src).type(lo), np.dtype(src).type(hi))
    assert (np.trunc(a.astype(np.float64)) >= lo).all() and (np.trunc(a.astype(np.float64)) <= float(info.max)).all()
    return a


@pytest.mark.parametrize("src", [np.float32, np.float64])
@pytest.mark.parametrize("dst", INT_TYPES)
def test_pack_float_to_int_truncates_like_numpy(cuda, src, dst):
    """hoomd.py:220-266 casts whatever it is given (ascontiguousarray(dtype=uint32/int32)): values whose truncation
    fits the destination must give numpy's bytes.  Out-of-range values and NaN are undefined in numpy ("invalid
    value encountered in cast") and differ between numpy builds; not tested."""
    rng = np.random.default_rng(np.dtype(src).num * 7 + np.dtype(dst).num)
    cols = [in_range_floats(rng, src, dst, 5003) for _ in range(2)]
    got = gpu_pack(cuda, cols, dst)
    assert got.dtype == np.dtype(dst)
    assert got.tobytes() == cast_oracle.pack_soa(cols, dst).tobytes()


@pytest.mark.parametrize("stride", [2, 4, 5])
def test_pack_strided_columns_aos(cuda, stride):
    """HOOMD-style AoS source (Scalar4 pos): M columns base+j with stride 4."""
    rng = np.random.default_rng(stride)
    N = 10007
    aos = rng.standard_normal((N, stride)).astype(np.float32)
    cols = [aos[:, j] for j in range(min(3, stride))]
    got = gpu_pack(cuda, cols, np.float32)
    assert got.tobytes() == np.ascontiguousarray(aos[:, :len(cols)]).tobytes()


def test_pack_unaligned_falls_back_to_generic(cuda):
    N = 5000
    base = DeviceArray.from_numpy(np.arange(N + 1, dtype=np.float32))
    out = DeviceArray((N, 1), np.float32)
    cols = (_lib.Column * 1)(_lib.Column(base.ptr + 4, 1))  # 4-byte aligned only
    _lib.check(cuda.pgsd_b200_pack_soa(out.ptr, _lib.TYPE_FLOAT, N, 1, _lib.TYPE_FLOAT, cols, None), "pack")
    assert (out.to_numpy()[:, 0] == np.arange(1, N + 1, dtype=np.float32)).all()


# ---------------------------------------------------------------------------- K2
def test_scan_sizes_matches_numpy(cuda):
    rng = np.random.default_rng(0)
    for P, Cn in [(1, 1), (2, 10), (8, 18), (8, 5000), (3, 4097)]:
        sizes = rng.integers(0, 2 ** 40, size=(P, Cn), dtype=np.uint64)
        for rank in {0, P - 1, P // 2}:
            excl = np.zeros(Cn, np.uint64)
            total = np.zeros(Cn, np.uint64)
            mx = np.zeros(Cn, np.uint64)
            p64 = C.POINTER(C.c_uint64)
            _lib.check(cuda.pgsd_b200_scan_sizes(sizes.ctypes.data_as(p64), P, Cn, rank, excl.ctypes.data_as(p64),
                                                 total.ctypes.data_as(p64), mx.ctypes.data_as(p64)), "scan")
            assert (excl == sizes[:rank].sum(axis=0, dtype=np.uint64)).all()
            assert (total == sizes.sum(axis=0, dtype=np.uint64)).all()
            assert (mx == sizes.max(axis=0)).all()


# ---------------------------------------------------------------------------- K4 / K5
def gpu_sort(lib, keys):
    n = len(keys)
    dk = DeviceArray.from_numpy(keys)
    ds = DeviceArray((n,), np.uint32)
    dp = DeviceArray((n,), np.uint32)
    _lib.check(lib.pgsd_b200_sort_ids(n, dk.ptr, ds.ptr, dp.ptr, None), "sort_ids")
    _lib.check(lib.pgsd_b200_synchronize(), "sync")
    return ds.to_numpy(), dp.to_numpy()


def key_cases():
    rng = np.random.default_rng(11)
    yield "perm_10k", rng.permutation(10000).astype(np.uint32)
    yield "perm_1M", rng.permutation(1 << 20).astype(np.uint32)
    yield "dups_small_range", rng.integers(0, 17, size=70001).astype(np.uint32)
    yield "dups_byte_patterns", (rng.integers(0, 4, size=50000) * 0x01010101).astype(np.uint32)
    yield "full_32bit", rng.integers(0, 2 ** 32, size=300007, dtype=np.uint64).astype(np.uint32)
    yield "all_equal", np.full(12345, 77, dtype=np.uint32)
    yield "sorted", np.arange(100000, dtype=np.uint32)
    yield "reversed", np.arange(100000, dtype=np.uint32)[::-1].copy()
    yield "one", np.array([5], dtype=np.uint32)
    yield "two", np.array([9, 3], dtype=np.uint32)
    yield "tile_edge", rng.permutation(8192 * 3 + 1).astype(np.uint32)
    yield "high_bytes_only", (rng.integers(0, 256, size=40000) << 24).astype(np.uint32)
    yield "max_values", np.array([0xffffffff, 0, 0xffffffff, 1, 0xfffffffe] * 1000, dtype=np.uint32)


@pytest.mark.parametrize("name,keys", list(key_cases()), ids=[k for k, _ in key_cases()])
def test_sort_ids_equals_stable_argsort(cuda, name, keys):
    s, p = gpu_sort(cuda, keys)
    o = np.argsort(keys, kind='stable')
    assert (p == o.astype(np.uint32)).all()
    assert (s == keys[o]).all()


def test_sort_empty(cuda):
    assert cuda.pgsd_b200_sort_ids(0, None, None, None, None) == 0


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 65537])
def test_reorder_host_all_field_widths(cuda, n):
    rng = np.random.default_rng(n)
    ids = rng.permutation(n).astype(np.uint32)
    fields = {
        "position": rng.standard_normal((n, 3)).astype(np.float32),
        "typeid": rng.integers(0, 3, size=n).astype(np.uint32),
        "image": rng.integers(-5, 5, size=(n, 3)).astype(np.int32),
        "mass64": rng.standard_normal(n),
        "flags8": rng.integers(0, 255, size=n).astype(np.uint8),
        "odd_bytes": rng.integers(0, 255, size=(n, 5)).astype(np.uint8),
        "h16": rng.integers(0, 60000, size=(n, 3)).astype(np.uint16),
        "wide": rng.standard_normal((n, 16)).astype(np.float32),
    }
    sid, out = reorder_by_id(ids, fields)
    rsid, rout, _ = reorder_oracle.reorder(ids, fields)
    assert (sid == rsid).all()
    for k in fields:
        assert out[k].tobytes() == rout[k].tobytes(), k


@pytest.mark.parametrize("slot", ["1", "0"])
def test_reorder_device_resident(cuda, monkeypatch, slot):
    """Below PGSD_B200_BUCKET_MIN_ROWS: unique ids take the slot path (slot=1), otherwise pair sort + gather."""
    monkeypatch.setenv("PGSD_B200_SLOT", slot)
    rng = np.random.default_rng(5)
    n = 200003
    ids = rng.permutation(n).astype(np.uint32)
    pos = rng.standard_normal((n, 3)).astype(np.float32)
    dens = rng.standard_normal(n).astype(np.float32)
    sid, out = reorder_by_id(DeviceArray.from_numpy(ids), {"p": DeviceArray.from_numpy(pos), "d": DeviceArray.from_numpy(dens)},
                             device=True)
    o = np.argsort(ids, kind='stable')
    assert (sid.to_numpy() == ids[o]).all()
    assert out["p"].to_numpy().tobytes() == pos[o].tobytes()
    assert out["d"].to_numpy().tobytes() == dens[o].tobytes()


@pytest.mark.slow
def test_reorder_16M_size_independent_properties(cuda):
    """BASELINE config-4 size (16 Mi particles): dense unique ids -> sorted ids are 0..N-1, the
    permutation is a bijection, and a payload that encodes its own id lands at row id."""
    n = 1 << 24
    rng = np.random.default_rng(2026)
    ids = rng.permutation(n).astype(np.uint32)
    payload = np.empty((n, 3), dtype=np.float32)
    payload[:, 0] = ids
    payload[:, 1] = ids * np.float32(0.5)
    payload[:, 2] = -ids.astype(np.float32)
    tag = (ids ^ np.uint32(0x9e3779b9)).astype(np.uint32)
    sid, out = reorder_by_id(ids, {"payload": payload, "tag": tag})
    ar = np.arange(n, dtype=np.uint32)
    assert (sid == ar).all()
    assert (out["tag"] == (ar ^ np.uint32(0x9e3779b9))).all()
    assert (out["payload"][:, 0] == ar.astype(np.float32)).all()
    assert (out["payload"][:, 2] == -ar.astype(np.float32)).all()
    # checksum of checksums: a permutation preserves the multiset
    assert int(out["tag"].astype(np.uint64).sum()) == int(tag.astype(np.uint64).sum())


# ---- bucket pass (rows grouped by the top key bits before the pair passes) and both rank modes
def _reorder_device_full(lib, keys, fields, want_perm=True):
    n = len(keys)
    dk = DeviceArray.from_numpy(keys)
    ds, dp = DeviceArray((n,), np.uint32), DeviceArray((n,), np.uint32)
    din = [DeviceArray.from_numpy(f) for f in fields]
    dout = [DeviceArray(f.shape, f.dtype) for f in fields]
    fl_ = (_lib.Field * max(len(fields), 1))(*[
        _lib.Field(i.ptr, o.ptr, f.dtype.itemsize * (int(np.prod(f.shape[1:])) if f.ndim > 1 else 1))
        for i, o, f in zip(din, dout, fields)])
    _lib.check(lib.pgsd_b200_reorder_device(n, dk.ptr, ds.ptr, dp.ptr if want_perm else None, len(fields), fl_, None),
               "reorder_device")
    _lib.check(lib.pgsd_b200_synchronize(), "sync")
    return ds.to_numpy(), (dp.to_numpy() if want_perm else None), [o.to_numpy() for o in dout]


def bucket_key_cases():
    rng = np.random.default_rng(23)
    yield "perm_70k", rng.permutation(70001).astype(np.uint32)                       # 17 bits: 3 byte passes
    yield "perm_300k", rng.permutation(300000).astype(np.uint32)
    yield "dups_2bytes", rng.integers(0, 40000, size=123457).astype(np.uint32)       # duplicates: stability
    yield "one_byte_direct", rng.integers(0, 200, size=50001).astype(np.uint32)      # npass == 1: bucket pass is the sort
    yield "high_byte_direct", (rng.integers(0, 256, size=40000) << 24).astype(np.uint32)
    yield "full_32bit", rng.integers(0, 2 ** 32, size=200003, dtype=np.uint64).astype(np.uint32)
    yield "const_high_bits", (rng.integers(0, 3000, size=99999) + 0xABC00000).astype(np.uint32)
    yield "all_equal", np.full(5000, 9, dtype=np.uint32)
    yield "tile_edges", rng.permutation(4096 * 5).astype(np.uint32)
    yield "skewed", np.minimum(rng.geometric(0.001, size=150000), 2 ** 20).astype(np.uint32)
    # one bucket of ~49 tiles (warp-cooperative segment scan) next to nearly empty ones
    yield "one_big_bucket", np.concatenate([rng.integers(0, 65536, size=400000), (1 << 23) + np.arange(100),
                                            (1 << 22) + rng.integers(0, 5, size=9000)]).astype(np.uint32)
    yield "empty_low_bits", (rng.integers(0, 300, size=60000) << 20).astype(np.uint32)   # nothing varies below the bucket digit


@pytest.mark.parametrize("rank_mode", ["ballot", "match"])
@pytest.mark.parametrize("name,keys", list(bucket_key_cases()), ids=[k for k, _ in bucket_key_cases()])
def test_reorder_bucket_path_equals_stable_argsort(cuda, monkeypatch, name, keys, rank_mode):
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    monkeypatch.setenv("PGSD_B200_SLOT", "0")   # the general (stable) path; the slot path has its own tests below
    monkeypatch.setenv("PGSD_B200_RANK_MODE", rank_mode)
    n = len(keys)
    rng = np.random.default_rng(n)
    fields = [rng.standard_normal((n, 3)).astype(np.float32), rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32),
              rng.standard_normal((n, 4)).astype(np.float32), rng.standard_normal(n), np.arange(n, dtype=np.uint32)]
    o = np.argsort(keys, kind='stable')
    for want_perm in (True, False):
        s, p, outs = _reorder_device_full(cuda, keys, fields, want_perm)
        assert (s == keys[o]).all()
        if want_perm:
            assert (p == o.astype(np.uint32)).all()
        for f, g in zip(fields, outs):
            assert g.tobytes() == f[o].tobytes()


def test_reorder_bucket_path_wide_rows(cuda, monkeypatch):
    """Rows wider than the staging buffer share (16 words) take several staging rounds."""
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    monkeypatch.setenv("PGSD_B200_SLOT", "0")
    rng = np.random.default_rng(3)
    n = 30011
    keys = rng.permutation(n).astype(np.uint32)
    fields = [rng.standard_normal((n, 16)).astype(np.float32), rng.standard_normal((n, 5)).astype(np.float32)]
    o = np.argsort(keys, kind='stable')
    s, p, outs = _reorder_device_full(cuda, keys, fields)
    assert (s == keys[o]).all() and (p == o.astype(np.uint32)).all()
    for f, g in zip(fields, outs):
        assert g.tobytes() == f[o].tobytes()


@pytest.mark.parametrize("rank_mode", ["ballot", "match"])
def test_sort_ids_rank_modes(cuda, monkeypatch, rank_mode):
    monkeypatch.setenv("PGSD_B200_RANK_MODE", rank_mode)
    rng = np.random.default_rng(8)
    keys = rng.integers(0, 2 ** 20, size=250001).astype(np.uint32)
    s, p = gpu_sort(cuda, keys)
    o = np.argsort(keys, kind='stable')
    assert (p == o.astype(np.uint32)).all() and (s == keys[o]).all()


@pytest.mark.parametrize("widths,layout", [((16, 16), "aos"), ((16, 16, 16), "aos"), ((3, 1, 1), "soa"), ((2, 7, 1), "aos"),
                                           ((1,), "aos"), ((3, 3, 1, 1, 1), "aos"), ((3, 3, 1, 1, 1), "soa")])
def test_reorder_bucket_layouts_and_row_widths(cuda, monkeypatch, widths, layout):
    """Interleaved (AoS) bucketed copy at 4/2/1 items per thread, its 40-word limit (wider rows take
    the per-field copy), and the per-field layout forced by environment."""
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    monkeypatch.setenv("PGSD_B200_SLOT", "0")
    monkeypatch.setenv("PGSD_B200_BUCKET_LAYOUT", layout)
    rng = np.random.default_rng(sum(widths))
    n = 50021
    keys = rng.integers(0, 2 ** 18, size=n).astype(np.uint32)
    fields = [rng.integers(0, 2 ** 32, size=(n, w), dtype=np.uint64).astype(np.uint32) for w in widths]
    o = np.argsort(keys, kind='stable')
    for want_perm in (True, False):
        s, p, outs = _reorder_device_full(cuda, keys, fields, want_perm)
        assert (s == keys[o]).all()
        if want_perm:
            assert (p == o.astype(np.uint32)).all()
        for f, g in zip(fields, outs):
            assert g.tobytes() == f[o].tobytes()


# ---- slot path (kernels_slot.cu): unique ids -> fine bucket scatter + per-bucket slot placement
def _launches(lib):
    st = _lib.Stats()
    lib.pgsd_b200_get_stats(st)
    return int(st.kernel_launches)


# histogram + scan + scatter + place with the key range guessed from n; a wrong guess (ids with offset or gaps)
# costs that attempt + census + the four kernels again; anything more = the general path ran
SLOT_LAUNCHES = (4, 9)


def slot_key_cases():
    rng = np.random.default_rng(41)
    yield "perm_1", np.array([0], dtype=np.uint32), None                             # nothing varies: identity
    yield "perm_1000", rng.permutation(1000).astype(np.uint32), True
    yield "perm_1024", rng.permutation(1024).astype(np.uint32), True                 # exactly one full bucket
    yield "perm_1025", rng.permutation(1025).astype(np.uint32), True
    yield "perm_70k", rng.permutation(70001).astype(np.uint32), True
    yield "perm_300k", rng.permutation(300000).astype(np.uint32), True
    yield "perm_1Mi", rng.permutation(1 << 20).astype(np.uint32), True
    yield "sorted", np.arange(123457, dtype=np.uint32), True                         # every row of a tile hits one cursor
    yield "reversed", np.arange(99999, dtype=np.uint32)[::-1].copy(), True
    yield "even_ids", (rng.permutation(200003) * 2).astype(np.uint32), True          # half-empty buckets: compaction
    yield "sparse_unique", rng.choice(1 << 22, size=150001, replace=False).astype(np.uint32), True
    yield "const_high_bits", (rng.permutation(50000) + 0xABC00000).astype(np.uint32), True
    yield "gaps_and_offset", (rng.permutation(40000) * 3 + 77777).astype(np.uint32), True
    yield "dense_with_offset_1", (rng.permutation(65536) + 1).astype(np.uint32), True   # one key just outside the guessed range
    yield "one_dup_pair", np.concatenate([rng.permutation(90000), [4242]]).astype(np.uint32), False
    yield "dups_2bytes", rng.integers(0, 40000, size=123457).astype(np.uint32), False  # bucket overflow -> flag in the scan
    yield "dups_low_density", rng.integers(0, 1 << 22, size=100000).astype(np.uint32), False  # dups found by the placement
    yield "full_32bit", rng.choice(1 << 32, size=100003, replace=False).astype(np.uint32), None  # too many buckets: not applicable


@pytest.mark.parametrize("layout", ["flat", "lines"])
@pytest.mark.parametrize("name,keys,slot", list(slot_key_cases()), ids=[k for k, _, _ in slot_key_cases()])
def test_reorder_slot_path_equals_stable_argsort(cuda, monkeypatch, name, keys, slot, layout):
    """Unique ids take the slot path (counted by its kernel launches); duplicate ids are detected on the
    device and the stable general path produces the result.  Either way: == stable argsort + gather."""
    monkeypatch.setenv("PGSD_B200_CLUSTER", "0")   # (the opt-in cluster path has its own tests: test_gpu_cluster.py)
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    monkeypatch.setenv("PGSD_B200_SLOT_LAYOUT", layout)
    n = len(keys)
    rng = np.random.default_rng(n)
    fields = [rng.standard_normal((n, 3)).astype(np.float32), rng.standard_normal((n, 3)).astype(np.float32),
              rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32),
              rng.integers(0, 3, size=n).astype(np.uint32)]
    o = np.argsort(keys, kind='stable')
    for want_perm in (True, False):
        cuda.pgsd_b200_reset_stats()
        s, p, outs = _reorder_device_full(cuda, keys, fields, want_perm)
        launches = _launches(cuda)
        assert (s == keys[o]).all()
        if want_perm:
            assert (p == o.astype(np.uint32)).all()
        for f, g in zip(fields, outs):
            assert g.tobytes() == f[o].tobytes()
        if slot is True:
            assert launches in SLOT_LAUNCHES, launches
        elif slot is False:
            assert launches > max(SLOT_LAUNCHES), launches


@pytest.mark.parametrize("bulk", ["1", "0", "flat", "flat-plain", "unit128", "unit1024"])
@pytest.mark.parametrize("tile", ["512", "1024", "2048"])
@pytest.mark.parametrize("bits", ["10", "11", "12"])
def test_reorder_slot_path_variants(cuda, monkeypatch, bits, tile, bulk):
    """Slot bits (bucket capacity 1024/2048/4096), scatter tile sizes, both layouts of the interleaved copy
    (128-byte lines of all buckets interleaved = default, buckets contiguous = flat), plain-load staging."""
    monkeypatch.setenv("PGSD_B200_CLUSTER", "0")   # (the opt-in cluster path has its own tests: test_gpu_cluster.py)
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    monkeypatch.setenv("PGSD_B200_SLOT_BITS", bits)
    monkeypatch.setenv("PGSD_B200_SLOT_TILE", tile)
    if bulk.startswith("flat"):
        monkeypatch.setenv("PGSD_B200_SLOT_LAYOUT", "flat")
    monkeypatch.setenv("PGSD_B200_SLOT_BULK", "0" if bulk in ("0", "flat-plain") else "1")
    if bulk.startswith("unit"):
        monkeypatch.setenv("PGSD_B200_SLOT_UNIT", {"unit128": "7", "unit1024": "10"}[bulk])
    rng = np.random.default_rng(int(bits) * 7 + int(tile))
    n = 150001
    keys = (rng.permutation(n) + rng.integers(0, 2)).astype(np.uint32)
    fields = [rng.integers(0, 2 ** 32, size=(n, 3), dtype=np.uint64).astype(np.uint32), rng.standard_normal(n),
              rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)]
    o = np.argsort(keys, kind='stable')
    cuda.pgsd_b200_reset_stats()
    s, p, outs = _reorder_device_full(cuda, keys, fields, True)
    assert _launches(cuda) in SLOT_LAUNCHES
    assert (s == keys[o]).all() and (p == o.astype(np.uint32)).all()
    for f, g in zip(fields, outs):
        assert g.tobytes() == f[o].tobytes()


@pytest.mark.parametrize("n", [1024, 1025, 2047, 5 * 1024, 100003, (1 << 20) + 77, 3 << 20])
def test_reorder_slot_sph_row(cuda, monkeypatch, n):
    """The 40-byte SPH row (key + position + velocity + three scalars) without the original-index column -- the call
    pgsd.hoomd makes -- at sizes around the tile boundaries: stable-argsort order bit for bit; ids with gaps and a
    duplicate (-> general path) included."""
    monkeypatch.setenv("PGSD_B200_CLUSTER", "0")
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    rng = np.random.default_rng(n)
    for case in ("dense", "gaps", "duplicate"):
        if case == "dense":
            keys = rng.permutation(n).astype(np.uint32)
        elif case == "gaps":
            keys = rng.permutation(n + n // 7 + 3)[:n].astype(np.uint32)
        else:
            keys = rng.permutation(n).astype(np.uint32)
            keys[n // 2] = keys[n // 3]
        fields = [rng.standard_normal((n, 3)).astype(np.float32), rng.standard_normal((n, 3)).astype(np.float32),
                  rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32),
                  rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)]
        o = np.argsort(keys, kind='stable')
        s, p, outs = _reorder_device_full(cuda, keys, fields, False)
        assert (s == keys[o]).all()
        for f, g in zip(fields, outs):
            assert g.tobytes() == f[o].tobytes(), case


@pytest.mark.parametrize("layout", ["flat", "lines"])
@pytest.mark.parametrize("tile", ["512", "1024", "2048"])
@pytest.mark.parametrize("per", ["1", "2", "4"])
def test_reorder_slot_scatter_rows_per_thread(cuda, monkeypatch, per, tile, layout):
    """Scatter launch shapes: 1, 2 or 4 rows per thread (PGSD_B200_SLOT_PER) for every tile size, with and without
    the original-index column, sizes that leave a ragged last tile."""
    monkeypatch.setenv("PGSD_B200_CLUSTER", "0")
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    monkeypatch.setenv("PGSD_B200_SLOT_PER", per)
    monkeypatch.setenv("PGSD_B200_SLOT_TILE", tile)
    monkeypatch.setenv("PGSD_B200_SLOT_LAYOUT", layout)
    for n in (1023, 4096, 300007):
        rng = np.random.default_rng(n + int(per) + int(tile))
        keys = rng.permutation(n).astype(np.uint32)
        fields = [rng.standard_normal((n, 3)).astype(np.float32), rng.standard_normal((n, 3)).astype(np.float32),
                  rng.standard_normal(n).astype(np.float32), rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)]
        o = np.argsort(keys, kind='stable')
        for want_perm in (True, False):
            cuda.pgsd_b200_reset_stats()
            s, p, outs = _reorder_device_full(cuda, keys, fields, want_perm)
            assert _launches(cuda) in SLOT_LAUNCHES
            assert (s == keys[o]).all() and (p is None or (p == o.astype(np.uint32)).all())
            for f, g in zip(fields, outs):
                assert g.tobytes() == f[o].tobytes()


@pytest.mark.parametrize("layout", ["flat", "lines"])
@pytest.mark.parametrize("widths", [(1,), (2, 2), (3, 3, 1, 1, 1), (4, 4, 4, 4), (5, 7), (16, 14), (16, 16)])
def test_reorder_slot_path_row_widths(cuda, monkeypatch, widths, layout):
    """Records of 2..31 words take the slot path; wider records fall back to the general path."""
    monkeypatch.setenv("PGSD_B200_CLUSTER", "0")   # (the opt-in cluster path has its own tests: test_gpu_cluster.py)
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    monkeypatch.setenv("PGSD_B200_SLOT_LAYOUT", layout)
    rng = np.random.default_rng(sum(widths))
    n = 70001
    keys = rng.permutation(n).astype(np.uint32)
    fields = [rng.integers(0, 2 ** 32, size=(n, w), dtype=np.uint64).astype(np.uint32) for w in widths]
    o = np.argsort(keys, kind='stable')
    for want_perm in (True, False):
        cuda.pgsd_b200_reset_stats()
        s, p, outs = _reorder_device_full(cuda, keys, fields, want_perm)
        took_slot = _launches(cuda) in SLOT_LAUNCHES
        assert took_slot == (1 + sum(widths) + (1 if want_perm else 0) <= 32)
        assert (s == keys[o]).all()
        if want_perm:
            assert (p == o.astype(np.uint32)).all()
        for f, g in zip(fields, outs):
            assert g.tobytes() == f[o].tobytes()


@pytest.mark.parametrize("seed", range(24))
def test_reorder_slot_path_random_geometry(cuda, monkeypatch, seed):
    """Seeded random cases: n, id range (10..27 bits, so every bucket count from 1 to 32768 and every slot width),
    density of the ids in that range, an id offset, record shape and whether the permutation is wanted.
    Always == numpy stable argsort + gather, bit for bit."""
    monkeypatch.setenv("PGSD_B200_CLUSTER", "0")   # (the opt-in cluster path has its own tests: test_gpu_cluster.py)
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    rng = np.random.default_rng(9000 + seed)
    bits = int(rng.integers(10, 28))
    n = int(min(rng.integers(1, 400000), 1 << bits))
    keys = rng.choice(1 << bits, size=n, replace=False).astype(np.uint32)
    if rng.random() < 0.3:
        keys = (keys + np.uint32(rng.integers(1, 1 << 30))).astype(np.uint32)       # ids with an offset
    if rng.random() < 0.2 and n > 10:
        keys[int(rng.integers(0, n))] = keys[int(rng.integers(0, n))]               # maybe one duplicate: general path
    widths = [int(w) for w in rng.integers(1, 6, size=int(rng.integers(1, 6)))]
    fields = [rng.integers(0, 2 ** 32, size=(n, w), dtype=np.uint64).astype(np.uint32) for w in widths]
    want_perm = bool(rng.integers(0, 2))
    o = np.argsort(keys, kind='stable')
    s, p, outs = _reorder_device_full(cuda, keys, fields, want_perm)
    assert (s == keys[o]).all()
    if want_perm:
        assert (p == o.astype(np.uint32)).all()
    for f, g in zip(fields, outs):
        assert g.tobytes() == f[o].tobytes()


def test_reorder_slot_path_unaligned_inputs(cuda, monkeypatch):
    """Field arrays that start 4 bytes off a 16-byte boundary are staged with plain loads."""
    monkeypatch.setenv("PGSD_B200_CLUSTER", "0")   # (the opt-in cluster path has its own tests: test_gpu_cluster.py)
    monkeypatch.setenv("PGSD_B200_BUCKET_MIN_ROWS", "0")
    rng = np.random.default_rng(77)
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
    assert _launches(cuda) in SLOT_LAUNCHES
    o = np.argsort(keys, kind='stable')
    assert (ds.to_numpy() == keys[o]).all()
    assert opos.to_numpy()[1:].tobytes() == pos[o].tobytes()
    assert odens.to_numpy()[1:].tobytes() == dens[o].tobytes()


@pytest.mark.parametrize("M", [1, 2, 3, 4])
@pytest.mark.parametrize("N", [1, 5, 4096, 100003])
@pytest.mark.parametrize("dt", [np.float32, np.int32])
def test_pack_scalar4_records(cuda, N, M, dt):
    """HOOMD Scalar4 / int4 arrays: the leading M components of 16-byte records, read in place
    (one device buffer, columns base+4j with stride 4) -> (N, M); takes K1's record path."""
    rng = np.random.default_rng(N + M)
    aos = (rng.standard_normal((N, 4)) * 1000).astype(dt)
    d = DeviceArray.from_numpy(aos)
    cols = (_lib.Column * M)(*[_lib.Column(d.ptr + 4 * j, 4) for j in range(M)])
    out = DeviceArray((N, M), dt)
    t = _NP_TO_PGSD[np.dtype(dt)]
    _lib.check(cuda.pgsd_b200_pack_soa(out.ptr, t, N, M, t, cols, None), "pack_soa")
    _lib.check(cuda.pgsd_b200_synchronize(), "sync")
    assert out.to_numpy().tobytes() == np.ascontiguousarray(aos[:, :M]).tobytes()
