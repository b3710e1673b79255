"""GPU parity tests of the device-resident write path (K1 -> arena -> K3 staging) and the
ID-reordered read path (K4 + K5) through the drop-in API, against reference-made goldens."""
import hashlib
import json
import os

import numpy as np
import pytest

import opscript
from golden.make_golden import GOLDEN_FRAMES, GOLDEN_N, hoomd_script, seed_nprocs
from oracle import reader_oracle, reorder_oracle
from pgsd_sph_b200 import fl, hoomd, synth
from pgsd_sph_b200.devmem import DeviceArray
from randscript import random_script

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def cuda(lib):
    assert lib.pgsd_b200_cuda_available() == 1, "no CUDA device: the device path has no CPU fallback"


@pytest.mark.parametrize("soa", [False, True], ids=["packed", "soa"])
@pytest.mark.parametrize("P", [1, 2, 8])
def test_device_write_matches_golden(golden, tmp_path, P, soa):
    """Chunk bytes come from device memory; P ranks share the GPU (shm communicator)."""
    gsd, prefix = opscript.run_replay(hoomd_script(), str(tmp_path), f"d{P}", P, device=True, soa=soa, timeout=300)
    assert opscript.read_bytes(gsd) == opscript.read_bytes(os.path.join(golden, f"hoomd_p{P}.gsd"))
    assert opscript.read_bytes(prefix + ".log") == opscript.read_bytes(os.path.join(golden, f"hoomd_p{P}.log"))


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 5, 7, 10, 12, 17, 22, 31, 40])
def test_device_random_scripts_match_golden_hashes(golden, tmp_path, seed):
    sums = json.load(open(os.path.join(golden, "script_sha256.json")))[str(seed)]
    P = seed_nprocs(seed)
    gsd, prefix = opscript.run_replay(random_script(seed, P, lookups=(seed % 3 != 0)), str(tmp_path), f"s{seed}", P,
                                      device=True, soa=(seed % 2 == 0), timeout=300)
    assert hashlib.sha256(opscript.read_bytes(gsd)).hexdigest() == sums["gsd"]
    assert hashlib.sha256(opscript.read_bytes(prefix + ".log")).hexdigest() == sums["log"]


def _write_frames_python(path, device, soa, f64_sources=False):
    """The golden P=1 file through the Python drop-in API with device-resident fields."""
    with fl.open(path, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
        for i in range(GOLDEN_FRAMES):
            fr = synth.make_frame(GOLDEN_N, i)
            for k, a in synth.frame_scalars(GOLDEN_N, i):
                f.write_chunk(k, a, write_all=False)
            for k, a in fr.items():
                if not device:
                    f.write_chunk(k, a)
                elif soa and a.ndim == 2:
                    src = a.astype(np.float64) if f64_sources else a
                    cols = [DeviceArray.from_numpy(np.ascontiguousarray(src[:, j])) for j in range(a.shape[1])]
                    f.write_chunk_soa(k, cols, dtype=a.dtype)
                elif soa and f64_sources and a.dtype == np.float32:
                    f.write_chunk_soa(k, [DeviceArray.from_numpy(a.astype(np.float64))], dtype=np.float32)
                else:
                    f.write_chunk(k, DeviceArray.from_numpy(a))
            f.write_chunk("log/value/kinetic_energy", np.array([0.5 * i + 1.25], dtype=np.float32), write_all=False)
            f.write_chunk("log/value/potential_energy", np.array([-3.0 * i], dtype=np.float32), write_all=False)
            f.end_frame()


@pytest.mark.parametrize("mode", ["host", "device", "device_soa", "device_soa_f64"])
def test_python_api_write_matches_golden(golden, tmp_path, mode):
    p = str(tmp_path / "w.gsd")
    _write_frames_python(p, device=mode != "host", soa="soa" in mode, f64_sources=mode.endswith("f64"))
    assert opscript.read_bytes(p) == opscript.read_bytes(os.path.join(golden, "hoomd_p1.gsd"))


def test_strided_device_array_is_packed_on_device(tmp_path):
    """A non-contiguous CUDA array goes through K1 (device-side ascontiguousarray, fl.pyx:571)."""
    n = 3000
    rng = np.random.default_rng(1)
    aos = rng.standard_normal((n, 4)).astype(np.float32)

    class View:  # (n, 3) view of the (n, 4) device buffer
        def __init__(self, d):
            self.d = d
            self.__cuda_array_interface__ = {"shape": (n, 3), "typestr": "<f4", "data": (d.ptr, False), "version": 3,
                                             "strides": (16, 4)}
    d = DeviceArray.from_numpy(aos)
    p1, p2 = str(tmp_path / "a.gsd"), str(tmp_path / "b.gsd")
    for p, data in ((p1, View(d)), (p2, aos[:, :3])):
        with fl.open(p, 'w', 'a', 's', [1, 0]) as f:
            f.write_chunk("pos", data)
            f.end_frame()
    assert opscript.read_bytes(p1) == opscript.read_bytes(p2)


def test_many_frames_async_staging_overlap(tmp_path):
    """Frames are queued faster than they drain: arena recycling + pinned ring keep the bytes right."""
    n, frames = 200000, 12
    p = str(tmp_path / "m.gsd")
    expect = []
    with fl.open(p, 'w', 'a', 'hoomd', [1, 4]) as f:
        for i in range(frames):
            fr = synth.make_frame(n, i, cheap=True)
            expect.append(fr)
            for k, a in fr.items():
                if a.ndim == 2:
                    f.write_chunk_soa(k, [DeviceArray.from_numpy(np.ascontiguousarray(a[:, j])) for j in range(3)])
                else:
                    f.write_chunk(k, DeviceArray.from_numpy(a))
            f.end_frame()
    o = reader_oracle.OracleFile(p)
    assert o.nframes == frames
    for i in (0, 5, frames - 1):
        for k, a in expect[i].items():
            assert o.read_chunk(i, k).tobytes() == a.tobytes(), (i, k)


# ---------------------------------------------------------------------------- read + reorder
@pytest.mark.parametrize("device", [False, True], ids=["host", "device"])
def test_hoomd_reorder_matches_reference_python_golden(golden, device):
    """HOOMDTrajectory(reorder='id') == the reference's pypgsd + hoomd + argsort (golden npz)."""
    ref = np.load(os.path.join(golden, "reorder_p2.npz"))
    with hoomd.open(os.path.join(golden, "hoomd_p2.gsd"), 'r', reorder='id', device=device) as t:
        assert len(t) == GOLDEN_FRAMES
        for i in range(len(t)):
            fr = t[i]
            assert fr.particles.N == ref[f"f{i}/N"][0]
            assert fr.configuration.step == ref[f"f{i}/step"][0]
            ids = fr.log['particles/id']
            ids = ids.to_numpy() if device else ids
            assert (ids == ref[f"f{i}/id"]).all()
            for name in reader_oracle.PARTICLE_DEFAULTS:
                a = getattr(fr.particles, name)
                a = a.to_numpy() if hasattr(a, "to_numpy") else a
                b = ref[f"f{i}/{name}"]
                assert a.dtype == b.dtype and a.shape == b.shape, name
                assert a.tobytes() == b.tobytes(), name
            assert fr.log['value/kinetic_energy'][0] == ref[f"f{i}/log/value/kinetic_energy"][0]


def test_hoomd_unsorted_read_matches_oracle(golden):
    f = reader_oracle.OracleFile(os.path.join(golden, "hoomd_p8.gsd"))
    with hoomd.open(os.path.join(golden, "hoomd_p8.gsd"), 'r') as t:
        for i in range(len(t)):
            dec = reader_oracle.decode_particles(f, i)
            fr = t[i]
            for name in reader_oracle.PARTICLE_DEFAULTS:
                assert getattr(fr.particles, name).tobytes() == dec[name].tobytes(), name


def test_roundtrip_1M_device_write_then_reordered_read(tmp_path):
    """BASELINE config-2 frame size: device write -> file -> decode -> K4/K5 == oracle reorder."""
    n = 1 << 20
    fr = synth.make_frame(n, 3, cheap=True)
    p = str(tmp_path / "r.gsd")
    with fl.open(p, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
        for k, a in synth.frame_scalars(n, 3):
            f.write_chunk(k, a, write_all=False)
        for k, a in fr.items():
            if a.ndim == 2:
                f.write_chunk_soa(k, [DeviceArray.from_numpy(np.ascontiguousarray(a[:, j])) for j in range(3)])
            else:
                f.write_chunk(k, DeviceArray.from_numpy(a))
        f.end_frame()
    o = np.argsort(fr["log/particles/id"], kind='stable')
    with hoomd.open(p, 'r', reorder='id') as t:
        got = t[0]
        assert (got.log['particles/id'] == np.arange(n, dtype=np.uint32)).all()
        assert got.particles.position.tobytes() == fr["particles/position"][o].tobytes()
        assert got.particles.velocity.tobytes() == fr["particles/velocity"][o].tobytes()
        assert got.particles.density.tobytes() == fr["particles/density"][o].tobytes()
        assert got.particles.typeid.tobytes() == fr["particles/typeid"][o].tobytes()


def test_vtu_columns_on_device_match_numpy(golden):
    """K6 = K1's strided cast path: (N,3) float32 on the device -> three contiguous float64 columns."""
    from pgsd_sph_b200 import vtu
    with hoomd.open(os.path.join(golden, "hoomd_p2.gsd"), 'r', reorder='id', device=True) as t:
        fr = t[2]
        x, y, z, pd = vtu.point_arrays(fr)
        pos = fr.particles.position.to_numpy()
        vel = fr.particles.velocity.to_numpy()
        for j, c in enumerate((x, y, z)):
            assert c.to_numpy().tobytes() == np.ascontiguousarray(pos[:, j], dtype=np.float64).tobytes()
        assert pd['velocity'][1].to_numpy().tobytes() == np.ascontiguousarray(vel[:, 1], dtype=np.float64).tobytes()
        assert pd['density'].to_numpy().tobytes() == fr.particles.density.to_numpy().astype(np.float64).tobytes()


def test_vtu_file_from_device_frame_matches_oracle(golden, tmp_path):
    """pgsd2vtu on a device-resident reordered frame: K1 column split / f64 cast / xyz interleave, D2H into the
    page-locked file image, threaded file stage == the sequential restatement of pyevtk's writer fed with the numpy
    arrays of the manual's listing (container parity with pyevtk itself unpinned, see oracle/vtu_oracle.py)."""
    from oracle import vtu_oracle
    from pgsd_sph_b200 import vtu
    path = os.path.join(golden, "hoomd_p2.gsd")
    with hoomd.open(path, 'r', reorder='id', device=True) as td, hoomd.open(path, 'r', reorder='id') as th:
        for i in (2, 0, 2):   # the second visit of frame 2 reuses the cached image
            x, y, z, pd = vtu.point_arrays(td[i])
            name = vtu.write_vtu(str(tmp_path / f"dev_{i}"), x, y, z, pd)
            hx, hy, hz, hpd = vtu.point_arrays(th[i])
            assert open(name, 'rb').read() == vtu_oracle.points_to_vtk_bytes(hx, hy, hz, hpd)


@pytest.mark.parametrize("n", [1, 3, 100003, (1 << 21) + 7])
def test_vtu_device_arrays_sizes(tmp_path, n):
    """Device arrays straight into write_vtu (pointsToVTK's signature): odd sizes, > 16 MiB images (several writer
    threads), float32 coordinates with float64 data."""
    from oracle import vtu_oracle
    from pgsd_sph_b200 import vtu
    rng = np.random.default_rng(n)
    x, y, z = (rng.standard_normal(n).astype(np.float32) for _ in range(3))
    pd = {'rho': rng.random(n), 'v': tuple(rng.standard_normal(n) for _ in range(3)), 'k': rng.integers(0, 9, n).astype(np.uint8)}
    dx, dy, dz = (DeviceArray.from_numpy(a) for a in (x, y, z))
    dpd = {'rho': DeviceArray.from_numpy(pd['rho']), 'v': tuple(DeviceArray.from_numpy(a) for a in pd['v']),
           'k': DeviceArray.from_numpy(pd['k'])}
    name = vtu.write_vtu(str(tmp_path / "d"), dx, dy, dz, dpd)
    assert open(name, 'rb').read() == vtu_oracle.points_to_vtk_bytes(x, y, z, pd)


def test_torch_cuda_tensors_cai_and_dlpack(tmp_path):
    """`data` may be any CUDA array: torch tensors through __cuda_array_interface__, a DLPack-only
    wrapper through __dlpack__, a strided torch view through the device-side pack (K1)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("torch sees no CUDA device")
    n = 4097
    g = torch.Generator(device="cuda").manual_seed(3)
    pos4 = torch.rand((n, 4), device="cuda", generator=g, dtype=torch.float32)      # HOOMD Scalar4-like
    dens = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
    ids = torch.randperm(n, device="cuda", generator=g).to(torch.int32)

    class DLPackOnly:  # hides __cuda_array_interface__
        def __init__(self, t):
            self.t = t

        def __dlpack__(self, stream=None):
            return self.t.__dlpack__()

        def __dlpack_device__(self):
            return self.t.__dlpack_device__()

    torch.cuda.synchronize()
    p = str(tmp_path / "torch.gsd")
    with fl.open(p, 'w', 'app', 'hoomd', [1, 4]) as f:
        f.write_chunk("particles/position", pos4[:, :3])                      # strided (n,3) view of (n,4)
        f.write_chunk("particles/density", DLPackOnly(dens))                  # DLPack intake
        f.write_chunk("log/particles/id", ids)                                # CAI intake, contiguous
        f.write_chunk_soa("particles/velocity", [pos4[:, 3], pos4[:, 0], pos4[:, 1]], dtype=np.float64)  # cast f32->f64
        f.end_frame()
    with fl.open(p, 'r') as f:
        h = pos4.cpu().numpy()
        assert f.read_chunk(0, "particles/position").tobytes() == np.ascontiguousarray(h[:, :3]).tobytes()
        assert f.read_chunk(0, "particles/density").tobytes() == dens.cpu().numpy().tobytes()
        assert f.read_chunk(0, "log/particles/id").tobytes() == ids.cpu().numpy().tobytes()
        want = np.ascontiguousarray(np.stack([h[:, 3], h[:, 0], h[:, 1]], 1), dtype=np.float64)
        assert f.read_chunk(0, "particles/velocity").tobytes() == want.tobytes()
        # and back into torch without a copy: DeviceArray exposes __cuda_array_interface__
        d = f.read_chunk(0, "particles/position", device=True)
        t = torch.as_tensor(d, device="cuda")
        assert torch.equal(t, pos4[:, :3])
