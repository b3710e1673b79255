"""Distributed reorder (pgsd_b200_reorder_distributed, SURVEY.md section 8e: one frame partitioned over the
ranks): P processes on cuda:0 with the shared-memory communicator; records travel between the processes through
CUDA IPC mappings exactly as they do between GPUs.  Oracle: numpy stable argsort of the whole frame."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import dist_reorder_worker as W
from oracle import reorder_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nprocs,mode", [(1, "partition"), (2, "partition"), (3, "partition"), (4, "partition"), (8, "partition"),
                                         (2, "partition-host"), (8, "partition-host"), (2, "fused"), (3, "fused"), (8, "fused")])
def test_reorder_distributed_equals_stable_argsort(lib, nprocs, mode):
    """mode: how records reach their owners -- "partition" (runs appended to the owner's inbox, then the local
    two-pass reorder), "partition-host" (the same with PGSD_B200_DIST_HOST=1) or "fused" (every record stored directly
    at its place in the owner's bucketed copy).  All ranks of this test share cuda:0, so the library keeps the
    exchange of counts and completion flags on the host communicator in every mode (kernels of different processes
    on one GPU must not wait for one another); the device-driven exchange runs in tests/test_gpu_multi.py, one GPU
    per rank."""
    assert lib.pgsd_b200_cuda_available() == 1, "no CUDA device: the device path has no CPU fallback"
    with tempfile.TemporaryDirectory() as d:
        # a frame file for the end-to-end leg (written here by one rank: whole chunks, like any reference file)
        from pgsd_sph_b200 import fl, synth
        FILE_N = 123457
        frame = synth.make_frame(FILE_N, 3)
        with fl.open(os.path.join(d, "frame.gsd"), 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
            for k, a in synth.frame_scalars(FILE_N, 0):
                f.write_chunk(k, a, write_all=False)
            for k, a in frame.items():
                f.write_chunk(k, a)
            f.end_frame()
        seg = f"/pgsd_dist_{os.getpid()}_{nprocs}_{mode}"
        env = dict(os.environ, PGSD_B200_DIST_MODE=mode.split("-")[0], PGSD_B200_DIST_HOST="1" if mode.endswith("-host") else "0")
        procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "dist_reorder_worker.py"), str(r), str(nprocs), seg, d],
                                  env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(nprocs)]
        outs = []
        for p in procs:
            try:
                o, _ = p.communicate(timeout=300)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
            outs.append(o)
        for r, (p, o) in enumerate(zip(procs, outs)):
            assert p.returncode == 0, f"rank {r}:\n{o}"
        for ci, (name, n, gen) in enumerate(W.cases()):
            rng = np.random.default_rng(1000 + ci)
            ids = gen(rng, n)
            pos = rng.standard_normal((n, 3)).astype(np.float32)
            tag = (ids ^ np.uint32(0x9e3779b9)).astype(np.uint32)
            dens = rng.standard_normal(n)
            res = [np.load(os.path.join(d, f"rank{r}_{name}.npz")) for r in range(nprocs)]
            rcs = {int(x["rc"]) for x in res}
            assert len(rcs) == 1, (name, rcs)          # a collective decision
            if name in ("one_dup", "dup_in_partial_bucket", "out_of_range"):
                assert rcs == {1}, (name, rcs)
                continue
            assert rcs == {0}, (name, rcs)
            shares = reorder_oracle.reorder_distributed(ids, {"pos": pos, "tag": tag, "dens": dens}, nprocs)
            for r, (x, (first, sid, exp)) in enumerate(zip(res, shares)):     # rank by rank against the oracle's shares
                assert int(x["id_first"]) == first and int(x["n_out"]) == len(sid), (name, r)
                assert x["ids"].tobytes() == sid.tobytes(), (name, r)
                for k in ("pos", "tag", "dens"):
                    assert x[k].tobytes() == exp[k].tobytes(), (name, r, k)
            o = np.argsort(ids, kind='stable')
            got_ids = np.concatenate([x["ids"] for x in res])
            assert got_ids.tobytes() == ids[o].tobytes(), name
            assert np.concatenate([x["pos"] for x in res]).tobytes() == pos[o].tobytes(), name
            assert np.concatenate([x["tag"] for x in res]).tobytes() == tag[o].tobytes(), name
            assert np.concatenate([x["dens"] for x in res]).tobytes() == dens[o].tobytes(), name
            # ownership: rank r holds exactly the ids of [id_first, id_first of the next rank)
            for r, x in enumerate(res):
                if x["n_out"] > 0:
                    assert x["ids"][0] >= x["id_first"]
                    if r + 1 < nprocs:
                        assert x["ids"][-1] < res[r + 1]["id_first"]
        # Python wrapper
        rng = np.random.default_rng(77)
        ids = rng.permutation(W.PYWRAP_N).astype(np.uint32)
        pos = rng.standard_normal((W.PYWRAP_N, 3)).astype(np.float32)
        res = [np.load(os.path.join(d, f"rank{r}_pywrap.npz")) for r in range(nprocs)]
        o = np.argsort(ids, kind='stable')
        assert np.concatenate([x["ids"] for x in res]).tobytes() == ids[o].tobytes()
        assert np.concatenate([x["pos"] for x in res]).tobytes() == pos[o].tobytes()
        assert all(len(x["ids"]) == 0 or x["ids"][0] == x["id_first"] for x in res)
        # file -> row-sliced device reads on every rank -> distributed reorder
        ids = frame['log/particles/id'].reshape(-1)
        o = np.argsort(ids, kind='stable')
        res = [np.load(os.path.join(d, f"rank{r}_file.npz")) for r in range(nprocs)]
        assert np.concatenate([x["ids"] for x in res]).tobytes() == ids[o].tobytes()
        for name in ("position", "velocity", "typeid", "density", "pressure"):
            want = frame['particles/' + name]
            got = np.concatenate([x[name] for x in res])
            assert got.tobytes() == want[o].tobytes(), name
