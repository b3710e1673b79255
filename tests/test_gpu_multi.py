"""Multi-GPU tests (skipped on a box with one GPU): the NCCL transport on real devices.  One process per GPU,
launched with torch.distributed.run; each worker checks its own share against the oracle.
  (i)  a frame written by N ranks under the NCCL communicator == the file the unmodified reference writes at N ranks
  (ii) pgsd_b200_reorder_distributed on N real GPUs == oracle/reorder_oracle.reorder_distributed, rank by rank"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


NG = _ngpus()


def _launch(n, what, tmp_path, extra=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + n), os.path.join(REPO, "tests", "multi_gpu_worker.py"), what, str(tmp_path)] + list(extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=REPO)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert f"WORKER-OK {what} ranks={n}" in r.stdout, r.stdout[-2000:]


@pytest.mark.skipif(NG < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("n", [k for k in (2, 4, 8) if k <= max(NG, 2)])
def test_nccl_transport_write_matches_reference_file(tmp_path, n):
    _launch(n, "write", tmp_path)


@pytest.mark.skipif(NG < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("n", [k for k in (2, 4, 8) if k <= max(NG, 2)])
def test_distributed_reorder_on_real_gpus_matches_oracle_shares(tmp_path, n):
    _launch(n, "reorder", tmp_path)
