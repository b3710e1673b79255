"""N > 1 host-side path on CPU: two processes, torch.distributed gloo backend, the library's 'host'
communicator (all-gather callback).  The file must equal the golden written by the unmodified
reference at 2 MPI ranks."""
import os
import socket
import subprocess
import sys

import pytest

import opscript

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode", ["counts", "auto"])
def test_two_rank_gloo_write_matches_reference_golden(golden, tmp_path, mode):
    path = str(tmp_path / "gloo.gsd")
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "gloo_worker.py"), path, mode], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert opscript.read_bytes(path) == opscript.read_bytes(os.path.join(golden, "hoomd_p2.gsd"))
