"""bench.py's JSON contract, checked on CPU through the reference arm (the only arm that runs without a GPU)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_has_the_contract_keys(tmp_path):
    env = dict(os.environ, PGSD_BENCH_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--particles", "200000", "--read-particles", "100000", "--quick"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    if "unavailable" in line:          # oracle/_ref not built on this host
        assert line["impl"] == "reference"
        return
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "frame_write_GBps" and line["unit"] == "GB/s"
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert "workload" in line["config"] and line["vs_baseline"] is None and line["value"] > 0
    rr = line["read_reorder"]
    assert rr["metric"] == "id_reordered_read_Mparticles_per_s" and rr["value"] > 0 and rr["cpu_baseline"]["kind"] in ("reference", "port")
    assert not os.listdir(str(tmp_path)) or all(not os.listdir(os.path.join(str(tmp_path), d)) for d in os.listdir(str(tmp_path)))


def test_other_ranks_of_the_reference_arm_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
