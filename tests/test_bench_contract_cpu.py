"""The reference arm of bench.py runs without a GPU (it times the oracle port of the reference's CPU algorithm on the host
cores): its JSON line must carry the contract's keys with the same metric, unit, direction and ``config`` object the GPU arm
prints, so that the driver can divide one by the other."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_arm(*flags):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *flags], capture_output=True,
                         text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                            # ONE JSON line
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_arm("--gpus", "1", "--steps", "2", "--warmup", "1", "--cpu-seconds", "1")
    assert d["impl"] == "reference" and d["metric"] == "cell_steps_per_sec" and d["unit"] == "cell-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["ms_per_step"] > 0 and d["value"] > 0
    assert d["dtype"] == "f64"                                        # the reference computes in float64
    cfg = d["config"]
    assert cfg["workload"].startswith("c5:") and cfg["genes"] == 3000 and cfg["cells_per_gpu"] == 1000000
    assert cfg["tree_parents"] == [-1, 0, 0, 1, 1, 2, 2] and cfg["alpha"] == 0.4 and "model" not in cfg
    cpu = d["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["unit"] == d["unit"] and cpu["value"] == d["value"]
    assert "findCellNextState" in cpu["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "unavailable" not in d


def test_reference_arm_under_torchrun_prints_on_rank_zero_only():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--cpu-seconds", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]      # other ranks: no work
