"""The reference arm of `bench.py` (`--impl reference`: the oracle port on the host cores) prints the contract's JSON
line without a GPU; under torchrun rank 0 alone runs it.  (The CUDA arm needs a B200 and is run by the driver.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _last_json(out):
    lines = [l for l in out.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out
    return json.loads(lines[0])


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1", "--envs", "256"],
                         cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    d = _last_json(res.stdout)
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["higher_is_better"] is True
    assert d["unit"] == "env-steps/s" and d["dtype"] == "f32" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and abs(d["value"] - 256 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["baseline_config"] == 2 and "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_under_torchrun_runs_on_rank_0_only():
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", "bench.py", "--impl", "reference",
                          "--gpus", "2", "--steps", "1", "--warmup", "1", "--envs", "128", "--config", "3"],
                         cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    d = _last_json(res.stdout)            # exactly one line: the other rank exits without work
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["config"]["baseline_config"] == 3
