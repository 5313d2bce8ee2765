"""CPU: the reference arm of bench.py (`--impl reference`: the burn-ndarray restatement on the host cores) runs without a
GPU, prints ONE JSON line with the contract's keys, and carries the same `config` as the GPU arm — the driver divides one
arm's `e2e` by the other's, so they must describe the same workload."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_reference_arm_prints_the_contract_line_with_the_gpu_arms_config():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["value"] > 0 and line["unit"] == "GB/s" and line["higher_is_better"] is True
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.bench_config(1, bench.step_bytes(bench.N_ROWS, bench.N_COLS))
    assert "configs[1]" in line["config"]["workload"]
