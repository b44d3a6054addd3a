"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the keys the driver reads, and the
optional child legs fail loudly (an error object, not a crash) when no device is present."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None, timeout=300):
    e = dict(os.environ, **(env or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                         timeout=timeout)
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert out.returncode == 0 and len(lines) == 1, out.stdout[-500:] + out.stderr[-500:]
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = run(["--impl", "reference", "--steps", "1", "--warmup", "1"], env={"SG_BENCH_CPU_ROWS": "20000"})
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "workload" in d["config"] and d["value"] > 0


def test_child_legs_report_errors_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    for leg in ("pq", "spann"):
        d = run(["--leg", leg], env={"SG_BENCH_SPANN_ROWS": "50000", "SG_BENCH_PQ_ROWS": "50000"})
        assert "error" in d and "no CUDA device" in d["error"], d
