#!/bin/bash
# call I: smoke, the whole GPU suite, the full default bench
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 400 python -m pytest tests -q -x --tb=short -m gpu -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/gpu_tests.log
tail -c 800 gpurun_out/gpu_tests.log
timeout 420 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
tail -c 300 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "e2e", "graph_index", "clocks")}, indent=None)[:6000])
except Exception as e:
    print("bench parse failed", e)
PY
