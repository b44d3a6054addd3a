#!/bin/bash
# call B: the whole GPU suite (regression + new), then the bench with the graph-index leg
mkdir -p gpurun_out
timeout 540 python -m pytest tests -q -x --tb=short -m gpu -p no:cacheprovider -s > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/gpu_tests.log
tail -c 2500 gpurun_out/gpu_tests.log
timeout 480 python bench.py --steps 50 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
tail -c 600 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "e2e", "graph_index", "clocks")}, indent=1)[:3500])
except Exception as e:
    print("bench parse failed", e)
PY
