#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_vamana_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider -s 2>&1 | tail -6
SG_BENCH_ROWS=1000000 SG_BENCH_ENCODER=0 SG_BENCH_CPU_DETAIL=0 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err
echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_graph.json").read().strip().splitlines()[-1])
    print(json.dumps(d["graph_index"])[:5000])
except Exception as e:
    print("bench parse failed", e)
PY
