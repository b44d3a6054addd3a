#!/bin/bash
# call C: graph tests after the staged-kernel change, then only the graph leg of the bench on a small exact corpus
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_vamana_gpu.py tests/test_scan_metrics_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider -s > gpurun_out/graph_tests.log 2>&1
echo "graph tests exit $?" >> gpurun_out/graph_tests.log
tail -c 1500 gpurun_out/graph_tests.log
SG_BENCH_ROWS=1000000 SG_BENCH_ENCODER=0 SG_BENCH_CPU_DETAIL=0 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err
echo "bench exit $?"
tail -c 400 gpurun_out/bench_graph.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_graph.json").read().strip().splitlines()[-1])
    print(json.dumps(d["graph_index"], indent=1)[:4000])
except Exception as e:
    print("bench parse failed", e)
PY
