#!/bin/bash
# round 2, call G: encoder after the issue-loop fix + graph replay; finalize; PQ/SPANN legs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py tests/test_scan_tc_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider > gpurun_out/g_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/g_tests.log; tail -c 1500 gpurun_out/g_tests.log
timeout 200 python scripts/probe_enc.py 2>&1 | tail -4
SG_ENC_GRAPH=0 timeout 200 python scripts/probe_enc.py 2>&1 | tail -4
CASES=256x10,256x120 NS=10000000,1250000 REPS=100 timeout 300 python scripts/probe_step.py 2>&1 | tail -5
CASES=256x10 NS=1250000 REPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g_launches_1250k.csv python scripts/probe_step.py > /dev/null 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --legs > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench exit $?"; tail -c 400 gpurun_out/g_bench.err; cp profiles/r02_legs.json gpurun_out/g_legs.json 2>/dev/null
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/g_bench.json") if l.startswith("{")][-1])
    r=d["roofline"]
    print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["config"]["parity_check"]["mismatches"], r["frac"], r["hbm"]["frac"], r["scan_ms_per_step"], r["batch1"]["ms_per_query"], r["k120"], r.get("encoder"), r.get("other_configs"), d.get("cpu_baseline"))
    L=json.load(open("gpurun_out/g_legs.json"))
    print(json.dumps(L.get("product_quantization"))[:1500]); print(json.dumps(L.get("spann"))[:1200])
except Exception as e: print("parse failed", e)
PY
