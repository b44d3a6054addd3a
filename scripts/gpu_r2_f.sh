#!/bin/bash
# round 2, call F (2 GPUs): sharded parity against the oracle + the N=2 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_scan_tc_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider > gpurun_out/f_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/f_tests.log; tail -c 2500 gpurun_out/f_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/f_bench2.json 2> gpurun_out/f_bench2.err
echo "bench exit $?"; tail -c 1500 gpurun_out/f_bench2.err; python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/f_bench2.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"], d["config"]["parity_check"], d["roofline"].get("encoder",{}).get("recall_e2e_ms_per_query"), d["roofline"]["scan_ms_per_step"], d["roofline"]["batch1"]["ms_per_query"])
except Exception as e: print("parse failed", e)
PY
