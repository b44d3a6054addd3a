#!/bin/bash
# round 2, call w (2 GPUs): sharded tests + the N=2 scaling point
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -5
n=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/w_bench$n.json 2> gpurun_out/w_bench$n.err
echo "bench $n exit $?"; tail -c 300 gpurun_out/w_bench$n.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/w_bench$n.json") if l.startswith("{")][-1])
    r=d["roofline"]
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], d["config"]["parity_check"]["mismatches"], "scan", r["scan_ms_per_step"], "b1", r["batch1"]["ms_per_query"], "k120", r["k120"]["ms_per_step"], "cfg5", r.get("encoder",{}).get("recall_e2e_ms_per_query"), d["config"]["parallelism"])
except Exception as e: print("parse failed", e)
PY
