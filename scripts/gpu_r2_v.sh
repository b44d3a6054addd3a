#!/bin/bash
# round 2, call v (8 GPUs): the scaling points N = 8 and 4
mkdir -p gpurun_out
for n in 8 4; do
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/v_bench$n.json 2> gpurun_out/v_bench$n.err
echo "bench $n exit $?"; tail -c 200 gpurun_out/v_bench$n.err | tail -2
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/v_bench$n.json") if l.startswith("{")][-1])
    r=d["roofline"]
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], d["config"]["parity_check"]["mismatches"], "scan", r["scan_ms_per_step"], "b1", r["batch1"]["ms_per_query"], "k120", r["k120"]["ms_per_step"], "cfg5", r.get("encoder",{}).get("recall_e2e_ms_per_query"))
except Exception as e: print("parse failed", e)
PY
done
