#!/bin/bash
# round 2, call al: final check of the committed tree -- full GPU suite, smoke, the driver's bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/al_tests.log; tail -5 gpurun_out/al_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/al_bench1.json 2> gpurun_out/al_bench1.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/al_bench1.json") if l.startswith("{")][-1])
r=d["roofline"]
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], d["config"]["parity_check"]["mismatches"], "scan", r["scan_ms_per_step"], "frac", r["frac"], "b1", r["batch1"]["ms_per_query"], "k120", r["k120"]["ms_per_step"], "enc", r["encoder"]["encode_ms_batch1"], r["encoder"]["encode_ms_batch256"], "cfg5", r["encoder"]["recall_e2e_ms_per_query"], d["clocks"], "launches", d["gpu_launches"])
PY
