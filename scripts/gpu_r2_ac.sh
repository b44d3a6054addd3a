#!/bin/bash
# round 2, call ac: validation after the last changes (fused QKV+attention, finalize rounds, page-locked DMA)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/ac_tests.log; tail -6 gpurun_out/ac_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
NS=10000000,1250000 CASES=256x120,64x120,256x10 timeout 300 python scripts/probe_step.py 2>&1 | tail -6
timeout 900 python bench.py --steps 20 --warmup 5 --legs > gpurun_out/ac_bench1.json 2> gpurun_out/ac_bench1.err; echo "bench exit $?"; tail -c 300 gpurun_out/ac_bench1.err
cp profiles/r02_legs.json gpurun_out/ac_legs.json 2>/dev/null
timeout 600 compute-sanitizer --tool memcheck --launch-timeout 120 --print-limit 20 python -m pytest tests/test_encoder_gpu.py -q -x --tb=line -m gpu -p no:cacheprovider -k "matches_torch or config5" > gpurun_out/ac_enc_memcheck.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/ac_enc_memcheck.log | tail -3
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/ac_bench1.json") if l.startswith("{")][-1])
r=d["roofline"]
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], d["config"]["parity_check"]["mismatches"], "scan", r["scan_ms_per_step"], "frac", r["frac"], "b1", r["batch1"]["ms_per_query"], "k120", r["k120"]["ms_per_step"], "enc", r["encoder"]["encode_ms_batch1"], r["encoder"]["encode_ms_batch256"], "cfg5", r["encoder"]["recall_e2e_ms_per_query"], d["clocks"])
PY
