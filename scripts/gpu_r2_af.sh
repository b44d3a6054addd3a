#!/bin/bash
# round 2, call af: finalize with batched loads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scan_tc_gpu.py tests/test_full_size_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -4
NS=10000000,1250000 CASES=256x120,256x10 timeout 300 python scripts/probe_step.py 2>&1 | tail -4
NS=1250000 CASES=256x120,256x10 REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/af_k120.csv python scripts/probe_step.py > /dev/null 2>&1
grep finalize gpurun_out/af_k120.csv | tail -8 | awk -F'","' '{print $NF}'
