#!/bin/bash
# round 2, call t: two-group seed computation (one L2 round trip), cheaper slow path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scan_tc_gpu.py tests/test_scan_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -6
NS=1250000 CASES=256x10 REPS=2 SG_TC_TRACE=1 timeout 300 python scripts/probe_step.py 2>&1 | grep -v "^$" | tail -3
NS=1250000,10000000 CASES=256x10,256x120,128x10,64x10,8x10 timeout 300 python scripts/probe_step.py 2>&1 | tail -10
