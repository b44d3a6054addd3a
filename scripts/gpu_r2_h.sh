#!/bin/bash
mkdir -p gpurun_out
SG_TC_TRACE=1 CASES=256x10 NS=1250000 REPS=1 timeout 200 python scripts/probe_step.py 2>&1 | grep -v "^{" | tail -3
SG_TC_TRACE=1 CASES=128x10,256x120 NS=1250000 REPS=1 timeout 200 python scripts/probe_step.py 2>&1 | grep -v "^{" | tail -6
timeout 600 python -m pytest tests/test_scan_tc_gpu.py tests/test_full_size_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -3
CASES=256x10,256x120,128x10 NS=1250000,10000000 REPS=200 timeout 300 python scripts/probe_step.py 2>&1 | tail -6
