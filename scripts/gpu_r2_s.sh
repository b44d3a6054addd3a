#!/bin/bash
# round 2, call s: cheaper slow path, one-round-trip refined seed, pair kernel from b >= 97
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scan_tc_gpu.py tests/test_scan_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -6
NS=1250000 CASES=256x10 REPS=2 SG_TC_TRACE=1 timeout 300 python scripts/probe_step.py 2>&1 | grep -v "^$" | tail -3
NS=1250000,10000000 CASES=256x10,256x120,128x10,96x10,64x10,32x10,8x10 timeout 300 python scripts/probe_step.py 2>&1 | tail -14
echo "== 1-CTA kernel for all b <= 128"
NS=10000000 CASES=128x10,96x10,64x10 SG_TC_PAIR_MIN_B=129 timeout 300 python scripts/probe_step.py 2>&1 | tail -3
echo "== pair kernel from b = 33"
NS=10000000,1250000 CASES=64x10,48x10 SG_TC_PAIR_MIN_B=33 timeout 300 python scripts/probe_step.py 2>&1 | tail -4
