#!/bin/bash
# round 2, call E: small-shard behaviour (the 8-GPU shard shape on one GPU) + PQ tests
mkdir -p gpurun_out
CASES=256x10,128x10 NS=1250000 REPS=300 timeout 200 python scripts/probe_step.py 2>&1 | tail -3
SG_TC_COOP=0 CASES=256x10 NS=1250000 REPS=300 timeout 200 python scripts/probe_step.py 2>&1 | tail -2
B=256 N=1250000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_tc2 -s 1 -c 1 -f -o gpurun_out/e_tc2_small python scripts/prof_tc_target.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_pq_gpu.py tests/test_spann_gpu.py tests/test_index_golden.py -q -x --tb=short -m gpu -p no:cacheprovider > gpurun_out/e_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/e_tests.log; tail -5 gpurun_out/e_tests.log
CASES=2x10,3x10,4x10 NS=10000000 timeout 200 python scripts/probe_step.py 2>&1 | tail -4
SG_TC_MIN_B=2 CASES=2x10,3x10,4x10 NS=10000000 timeout 200 python scripts/probe_step.py 2>&1 | tail -4
