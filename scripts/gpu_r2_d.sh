#!/bin/bash
# round 2, call D: 8-warp epilogue + pipelined threshold loads + K-block rotation
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_scan_tc_gpu.py tests/test_full_size_gpu.py tests/test_scan_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider > gpurun_out/d_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/d_tests.log
tail -c 1500 gpurun_out/d_tests.log
for rot in 1; do
  echo "== kbrot $rot"; SG_TC_KBROT=$rot CASES=256x10,256x120,128x10,8x10 NS=10000000,1250000 timeout 300 python scripts/probe_step.py 2>&1 | tail -8
done
B=256 N=4000000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_tc2 -s 1 -c 1 -f -o gpurun_out/d_tc2 python scripts/prof_tc_target.py 2>&1 | tail -1
B=128 N=4000000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 1 -c 1 -f -o gpurun_out/d_tc1 python scripts/prof_tc_target.py 2>&1 | tail -1
