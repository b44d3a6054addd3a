#!/bin/bash
# round 2, call C: epilogue-optimised sweep -- parity, probes, ncu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_scan_tc_gpu.py tests/test_full_size_gpu.py tests/test_encoder_gpu.py tests/test_vama_file.py tests/test_spann_file.py -q -x --tb=short -m gpu -p no:cacheprovider -s > gpurun_out/c_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c_tests.log
tail -c 2500 gpurun_out/c_tests.log
CASES=256x10,256x120,128x10,8x10 NS=10000000,1250000 timeout 300 python scripts/probe_step.py 2>&1 | tail -8
B=256 N=4000000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_tc2 -s 1 -c 1 -f -o gpurun_out/c_tc2 python scripts/prof_tc_target.py 2>&1 | tail -1
B=128 N=4000000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 1 -c 1 -f -o gpurun_out/c_tc1 python scripts/prof_tc_target.py 2>&1 | tail -1
B=256 N=4000000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:finalize_tc -s 1 -c 1 -f -o gpurun_out/c_fin python scripts/prof_tc_target.py 2>&1 | tail -1
