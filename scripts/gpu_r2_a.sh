#!/bin/bash
# round 2, call A: the rewritten tensor-core path -- parity first, then step timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scan_tc_gpu.py tests/test_scan_gpu.py tests/test_full_size_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider -s > gpurun_out/a_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/a_tests.log
tail -c 3000 gpurun_out/a_tests.log
timeout 300 python scripts/probe_step.py > gpurun_out/a_probe.log 2>&1
cat gpurun_out/a_probe.log | tail -20
SG_TC_COOP=0 NS=1250000 timeout 200 python scripts/probe_step.py > gpurun_out/a_probe_nocoop.log 2>&1
tail -8 gpurun_out/a_probe_nocoop.log
