#!/bin/bash
# call A: the new GPU tests only (metrics + graph index), all failures shown
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scan_metrics_gpu.py tests/test_vamana_gpu.py tests/test_host_cpp.py -q --tb=short -m gpu -p no:cacheprovider -s > gpurun_out/new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/new_tests.log
tail -c 6000 gpurun_out/new_tests.log
