#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scan_tc_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider -k "matches_oracle or pair_kernel or large_k" 2>&1 | tail -3
timeout 300 python scripts/probe_coop.py 2>&1 | tail -7
NS=1250000 CASES=256x10 timeout 200 python scripts/probe_step.py 2>&1 | tail -1
