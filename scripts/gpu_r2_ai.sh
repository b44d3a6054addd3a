#!/bin/bash
# round 2, call ai: bounded address reservation + zero-copy re-mapping of the row chunks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_store_growth_gpu.py tests/test_scan_gpu.py tests/test_vamana_gpu.py tests/test_vama_file.py -q -x --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -6
