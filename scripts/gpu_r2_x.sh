#!/bin/bash
# round 2, call x: split/persistent LayerNorm GEMM; page-locked queries read in place
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py tests/test_scan_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -8
echo "== default"; timeout 200 python scripts/probe_enc.py 2>&1 | tail -3
echo "== no ln split"; SG_ENC_NO_LN_SPLIT=1 timeout 200 python scripts/probe_enc.py 2>&1 | tail -1
for m in 4 16; do echo "== skip $m"; SG_ENC_SKIP=$m timeout 200 python scripts/probe_enc.py 2>&1 | tail -1; done
