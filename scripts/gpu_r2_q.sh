#!/bin/bash
# round 2, call q: persistent double-buffered attention, GELU trim, no per-forward memset
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -8
echo "== default"; timeout 200 python scripts/probe_enc.py 2>&1 | tail -3
for m in 2 8; do echo "== skip $m"; SG_ENC_SKIP=$m timeout 200 python scripts/probe_enc.py 2>&1 | tail -1; done
