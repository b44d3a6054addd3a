#!/bin/bash
# round 2, call ab: fused QKV + attention kernel for small batches
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -6
echo "== default"; timeout 200 python scripts/probe_enc.py 2>&1 | tail -3
echo "== separate QKV GEMM + attention"; SG_ENC_NO_FUSED_QKV=1 timeout 200 python scripts/probe_enc.py 2>&1 | tail -3 | head -1
