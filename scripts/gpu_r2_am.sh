#!/bin/bash
# round 2, call am: compute-sanitizer synccheck (barrier usage) on the kernels with named barriers / cluster syncs
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_encoder_gpu.py -q -x --tb=line -m gpu -p no:cacheprovider -k "matches_torch or config5" > gpurun_out/am_enc_synccheck.log 2>&1; echo "enc synccheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/am_enc_synccheck.log | tail -3
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_scan_tc_gpu.py -q -x --tb=line -m gpu -p no:cacheprovider -k "matches_oracle or large_k or duplicates" > gpurun_out/am_tc_synccheck.log 2>&1; echo "tc synccheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/am_tc_synccheck.log | tail -3
