#!/bin/bash
# round 2, call y: compute-sanitizer memcheck over the kernels rewritten this round (small shapes)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 120 --print-limit 20 python -m pytest tests/test_encoder_gpu.py -q -x --tb=line -m gpu -p no:cacheprovider -k "matches_torch or config5 or all_masked or pool_finalize" > gpurun_out/y_enc_memcheck.log 2>&1; echo "enc memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/y_enc_memcheck.log | tail -6
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 120 --print-limit 20 python -m pytest tests/test_scan_tc_gpu.py tests/test_store_growth_gpu.py -q -x --tb=line -m gpu -p no:cacheprovider -k "not accumulation_error and not pair_kernel_large and not copy_fallback" > gpurun_out/y_tc_memcheck.log 2>&1; echo "tc memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/y_tc_memcheck.log | tail -6
