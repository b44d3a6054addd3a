#!/bin/bash
# round 2, call ah: compute-sanitizer racecheck (shared-memory hazards) on the sweep's seed exchange and the encoder kernels
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 python -m pytest tests/test_scan_tc_gpu.py -q -x --tb=line -m gpu -p no:cacheprovider -k "matches_oracle or large_k or duplicates" > gpurun_out/ah_tc_racecheck.log 2>&1; echo "tc racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|Race reported|hazard" gpurun_out/ah_tc_racecheck.log | sort | uniq -c | sort -rn | head -12
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 python -m pytest tests/test_encoder_gpu.py -q -x --tb=line -m gpu -p no:cacheprovider -k "config5 or all_masked" > gpurun_out/ah_enc_racecheck.log 2>&1; echo "enc racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|Race reported|hazard" gpurun_out/ah_enc_racecheck.log | sort | uniq -c | sort -rn | head -12
