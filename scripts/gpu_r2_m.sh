#!/bin/bash
# round 2, call m: encoder after pair GEMM + ldmatrix attention + PDL chain
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -8
echo "== default"; timeout 200 python scripts/probe_enc.py 2>&1 | tail -3
echo "== no pair"; SG_ENC_NO_PAIR=1 timeout 200 python scripts/probe_enc.py 2>&1 | tail -3
echo "== no graph"; SG_ENC_GRAPH=0 timeout 200 python scripts/probe_enc.py 2>&1 | tail -3
for m in 1 2 4 8 16; do echo "== skip $m"; SG_ENC_SKIP=$m timeout 200 python scripts/probe_enc.py 2>&1 | tail -3; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/m_enc_launches.csv python scripts/prof_enc_target.py > /dev/null 2>&1; wc -l gpurun_out/m_enc_launches.csv
