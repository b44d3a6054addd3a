#!/bin/bash
# round 2, call n: ncu --set full of one encoder layer at 256 queries (pair GEMMs, ldmatrix attention, LN GEMMs)
mkdir -p gpurun_out
B=256 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"gemm_tc|attention_mma" -s 30 -c 5 -f -o gpurun_out/n_enc python scripts/prof_enc_target.py 2>&1 | tail -3
