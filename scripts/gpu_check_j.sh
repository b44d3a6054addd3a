#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ann_search -c 3 -f -o gpurun_out/ann_r01 python scripts/prof_ann_target.py 2>&1 | tail -6
ls -la gpurun_out/ann_r01.ncu-rep
