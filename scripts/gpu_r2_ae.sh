#!/bin/bash
mkdir -p gpurun_out
NS=1250000 CASES=256x120 REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:finalize_tc -s 3 -c 1 -f -o gpurun_out/ae_fin120 python scripts/probe_step.py 2>&1 | tail -2
