#!/bin/bash
# round 2, call r: sweep phase trace at the 8-GPU shard shape and at 10M rows
mkdir -p gpurun_out
NS=1250000 CASES=256x10 REPS=4 SG_TC_TRACE=1 timeout 300 python scripts/probe_step.py 2>&1 | grep -v "^$" | tail -9
NS=10000000 CASES=256x10 REPS=2 SG_TC_TRACE=1 timeout 300 python scripts/probe_step.py 2>&1 | grep -v "^$" | tail -4
NS=1250000,10000000 CASES=256x10,128x10,8x10 timeout 300 python scripts/probe_step.py 2>&1 | tail -6
