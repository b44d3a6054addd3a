#!/bin/bash
mkdir -p gpurun_out
SG_TC_EXPERIMENT=2 B=256 N=4000000 timeout 300 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:scan_tc2 -s 1 -c 1 python scripts/prof_tc_target.py 2>&1 | grep -E "tensor|cycles|duration" 
B=256 N=4000000 timeout 300 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:scan_tc2 -s 1 -c 1 python scripts/prof_tc_target.py 2>&1 | grep -E "tensor|cycles|duration"
