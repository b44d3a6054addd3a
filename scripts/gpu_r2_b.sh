#!/bin/bash
# round 2, call B: whole GPU suite on the rewritten path, knob probes, ncu captures, first bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x --tb=short -m gpu -p no:cacheprovider -s > gpurun_out/b_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/b_tests.log
tail -c 2500 gpurun_out/b_tests.log
for st in 8 6; do
  echo "== stages $st"; SG_TC_STAGES=$st CASES=128x10,256x10,8x10 NS=10000000,1250000 timeout 200 python scripts/probe_step.py 2>&1 | tail -6
done
# launch list of one step at the 8-GPU shard shape and at full size (shares only)
CASES=256x10 NS=1250000 REPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/b_launches_1250k.csv python scripts/probe_step.py > /dev/null 2>&1
# full captures: pair kernel (b=256) and 1-CTA kernel (b=128) at 4M rows (comparable with profiles/r01_scan_tc2_ncu.md)
B=256 N=4000000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_tc2 -s 1 -c 1 -f -o gpurun_out/b_tc2 python scripts/prof_tc_target.py 2>&1 | tail -2
B=128 N=4000000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 1 -c 1 -f -o gpurun_out/b_tc1 python scripts/prof_tc_target.py 2>&1 | tail -2
B=256 N=4000000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:finalize_tc -s 1 -c 1 -f -o gpurun_out/b_fin python scripts/prof_tc_target.py 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
echo "bench exit $?"; tail -c 600 gpurun_out/b_bench.err; head -c 3000 gpurun_out/b_bench.json
