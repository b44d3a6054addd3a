#!/bin/bash
# round 2, call l: in-place growth tests, full GPU suite, SPANN leg after the rank-kernel operand batching
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_store_growth_gpu.py -q -x --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -15
timeout 1500 python -m pytest tests -q --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/l_tests.log; tail -8 gpurun_out/l_tests.log
timeout 300 python bench.py --leg spann > gpurun_out/l_spann.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/l_spann.json')); print('SPANN', d.get('error') or {k:d[k] for k in ('batch256','batch1_host_api_ms','achieved_GBps_batch256')})"
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3
CASES="10000000:256:10,10000000:128:10,10000000:8:10,1250000:256:10" timeout 300 python scripts/probe_step.py 2>&1 | tail -5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_enc_launches.csv python scripts/prof_enc_target.py > /dev/null 2>&1; wc -l gpurun_out/l_enc_launches.csv
