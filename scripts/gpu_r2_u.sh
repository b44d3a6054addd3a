#!/bin/bash
# round 2, call u: validation pass -- full GPU suite, smoke, bench (N=1, with the 8(f) legs), launch list, ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/u_tests.log; tail -6 gpurun_out/u_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 --legs > gpurun_out/u_bench1.json 2> gpurun_out/u_bench1.err; echo "bench exit $?"; tail -c 400 gpurun_out/u_bench1.err
cp profiles/r02_legs.json gpurun_out/u_legs.json 2>/dev/null
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/u_ref.json 2> gpurun_out/u_ref.err; echo "ref exit $?"
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/u_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/u_bench_under_ncu.log 2>&1; wc -l gpurun_out/u_launches.csv
# full captures: pair sweep + finalize at the bench shape (10M rows, 256 queries)
B=256 N=10000000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan_tc2|finalize_tc" -s 2 -c 2 -f -o gpurun_out/u_sweep10m python scripts/prof_tc_target.py 2>&1 | tail -2
# batch-1 exact sweep at 10M
timeout 600 ncu --set full --clock-control none -k regex:"scan_exact" -c 1 -f -o gpurun_out/u_exact10m python - <<'PY' 2>&1 | tail -2
import sys, torch
sys.path.insert(0, ".")
import shodh_memory_b200 as sg
st = sg.GpuVectorStore(384, "f16"); st.append_synthetic(23821, 0, 10_000_000)
q = torch.nn.functional.normalize(torch.randn(4, 384, generator=torch.Generator().manual_seed(1)), dim=1).cuda()
ids = torch.empty((1, 100), dtype=torch.int32, device="cuda"); dist = torch.empty((1, 100), device="cuda"); cnt = torch.empty(1, dtype=torch.int32, device="cuda")
st.search_device(q.data_ptr(), 1, 100, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
PY
# one encoder layer at 256 queries
B=256 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"gemm_tc|gemm_ln|attention_mma" -s 30 -c 5 -f -o gpurun_out/u_enc python scripts/prof_enc_target.py 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/u_enc_launches.csv python scripts/prof_enc_target.py > /dev/null 2>&1; wc -l gpurun_out/u_enc_launches.csv
