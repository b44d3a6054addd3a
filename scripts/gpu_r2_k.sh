#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pq_gpu.py tests/test_spann_gpu.py tests/test_spann_file.py tests/test_index_golden.py -q -x --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 300 python bench.py --leg pq > gpurun_out/k_pq.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/k_pq.json')); print('PQ', d.get('error') or {k:d[k] for k in ('batch256','batch1_host_api_ms')})"
SG_BENCH_PQ_ROWS=10000000 timeout 300 python bench.py --leg pq > gpurun_out/k_pq10m.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/k_pq10m.json')); print('PQ10M', d.get('error') or {k:d[k] for k in ('batch256','batch1_host_api_ms')})"
timeout 300 python bench.py --leg spann > gpurun_out/k_spann.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/k_spann.json')); print('SPANN', d.get('error') or {k:d[k] for k in ('batch256','batch1_host_api_ms','achieved_GBps_batch256','cpu_reference')})"
cat > /tmp/pq_target.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import shodh_memory_b200 as sg
rng = np.random.default_rng(0)
pq = sg.GpuProductQuantizer.from_centroids((rng.standard_normal((48, 256, 8)) * 0.05).astype(np.float32))
for lo in range(0, 2_000_000, 200_000):
    pq.add(np.arange(lo, lo + 200_000, dtype=np.uint32), rng.standard_normal((200_000, 384)).astype(np.float32) * 0.05)
q = rng.standard_normal((8, 384)).astype(np.float32) * 0.05
pq.search_batch(q, 10)
P = 1415
cent = rng.standard_normal((P, 384)).astype(np.float32); cent /= np.linalg.norm(cent, axis=1, keepdims=True)
off = np.linspace(0, 2_000_000, P + 1).astype(np.uint64)
sp = sg.GpuSpannIndex.from_arrays(cent, (rng.standard_normal((48, 256, 8)) * 0.05).astype(np.float32), off,
                                  np.arange(2_000_000, dtype=np.uint32), rng.integers(0, 256, (2_000_000, 48), dtype=np.uint8))
sp.search_batch(rng.standard_normal((256, 384)).astype(np.float32), 10)
torch.cuda.profiler.start()
pq.search_batch(q, 10)
sp.search_batch(rng.standard_normal((256, 384)).astype(np.float32), 10)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
PY
timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"pq_scan|spann_search|spann_rank" -c 6 -f -o gpurun_out/k_pq_spann python /tmp/pq_target.py 2>&1 | tail -2
