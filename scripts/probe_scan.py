"""Scratch tuning probe (not the bench): times the exact sweep for a few (warps, ring) shapes."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
import shodh_memory_b200 as sg
from oracle import oracle as o

n = int(os.environ.get("N", 10_000_000))
st = sg.GpuVectorStore(384, "f16")
t0 = time.time(); st.append_synthetic(23821, 0, n); print("fill s", time.time() - t0, flush=True)
q, _ = o.synth_queries(23821, n, 8, 384)
dq = torch.tensor(q).cuda()
stream = torch.cuda.current_stream().cuda_stream
st.set_profiling(True)
bytes_sweep = n * 384 * 2
for (w, r) in [(8, 6), (4, 6), (5, 6), (6, 6), (8, 4), (10, 4), (12, 4), (6, 5), (8, 5), (16, 3)]:
    os.environ["SG_SCAN_WARPS"] = str(w); os.environ["SG_SCAN_RING"] = str(r)
    for b, k in [(1, 10), (1, 100), (2, 10), (4, 10)]:
        ids = torch.empty((b, k), dtype=torch.int32, device="cuda"); dist = torch.empty((b, k), device="cuda")
        cnt = torch.empty(b, dtype=torch.int32, device="cuda")
        try:
            for _ in range(3):
                st.search_device(dq.data_ptr(), b, k, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream)
            torch.cuda.synchronize()
            ms = []
            for _ in range(40):
                st.search_device(dq.data_ptr(), b, k, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream)
                torch.cuda.synchronize()
                ms.append(st.last_scan_ms()[0])
            m = float(np.median(ms))
            print(json.dumps(dict(W=w, ring=r, b=b, k=k, scan_ms=round(m, 4), GBps=round(bytes_sweep / m / 1e6, 1),
                                  frac=round(bytes_sweep / m / 1e6 / 6569.3, 3))), flush=True)
        except Exception as e:
            print("ERR", w, r, b, k, e, flush=True)
