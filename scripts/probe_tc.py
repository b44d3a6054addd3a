"""Scratch probe: tensor-core sweep timing at 10M rows for a few batch sizes."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
import shodh_memory_b200 as sg

n = int(os.environ.get("N", 10_000_000))
st = sg.GpuVectorStore(384, "f16")
st.append_synthetic(23821, 0, n)
g = torch.Generator().manual_seed(1)
q = torch.nn.functional.normalize(torch.randn(512, 384, generator=g), dim=1).cuda()
stream = torch.cuda.current_stream().cuda_stream
st.set_profiling(True)
bytes_sweep = n * 384 * 2
for b in [int(x) for x in os.environ.get("BS", "5,16,64,128,256,512").split(",")]:
    k = 10
    ids = torch.empty((b, k), dtype=torch.int32, device="cuda"); dist = torch.empty((b, k), device="cuda")
    cnt = torch.empty(b, dtype=torch.int32, device="cuda")
    for _ in range(3):
        st.search_device(q.data_ptr(), b, k, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream)
    torch.cuda.synchronize(); st.profile_read()
    s0 = st.tc_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        st.search_device(q.data_ptr(), b, k, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    ms, sweeps = st.profile_read()
    s1 = st.tc_stats()
    tot = e0.elapsed_time(e1) / reps
    m = float(np.mean(ms))
    print(json.dumps(dict(b=b, total_ms=round(tot, 4), scan_ms=round(m, 4), sweeps=sweeps, qps=round(b / tot * 1e3),
                          GBps_per_sweep=round(bytes_sweep * sweeps / m / 1e6, 1), flagged=s1[1] - s0[1],
                          served=s1[0] - s0[0])), flush=True)
