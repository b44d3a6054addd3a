"""Scratch probe: the batched search step (path B) at a given shard size -- total step vs sweep time, by k."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
import shodh_memory_b200 as sg

stream = torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(1)
q = torch.nn.functional.normalize(torch.randn(512, 384, generator=g), dim=1).cuda()
for n in [int(x) for x in os.environ.get("NS", "10000000,1250000").split(",")]:
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(23821, 0, n)
    st.set_profiling(True)
    cases = [(256, 10), (256, 120), (128, 10), (64, 120), (8, 10), (4, 10), (1, 100)]
    if os.environ.get("CASES"):
        cases = [tuple(int(v) for v in c.split("x")) for c in os.environ["CASES"].split(",")]
    for b, k in cases:
        ids = torch.empty((b, k), dtype=torch.int32, device="cuda"); dist = torch.empty((b, k), device="cuda")
        cnt = torch.empty(b, dtype=torch.int32, device="cuda")
        for _ in range(3):
            st.search_device(q.data_ptr(), b, k, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream)
        torch.cuda.synchronize(); st.profile_read()
        s0 = st.tc_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = int(os.environ.get("REPS", 20))
        e0.record()
        for _ in range(reps):
            st.search_device(q.data_ptr(), b, k, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream)
        e1.record(); torch.cuda.synchronize()
        ms, sweeps = st.profile_read()
        s1 = st.tc_stats()
        tot = e0.elapsed_time(e1) / reps
        m = float(np.mean(ms))
        print(json.dumps(dict(n=n, b=b, k=k, step_ms=round(tot, 4), scan_ms=round(m, 4), other_us=round((tot - m) * 1e3, 1),
                              sweeps=sweeps, qps=round(b / tot * 1e3), GBps=round(n * 768 * sweeps / m / 1e6, 1),
                              served=s1[0] - s0[0], flagged=s1[1] - s0[1])), flush=True)
    st.close()
