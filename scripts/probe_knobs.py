"""Scratch probe: ring depth / K-blocks per stage of the pair sweep at the bench shape, alternating in one process."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shodh_memory_b200 as sg
stream = torch.cuda.current_stream().cuda_stream
q = torch.nn.functional.normalize(torch.randn(512, 384, generator=torch.Generator().manual_seed(1)), dim=1).cuda()
st = sg.GpuVectorStore(384, "f16"); st.append_synthetic(23821, 0, 10_000_000); st.set_profiling(True)
b = 256
ids = torch.empty((b, 10), dtype=torch.int32, device="cuda"); dist = torch.empty((b, 10), device="cuda"); cnt = torch.empty(b, dtype=torch.int32, device="cuda")
cfgs = [("8", "2"), ("6", "2"), ("4", "2"), ("8", "1"), ("6", "1")]
res = {c: [] for c in cfgs}
for rnd in range(3):
    for c in cfgs:
        os.environ["SG_TC_STAGES"], os.environ["SG_TC_KBS"] = c
        for _ in range(3): st.search_device(q.data_ptr(), b, 10, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream)
        torch.cuda.synchronize(); st.profile_read()
        for _ in range(12): st.search_device(q.data_ptr(), b, 10, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream)
        torch.cuda.synchronize(); ms, _ = st.profile_read()
        res[c].append(round(float(np.mean(ms)), 4))
for c in cfgs: print("K-blocks in the ring", c[0], "per stage", c[1], "sweep ms per round:", res[c], flush=True)
