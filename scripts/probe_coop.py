"""Scratch probe: cooperative vs plain sweep launch, host API and back-to-back device loop, alternating in ONE process."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shodh_memory_b200 as sg
n = int(os.environ.get("N", 10_000_000))
st = sg.GpuVectorStore(384, "f16"); st.append_synthetic(23821, 0, n)
q = torch.nn.functional.normalize(torch.randn(256, 384, generator=torch.Generator().manual_seed(1)), dim=1)
qp = q.pin_memory().numpy()
dq = q.cuda(); ids = torch.empty((256, 10), dtype=torch.int32, device="cuda"); dist = torch.empty((256, 10), device="cuda"); cnt = torch.empty(256, dtype=torch.int32, device="cuda")
cs = torch.cuda.current_stream().cuda_stream
def dev(): st.search_device(dq.data_ptr(), 256, 10, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), cs)
for rnd in range(3):
    for coop in ("1", "0"):
        os.environ["SG_TC_COOP"] = coop
        for _ in range(4): st.search_batch(qp, 10)
        t0 = time.perf_counter()
        for _ in range(48): st.search_batch(qp, 10)
        host = (time.perf_counter() - t0) / 48 * 1e3
        for _ in range(4): dev()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(48): dev()
        torch.cuda.synchronize(); b2b = (time.perf_counter() - t0) / 48 * 1e3
        print(f"round {rnd} coop={coop}: host API {host:.4f} ms/call | device back-to-back {b2b:.4f} ms/call | gap {1e3*(host-b2b):.0f} us", flush=True)
