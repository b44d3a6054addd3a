"""Scratch probe: host-buffer search (sg_store_search) per-call time vs the device-resident step, with SG_HOST_TRACE=1."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shodh_memory_b200 as sg
n = int(os.environ.get("N", 10_000_000))
st = sg.GpuVectorStore(384, "f16"); st.append_synthetic(23821, 0, n)
q = torch.nn.functional.normalize(torch.randn(256, 384, generator=torch.Generator().manual_seed(1)), dim=1)
qp = q.pin_memory().numpy(); qn = q.numpy().copy()
for name, arr in (("pinned", qp), ("pageable", qn)):
    for _ in range(5): st.search_batch(arr, 10)
    t0 = time.perf_counter()
    for _ in range(64): st.search_batch(arr, 10)
    print(name, "ms per call", (time.perf_counter() - t0) / 64 * 1e3, flush=True)
dq = q.cuda(); ids = torch.empty((256, 10), dtype=torch.int32, device="cuda"); dist = torch.empty((256, 10), device="cuda"); cnt = torch.empty(256, dtype=torch.int32, device="cuda")
cs = torch.cuda.current_stream().cuda_stream
def dev(): st.search_device(dq.data_ptr(), 256, 10, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), cs)
for _ in range(5): dev()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(64): dev()
torch.cuda.synchronize(); print("device-resident, back to back: ms per call", (time.perf_counter() - t0) / 64 * 1e3, flush=True)
t0 = time.perf_counter()
for _ in range(64):
    dev(); torch.cuda.synchronize()
print("device-resident, synchronised after every call: ms per call", (time.perf_counter() - t0) / 64 * 1e3, flush=True)
