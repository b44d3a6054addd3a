"""Tiny target for `ncu --set full`: a few exact sweeps over the 10M x 384 fp16 corpus."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shodh_memory_b200 as sg

n = int(os.environ.get("N", 10_000_000))
st = sg.GpuVectorStore(384, "f16")
st.append_synthetic(23821, 0, n)
g = torch.Generator().manual_seed(1)
q = torch.nn.functional.normalize(torch.randn(256, 384, generator=g), dim=1).cuda()
stream = torch.cuda.current_stream().cuda_stream
for b, k in [(1, 100)] * 3 + [(4, 10)] * 2 + ([(256, 10)] if os.environ.get("B256") else []):
    ids = torch.empty((b, k), dtype=torch.int32, device="cuda"); dist = torch.empty((b, k), device="cuda")
    cnt = torch.empty(b, dtype=torch.int32, device="cuda")
    st.search_device(q.data_ptr(), b, k, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream)
torch.cuda.synchronize()
print("done")
