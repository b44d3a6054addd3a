"""Tiny target for `ncu --set full` on the tensor-core sweep."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shodh_memory_b200 as sg

n = int(os.environ.get("N", 10_000_000))
b = int(os.environ.get("B", 128))
st = sg.GpuVectorStore(384, "f16")
st.append_synthetic(23821, 0, n)
g = torch.Generator().manual_seed(1)
q = torch.nn.functional.normalize(torch.randn(512, 384, generator=g), dim=1).cuda()
stream = torch.cuda.current_stream().cuda_stream
k = 10
ids = torch.empty((b, k), dtype=torch.int32, device="cuda"); dist = torch.empty((b, k), device="cuda")
cnt = torch.empty(b, dtype=torch.int32, device="cuda")
for _ in range(2):
    st.search_device(q.data_ptr(), b, k, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream)
torch.cuda.synchronize()
print("done", st.tc_stats())
