"""Target for ncu captures of the graph-index kernels: build a 200k x 384 graph (not profiled), then -- between
cudaProfilerStart/Stop -- beam searches at batch 1, 256 (staged path) and 8192 (unstaged path), beam 40."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shodh_memory_b200 as sg
from oracle import oracle as o

n, dim = int(os.environ.get("ROWS", 200_000)), 384
x = o.manifold_corpus(31, n, dim)
st = sg.GpuVectorStore(dim, "f16")
st.add_vectors(x)
st.build_graph()
rng = np.random.default_rng(1)
q = x[rng.integers(0, n, 8192)] + 0.02 * rng.standard_normal((8192, dim)).astype(np.float32)
q = np.ascontiguousarray(q / np.linalg.norm(q, axis=1, keepdims=True), np.float32)
for b in (1, 256, 8192):
    st.search_ann_batch(q[:b], 40)          # warm (not profiled)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for b in (1, 256, 8192):
    _, _, _, stats = st.search_ann_batch(q[:b], 40)
    print("batch", b, "distance evaluations", int(stats[:, 1].sum()), "row bytes", int(stats[:, 1].sum()) * dim * 2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
