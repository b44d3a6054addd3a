"""Where does a beam-search step spend its time?  (profiling hook sg_store_ann_phase_cycles)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
import shodh_memory_b200 as sg
from oracle import oracle as o

n, dim = int(os.environ.get("ROWS", 300_000)), 384
x = o.manifold_corpus(31, n, dim)
st = sg.GpuVectorStore(dim, "f16")
st.add_vectors(x)
t0 = time.perf_counter(); st.build_graph(); print("build_s", time.perf_counter() - t0)
rng = np.random.default_rng(1)
NQ = 32768
q = x[rng.integers(0, n, NQ)] + 0.02 * rng.standard_normal((NQ, dim)).astype(np.float32)
q = np.ascontiguousarray(q / np.linalg.norm(q, axis=1, keepdims=True), np.float32)
names = ("pop", "list+visited", "row fetch", "scoring", "insertion")
for beam in (40, 120):
    for b in (1, 256, 1024, 8192, 32768):
        st.search_ann_batch(q[:b], beam)
        st.ann_phase_cycles(True)
        st.set_profiling(True); st.profile_read()
        reps = 20 if b <= 1024 else 3
        for _ in range(reps):
            _, _, _, stats = st.search_ann_batch(q[:b], beam)
        ms, _ = st.profile_read()
        cyc = st.ann_phase_cycles(False).astype(np.float64)
        hops = float(stats[:, 0].sum()) * reps
        gbs = float(stats[:, 1].sum()) * dim * 2 / (np.mean(ms) * 1e-3) / 1e9
        print(f"beam {beam} batch {b}: kernel {np.mean(ms):.3f} ms = {b / np.mean(ms) / 1e3:.2f} Mq/s, {gbs:.0f} GB/s of rows, {stats[:,0].mean():.0f} hops/query, cycles per hop: " +
              ", ".join(f"{nm} {c / hops:.0f}" for nm, c in zip(names, cyc)) + f"  (total {cyc.sum() / hops:.0f})")
