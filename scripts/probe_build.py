"""Graph build variants vs recall (tuning knobs are read by sg_store_build_graph from the environment)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shodh_memory_b200 as sg
from oracle import oracle as o

n, dim = int(os.environ.get("ROWS", 300_000)), 384
x = o.manifold_corpus(31, n, dim)
rng = np.random.default_rng(1)
q = x[rng.integers(0, n, 256)] + 0.02 * rng.standard_normal((256, dim)).astype(np.float32)
q = np.ascontiguousarray(q / np.linalg.norm(q, axis=1, keepdims=True), np.float32)
st = sg.GpuVectorStore(dim, "f16")
st.add_vectors(x)
ex, _, ec = st.search_batch(q, 10)
variants = [dict(), dict(SG_GRAPH_BACKEDGE_BY_ID="1"), dict(SG_GRAPH_BATCH="2048"), dict(SG_GRAPH_PASSES="3"),
            dict(SG_GRAPH_INBOX="96"), dict(SG_GRAPH_BATCH="2048", SG_GRAPH_INBOX="96"),
            dict(SG_GRAPH_BATCH="1024", SG_GRAPH_PASSES="3", SG_GRAPH_INBOX="96")]
for v in variants:
    for k in ("SG_GRAPH_BACKEDGE_BY_ID", "SG_GRAPH_BATCH", "SG_GRAPH_PASSES", "SG_GRAPH_INBOX"):
        os.environ.pop(k, None)
    os.environ.update(v)
    t0 = time.perf_counter(); st.build_graph(); bs = time.perf_counter() - t0
    deg, _, _ = st.get_graph()
    out = []
    for beam in (40, 120):
        a, _, ac, stt = st.search_ann_batch(q, beam)
        hits = sum(len(set(a[i, :10].tolist()) & set(ex[i, :ec[i]].tolist())) for i in range(256))
        out.append(f"beam {beam}: recall@10 {hits / ec.sum():.3f} ({stt[:,1].mean():.0f} evals)")
    print(v or "default", f"build {bs:.2f}s mean degree {deg.mean():.1f} min {deg.min()}", "; ".join(out), flush=True)
