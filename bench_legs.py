"""bench_legs.py -- the SURVEY 8(f) legs of the benchmark (graph index, Product Quantization, SPANN, CPU detail).

`python bench.py --legs` runs them after the headline measurement, each GPU leg in its OWN child process
(`bench.py --leg NAME`): a fault or a hang inside one of them cannot take the headline line down with it.  The results
go to profiles/r02_legs.json, not to the JSON line the driver reads.  The CPU sides use oracle/ as the timed baseline /
checker only."""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from bench import BATCH, DIM, ROWS, SEED, TOPK, UNIT, host_threads, peaks  # noqa: E402


def cpu_detail(budget_rows: int = 100_000):
    """SURVEY 8(d) CPU baselines beside the headline one, all on a small row sample and scaled linearly to the
    full corpus: the reference loop single-threaded (it IS sequential, vamana.rs:1173) as written and with the
    incidental waste removed (heap top-k, no row clone), and the waste-free loop on all host threads."""
    from oracle import oracle as o
    threads = host_threads()
    corpus = o.synth_corpus(SEED, 0, budget_rows, DIM, f16_round=True)
    queries, _ = o.synth_queries(SEED, ROWS, max(threads, 4), DIM)
    scale = budget_rows / ROWS
    out = {}

    def run(name, nq, as_written, thr):
        o.brute_force_search_batch(corpus, queries[:nq], TOPK, as_written=as_written, threads=thr)
        t0 = time.perf_counter()
        o.brute_force_search_batch(corpus, queries[:nq], TOPK, as_written=as_written, threads=thr)
        out[name] = nq / (time.perf_counter() - t0) * scale
    run("as_written_1_thread", 4, True, 1)
    run("heap_topk_no_clone_1_thread", 4, False, 1)
    run(f"heap_topk_no_clone_{threads}_threads", threads, False, threads)
    out["unit"] = UNIT
    out["sample"] = f"{budget_rows} rows of {ROWS}, scaled linearly"
    return out


def cpu_encoder_ms(L: int, threads: int):
    """PyTorch-CPU fp32 forward of the MiniLM-L6 architecture (the checker model of tests/test_encoder_gpu.py),
    one query of L tokens, `threads` intra-op threads (reference default: 2, minilm.rs:67)."""
    try:
        import torch
        from transformers import BertConfig, BertModel
    except Exception:
        return None
    cfg = BertConfig(vocab_size=30522, hidden_size=384, num_hidden_layers=6, num_attention_heads=12,
                     intermediate_size=1536, max_position_embeddings=512)
    model = BertModel(cfg, add_pooling_layer=False).eval()
    prev = torch.get_num_threads()
    torch.set_num_threads(threads)
    ids = torch.randint(1000, 30522, (1, L))
    mask = torch.ones((1, L), dtype=torch.int64)
    with torch.no_grad():
        for _ in range(2):
            model(input_ids=ids, attention_mask=mask)
        t0 = time.perf_counter()
        for _ in range(5):
            model(input_ids=ids, attention_mask=mask)
        ms = (time.perf_counter() - t0) * 1e3 / 5
    torch.set_num_threads(prev)
    return ms


def leg_pq(sg, local_rank):
    """Product Quantization: CompressedVectorStore::search on the device."""
    from oracle import oracle as o
    pq_rows = int(os.environ.get("SG_BENCH_PQ_ROWS", 1_000_000))
    rng = np.random.default_rng(SEED + 11)
    A = rng.standard_normal((12, DIM)).astype(np.float32)

    def pq_rows_gen(m):
        x = rng.standard_normal((m, 12)).astype(np.float32) @ A + 0.17 * rng.standard_normal((m, DIM)).astype(np.float32)
        return np.ascontiguousarray(x / np.linalg.norm(x, axis=1, keepdims=True), np.float32)
    train = pq_rows_gen(5000)
    orders = np.stack([rng.permutation(5000) for _ in range(DIM // 8)]).astype(np.uint32)
    t0 = time.perf_counter()
    pq = sg.GpuProductQuantizer.train(train, orders, kmeans_iterations=20, device=local_rank)   # PQConfig::minilm()
    train_s = time.perf_counter() - t0
    t_add = 0.0
    first_chunk = None
    for lo in range(0, pq_rows, 100_000):
        m = min(100_000, pq_rows - lo)
        x = pq_rows_gen(m)
        if first_chunk is None:
            first_chunk = x
        t0 = time.perf_counter()
        pq.add(np.arange(lo, lo + m, dtype=np.uint32), x)
        t_add += time.perf_counter() - t0
    pq_q = np.ascontiguousarray(first_chunk[:BATCH] + 0.01 * rng.standard_normal((BATCH, DIM)).astype(np.float32), np.float32)
    for _ in range(2):
        pq.search_batch(pq_q, TOPK)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        gi, gd, gc = pq.search_batch(pq_q, TOPK)
    ms = (time.perf_counter() - t0) * 1e3 / reps
    t0 = time.perf_counter()
    for i in range(20):
        pq.search_batch(pq_q[i:i + 1], TOPK)
    ms1 = (time.perf_counter() - t0) * 1e3 / 20
    # CPU: the oracle port of CompressedVectorStore::search on a 100k-vector slice of the same store, 1 thread
    ref = o.PQOracle(DIM)
    ref.centroids, ref.k = pq.centroids, pq.num_centroids
    ref.ids = np.arange(100_000, dtype=np.uint32)
    ref.codes = pq.encode(first_chunk)            # device codes == oracle codes (tests/test_pq_gpu.py); saves 10 s of host encode
    t0 = time.perf_counter()
    for i in range(4):
        ref.search(pq_q[i], TOPK)
    cpu_ms = (time.perf_counter() - t0) * 1e3 / 4 * (pq_rows / 100_000)
    sweeps = (BATCH + 3) // 4                     # pq_scan4_kernel: four queries share one pass over the codes
    nbytes = float(pq_rows) * (DIM // 8 + 4)
    pq_res = dict(rows=pq_rows, code_bytes=DIM // 8, train_s_5000x20=train_s, encode_add_rows_per_s=pq_rows / t_add,
                  batch256=dict(host_api_ms=ms, queries_per_s=BATCH / (ms * 1e-3), sweeps=sweeps,
                                roofline=dict(bound="hbm (code stream) / shared-memory gathers", algorithmic_bytes_per_sweep=nbytes,
                                              achieved=nbytes * sweeps / (ms * 1e-3) / 1e9, peak=peaks()[0], unit="GB/s",
                                              frac=nbytes * sweeps / (ms * 1e-3) / 1e9 / peaks()[0],
                                              note="wall clock of sg_pq_search incl. H2D/D2H and table build; 4 queries per sweep; the sweep is bound by its "
                                                   "48 shared-memory table look-ups per vector and query, not by the code stream")),
                  batch1_host_api_ms=ms1,
                  cpu_reference=dict(ms_per_query_1_thread=cpu_ms, kind="port (oracle/pq_oracle.c), 100k-vector slice scaled linearly"),
                  kernel="pq_scan4_kernel (interleaved float4 table: one LDS.128 serves four queries; 48 look-ups per vector in the reference's order; ids and distances bit-identical to the CPU path)")
    pq.close()
    return pq_res


def leg_spann(sg, local_rank):
    """SPANN: one CTA per query over a 10M-entry partitioned PQ index.  The index is synthetic at full
    scale (random unit centroids, random codebooks and codes, equal-sized posting lists): the search cost depends on the
    shapes (P = ceil(sqrt(n)) centroids, num_probes x n / P entries per query), not on the values; 4 queries are
    checked id-for-id against the oracle port of SpannIndex::search on the same arrays."""
    import ctypes as C
    from oracle import oracle as o
    sp_rows = int(os.environ.get("SG_BENCH_SPANN_ROWS", 10_000_000))
    rng = np.random.default_rng(SEED + 13)
    P = int(np.ceil(np.sqrt(sp_rows)))
    cent = rng.standard_normal((P, DIM)).astype(np.float32)
    cent /= np.linalg.norm(cent, axis=1, keepdims=True)
    pqc = (rng.standard_normal((DIM // 8, 256, 8)) * 0.05).astype(np.float32)
    off = np.linspace(0, sp_rows, P + 1).astype(np.uint64)
    ids_sp = np.arange(sp_rows, dtype=np.uint32)
    codes_sp = rng.integers(0, 256, (sp_rows, DIM // 8), dtype=np.uint8)
    t0 = time.perf_counter()
    sp = sg.GpuSpannIndex.from_arrays(cent, pqc, off, ids_sp, codes_sp, num_probes=10, device=local_rank)
    load_s = time.perf_counter() - t0
    sq = rng.standard_normal((BATCH, DIM)).astype(np.float32)
    sq = np.ascontiguousarray(sq / np.linalg.norm(sq, axis=1, keepdims=True), np.float32)
    for _ in range(3):
        gi, gd, gc = sp.search_batch(sq, TOPK)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        sp.search_batch(sq, TOPK)
    ms = (time.perf_counter() - t0) * 1e3 / reps
    t0 = time.perf_counter()
    for i in range(50):
        sp.search_batch(sq[i:i + 1], TOPK)
    ms1 = (time.perf_counter() - t0) * 1e3 / 50
    f32p, u32p, u64p, u8p = (C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8))
    same, cpu_ms = True, 0.0
    for i in range(4):
        oi = np.zeros(TOPK, np.uint32); od = np.zeros(TOPK, np.float32)
        t0 = time.perf_counter()
        r = o.lib().so_spann_search(cent.ctypes.data_as(f32p), P, DIM, 0, 10, off.ctypes.data_as(u64p),
                                    ids_sp.ctypes.data_as(u32p), codes_sp.ctypes.data_as(u8p), pqc.ctypes.data_as(f32p),
                                    DIM // 8, 256, 8, sq[i].ctypes.data_as(f32p), TOPK, oi.ctypes.data_as(u32p),
                                    od.ctypes.data_as(f32p), None)
        cpu_ms += (time.perf_counter() - t0) * 1e3 / 4
        same = same and r == int(gc[i]) and np.array_equal(gi[i, :r], oi[:r]) and \
            np.array_equal(gd[i, :r].view(np.uint32), od[:r].view(np.uint32))
    entries = 10.0 * sp_rows / P
    bytes_q = P * DIM * 4 + entries * (DIM // 8 + 4)
    spann_res = dict(rows=sp_rows, partitions=P, num_probes=10, entries_scanned_per_query=entries, upload_s=load_s,
                     batch256=dict(host_api_ms=ms, queries_per_s=BATCH / (ms * 1e-3)), batch1_host_api_ms=ms1,
                     algorithmic_bytes_per_query=bytes_q,
                     achieved_GBps_batch256=bytes_q * BATCH / (ms * 1e-3) / 1e9,
                     cpu_reference=dict(ms_per_query_1_thread=cpu_ms, ids_and_distances_identical_to_device=bool(same),
                                        kind="port (oracle/spann_oracle.c restatement of SpannIndex::search), 4 queries"),
                     kernel="spann_search_kernel (one CTA per query: ADC table, centroid ranking, probed posting lists, top-k)")
    sp.close()
    return spann_res


def leg_ann(sg, local_rank):
    """Graph index: device build (R 32, L 75, alpha 1.2) of a low-intrinsic-dimension corpus, then batch-256 and
    batch-1 beam searches with the reference's beams (k = 12 x max_results = 120 and search_ids' own 4 x limit = 40), recall@10
    against the exact sweep of the same store, and the oracle port of greedy_search on the same graph beside it."""
    import torch
    ann_res = None
    try:
        ann_rows = int(os.environ.get("SG_BENCH_ANN_ROWS", 1_000_000))
        rng = np.random.default_rng(SEED + 7)
        latent = 12
        A = rng.standard_normal((latent, DIM)).astype(np.float32)
        g = sg.GpuVectorStore(DIM, "f16", local_rank)
        g.reserve(ann_rows)
        for lo in range(0, ann_rows, 100_000):
            m = min(100_000, ann_rows - lo)
            x = rng.standard_normal((m, latent)).astype(np.float32) @ A + \
                (0.05 * np.sqrt(latent)) * rng.standard_normal((m, DIM)).astype(np.float32)
            x /= np.linalg.norm(x, axis=1, keepdims=True)
            g.add_vectors(x.astype(np.float32))
        t0 = time.perf_counter()
        g.build_graph(32, 75, 1.2, SEED)
        build_s = time.perf_counter() - t0
        picks = rng.integers(0, ann_rows, BATCH)
        aq = np.stack([g.get_vectors(int(j), 1)[0] for j in picks]) + 0.02 * rng.standard_normal((BATCH, DIM)).astype(np.float32)
        aq = np.ascontiguousarray(aq / np.linalg.norm(aq, axis=1, keepdims=True), np.float32)
        ids_e, _, cnt_e = g.search_batch(aq, TOPK)
        t0 = time.perf_counter()
        for _ in range(10):
            g.search_batch(aq, TOPK)
        exact_ms = (time.perf_counter() - t0) * 1e3 / 10
        g.set_profiling(True)
        beams = {}
        # beam 120 = the reference's request for top-10 (k = 12 x max_results: retrieval.rs:916 x memory/mod.rs:3585);
        # beam 40 = search_ids' own 4 x limit (retrieval.rs:915-916)
        for KB in (12 * TOPK, 4 * TOPK):
            ids_a, _, cnt_a, st_a = g.search_ann_batch(aq, KB)
            hits = sum(len(set(ids_a[i, :TOPK].tolist()) & set(ids_e[i, :cnt_e[i]].tolist())) for i in range(BATCH))
            for _ in range(3):
                g.search_ann_batch(aq, KB)
            g.profile_read()
            reps = 30
            t0 = time.perf_counter()
            for _ in range(reps):
                g.search_ann_batch(aq, KB)
            host_ms = (time.perf_counter() - t0) * 1e3 / reps
            kms, _ = g.profile_read()
            kernel_ms = float(np.mean(kms)) if len(kms) else float("nan")
            for _ in range(3):
                g.search_ann_batch(aq[:1], KB)
            g.profile_read()
            t0 = time.perf_counter()
            for i in range(100):
                g.search_ann_batch(aq[i % BATCH:i % BATCH + 1], KB)
            b1_host = (time.perf_counter() - t0) * 1e3 / 100
            k1, _ = g.profile_read()
            nbytes = float(st_a[:, 1].sum()) * DIM * 2
            beams["beam_%d" % KB] = dict(
                recall_at_10_vs_exact=hits / float(cnt_e.sum()),
                batch256=dict(kernel_ms=kernel_ms, queries_per_s_kernel=BATCH / (kernel_ms * 1e-3), host_api_ms=host_ms,
                              queries_per_s_host_api=BATCH / (host_ms * 1e-3)),
                batch1=dict(kernel_ms=float(np.mean(k1)) if len(k1) else None, host_api_ms=b1_host),
                nodes_expanded_per_query=float(st_a[:, 0].mean()), distance_evals_per_query=float(st_a[:, 1].mean()),
                roofline=dict(bound="hbm (random 768-byte row reads, latency-bound per query)", algorithmic_bytes_per_step=nbytes,
                              achieved=nbytes / (kernel_ms * 1e-3) / 1e9, peak=peaks()[0], unit="GB/s",
                              frac=nbytes / (kernel_ms * 1e-3) / 1e9 / peaks()[0]))
        # throughput form: 8192 queries in flight (the unstaged path, ~28 query warps resident per SM)
        big = None
        try:
            NB = 8192
            pb = rng.integers(0, ann_rows, NB)
            bq = np.concatenate([g.get_vectors(int(lo), 1) for lo in pb[:64]] * (NB // 64)).astype(np.float32)
            bq += 0.02 * rng.standard_normal((NB, DIM)).astype(np.float32)
            bq = np.ascontiguousarray(bq / np.linalg.norm(bq, axis=1, keepdims=True), np.float32)
            KB = 4 * TOPK
            for _ in range(2):
                _, _, _, st_b = g.search_ann_batch(bq, KB)
            g.profile_read()
            for _ in range(5):
                g.search_ann_batch(bq, KB)
            kms, _ = g.profile_read()
            kb_ms = float(np.mean(kms))
            nb_bytes = float(st_b[:, 1].sum()) * DIM * 2
            big = dict(queries=NB, beam_k=KB, kernel_ms=kb_ms, queries_per_s_kernel=NB / (kb_ms * 1e-3),
                       roofline=dict(bound="hbm", algorithmic_bytes_per_step=nb_bytes, achieved=nb_bytes / (kb_ms * 1e-3) / 1e9,
                                     peak=peaks()[0], unit="GB/s", frac=nb_bytes / (kb_ms * 1e-3) / 1e9 / peaks()[0]))
        except Exception as ex:
            big = dict(error=repr(ex))
        g.set_profiling(False)
        # the reference's greedy_search on the host, same graph, same queries (oracle port, ONE thread), and
        # an id-for-id check of the device answers against it
        cpu_graph = None
        try:
            from oracle import oracle as o
            deg, nbr, medoid = g.get_graph()
            V = o.VamanaOracle(DIM, max_degree=32, search_list_size=75)
            V.set_graph(g.get_vectors(0, ann_rows), deg, np.where(nbr == 0xFFFFFFFF, 0, nbr), medoid)
            cpu_graph = {}
            for KB in (12 * TOPK, 4 * TOPK):
                ids_a, _, cnt_a, _ = g.search_ann_batch(aq[:64], KB)
                t0 = time.perf_counter()
                ref = [V.search(aq[i], KB)[0] for i in range(64)]
                ms = (time.perf_counter() - t0) * 1e3 / 64
                same = all(np.array_equal(ids_a[i, :cnt_a[i]], ref[i]) for i in range(64))
                cpu_graph["beam_%d" % KB] = dict(ms_per_query_1_thread=ms, ids_identical_to_device=bool(same))
            cpu_graph["kind"] = "port (oracle/vamana_oracle.c restatement of greedy_search), 64 queries, 1 thread"
            del V
        except Exception as ex:
            cpu_graph = dict(error=repr(ex))
        ann_res = dict(cpu_reference=cpu_graph, batch8192=big, rows=ann_rows, corpus="unit vectors on a %d-dimensional manifold (z A + noise), fp16" % latent,
                       max_degree=32, search_list_size=75, alpha=1.2, build_s=build_s, **beams,
                       exact_sweep_same_store_batch256_host_api_ms=exact_ms, retries=g.ann_stats()[1],
                       kernel="ann_search_kernel (one warp per query; result == the reference's greedy_search on the same graph)")
        g.close()
    except Exception as ex:  # the headline line must still print
        ann_res = dict(error=repr(ex))
    return ann_res


LEGS = {"pq": leg_pq, "spann": leg_spann, "ann": leg_ann}


def run_leg_child(name, local_rank):
    """child side of --leg: one JSON line on stdout"""
    try:
        import __graft_entry__ as ge
        ge.build()
        import shodh_memory_b200 as sg
        res = LEGS[name](sg, local_rank)
    except Exception as ex:  # noqa: BLE001
        res = dict(error=repr(ex))
    print(json.dumps(res), flush=True)
    return 0


def run_leg(name, local_rank, timeout_s=300):
    """parent side: run the leg in its own process on the same GPU"""
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--leg", name, "--leg-device", str(local_rank)],
                             capture_output=True, text=True, timeout=timeout_s)
        lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
        if not lines:
            return dict(error=f"leg {name}: no result (exit {out.returncode}): {out.stderr[-300:]}")
        return json.loads(lines[-1])
    except subprocess.TimeoutExpired:
        return dict(error=f"leg {name}: timed out after {timeout_s} s (child killed)")
    except Exception as ex:  # noqa: BLE001
        return dict(error=f"leg {name}: {ex!r}")


def run_all(sg, local_rank, path):
    res = dict(graph_index=run_leg("ann", local_rank, 600), product_quantization=run_leg("pq", local_rank),
               spann=run_leg("spann", local_rank))
    try:
        res["cpu_detail"] = cpu_detail()
        res["cpu_encoder_torch_fp32_ms"] = {"L64_2_threads": cpu_encoder_ms(64, 2), "L256_2_threads": cpu_encoder_ms(256, 2),
                                            "L64_8_threads": cpu_encoder_ms(64, min(8, os.cpu_count() or 2)),
                                            "note": "HF BertModel fp32 on this box's host, one query; the reference pads to 256 "
                                                    "and runs ONNX Runtime with 2 intra-op threads (minilm.rs:67,912-921)"}
    except Exception as ex:  # noqa: BLE001
        res["cpu_detail"] = dict(error=repr(ex))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(res, f, indent=1)
    return path
