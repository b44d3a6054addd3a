#!/usr/bin/env python
"""bench.py -- recall queries/sec on the 10M x 384 fp16 corpus (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port), rank 0 only

A "step" is one pass of the hot path over one batch of 256 synthetic queries: exact top-10
over the whole corpus (row-sharded over the N GPUs, one all-gather of per-shard top-k + merge
when N > 1).  `value` = queries/s with queries already resident in HBM; `e2e` = the same through
the host-buffer API (H2D of the queries and D2H of the results inside the timed region).
The corpus is fixed at 10M rows as N grows => "scaling": "strong".
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 23821
ROWS = int(os.environ.get("SG_BENCH_ROWS", 10_000_000))
DIM = 384
BATCH = int(os.environ.get("SG_BENCH_BATCH", 256))
TOPK = int(os.environ.get("SG_BENCH_K", 10))
B1_K = 100  # BASELINE configs[2]: batch-1 latency, top-100
METRIC = "recall_queries_per_sec_batch256_top10_exact_10Mx384_fp16"
UNIT = "queries/s"
CPU_SAMPLE_ROWS = int(os.environ.get("SG_BENCH_CPU_ROWS", 1_000_000))


def peaks():
    """(hbm GB/s, sustained dense bf16/fp16 TFLOP/s, source)"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return (float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                "measured (MEASURED_PEAKS.json: hbm_gbs = copy read+write; bf16_tflops_sustained = cuBLAS, "
                "the kernel is timed inside a long step)")
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of VamanaIndex::brute_force_search "as written"
# (per-row clone + full sort, vamana.rs:1167-1188), one query per host thread.
# ----------------------------------------------------------------------------------------------
def traffic_from_capture(kernel: str, local_rows: int):
    """DRAM bytes per launch from the committed ncu --set full capture, scaled by rows to this shard."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        c = json.load(open(p))[kernel]
    except Exception:
        return None, None
    per_row = (c["dram_bytes_read"] + c["dram_bytes_write"]) / c["rows"]
    return per_row * local_rows, (f"profiles/r01_traffic.json: ncu dram__bytes_read+write of {kernel} at {c['rows']} rows "
                                  f"({(c['dram_bytes_read'] + c['dram_bytes_write']) / c['algorithmic_bytes']:.4f}x algorithmic), "
                                  f"scaled by rows")


def cpu_detail(budget_rows: int = 100_000):
    """SURVEY 8(d) CPU baselines beside the headline one, all on a small row sample and scaled linearly to the
    full corpus: the reference loop single-threaded (it IS sequential, vamana.rs:1173) as written and with the
    incidental waste removed (heap top-k, no row clone), and the waste-free loop on all host threads."""
    from oracle import oracle as o
    threads = o.max_threads()
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


def cpu_reference(steps: int, warmup: int, budget_s: float = 90.0):
    """Times the oracle port of brute_force_search AS WRITTEN on all host threads.  The row sample is sized so
    the whole (steps + warmup) run stays within ~budget_s: scan cost is linear in rows, and the reported
    queries/s is scaled to the full corpus."""
    from oracle import oracle as o  # the ONE place bench.py executes oracle/ (cpu_baseline / --impl reference)
    threads = o.max_threads()
    rows = CPU_SAMPLE_ROWS
    total = steps + warmup
    # calibrate on 50k rows, then pick the sample size that fits the time budget
    cal_rows = min(50_000, rows)
    cal = o.synth_corpus(SEED, 0, cal_rows, DIM, f16_round=True)
    calq, _ = o.synth_queries(SEED, ROWS, threads, DIM)
    t0 = time.perf_counter()
    o.brute_force_search_batch(cal, calq, TOPK, as_written=True, threads=threads)
    per_row = (time.perf_counter() - t0) / cal_rows
    fit = int(budget_s / total / max(per_row, 1e-12))
    rows = max(20_000, min(rows, fit))
    corpus = o.synth_corpus(SEED, 0, rows, DIM, f16_round=True)
    queries, _ = o.synth_queries(SEED, ROWS, max(threads, 1) * total, DIM)
    scale = rows / ROWS
    times = []
    used = threads
    for s in range(total):
        q = queries[s * threads:(s + 1) * threads]
        t0 = time.perf_counter()
        _, _, _, used = o.brute_force_search_batch(corpus, q, TOPK, as_written=True, threads=threads)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    t = float(np.mean(times))
    qps_sample = threads / t
    return dict(value=qps_sample * scale, unit=UNIT, cores=int(used), kind="port",
                sample=(f"{threads} queries/step (one per host thread) x {rows} rows of the {ROWS}-row corpus, "
                        f"top-{TOPK}, reference as written (row clone + full sort); qps scaled by {scale:g} "
                        f"(scan cost is linear in rows)"),
                ms_per_step=t * 1e3)


# ----------------------------------------------------------------------------------------------
# Optional legs that run in a CHILD process (`bench.py --leg NAME`): a fault or a hang inside one of them
# cannot take the headline line down with it (the parent waits with a timeout and records the outcome).
# ----------------------------------------------------------------------------------------------
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
    sweeps = (BATCH + 1) // 2
    nbytes = float(pq_rows) * (DIM // 8 + 4)
    pq_res = dict(rows=pq_rows, code_bytes=DIM // 8, train_s_5000x20=train_s, encode_add_rows_per_s=pq_rows / t_add,
                  batch256=dict(host_api_ms=ms, queries_per_s=BATCH / (ms * 1e-3), sweeps=sweeps,
                                roofline=dict(bound="hbm (code stream) / shared-memory gathers", algorithmic_bytes_per_sweep=nbytes,
                                              achieved=nbytes * sweeps / (ms * 1e-3) / 1e9, peak=peaks()[0], unit="GB/s",
                                              frac=nbytes * sweeps / (ms * 1e-3) / 1e9 / peaks()[0],
                                              note="wall clock of sg_pq_search incl. H2D/D2H and table build; 2 queries per sweep")),
                  batch1_host_api_ms=ms1,
                  cpu_reference=dict(ms_per_query_1_thread=cpu_ms, kind="port (oracle/pq_oracle.c), 100k-vector slice scaled linearly"),
                  kernel="pq_scan_kernel (48 table look-ups per vector in the reference's order; ids and distances bit-identical to the CPU path)")
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


LEGS = {"pq": leg_pq, "spann": leg_spann}


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


def run_leg(name, local_rank, timeout_s=150):
    """parent side: run the leg in its own process on the same GPU"""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--leg", name, "--leg-device", str(local_rank)],
                             capture_output=True, text=True, timeout=timeout_s)
        lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
        if not lines:
            return dict(error=f"leg {name}: no result (exit {out.returncode}): {out.stderr[-300:]}")
        return json.loads(lines[-1])
    except subprocess.TimeoutExpired:
        return dict(error=f"leg {name}: timed out after {timeout_s} s (child killed)")
    except Exception as ex:  # noqa: BLE001
        return dict(error=f"leg {name}: {ex!r}")


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid: str | None):
        self.proc, self.path = None, os.path.join("/tmp", f"sg_clocks_{os.getpid()}.csv")
        cmd = ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"]
        if uuid:
            cmd += ["-i", uuid]
        try:
            self.proc = subprocess.Popen(cmd, stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        # "under load": samples at or above the median power draw
        med_p = float(np.median(pw))
        load = [c for c, p in zip(sm, pw) if p >= med_p] or sm
        return dict(sm_mhz=float(np.median(load)), sm_max_mhz=float(max(smax)), power_w_max=float(max(pw)),
                    samples=len(sm), reasons=sorted(reasons))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--leg", default=None, choices=sorted(LEGS))          # internal: run one optional leg and print its JSON
    ap.add_argument("--leg-device", type=int, default=0)
    args = ap.parse_args()
    if args.leg:
        return run_leg_child(args.leg, args.leg_device)
    steps, warmup = max(args.steps, 1), max(args.warmup, 3 if args.impl == "b200" else 1)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference(steps, warmup)
        line = dict(metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warmup,
                    ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="strong", vs_baseline=None,
                    dtype="f32", data="synthetic", impl="reference",
                    config=dict(workload=f"{ROWS}x{DIM} corpus (f16-rounded values held as f32, like the reference's "
                                         f"Vec<Vec<f32>>), exact top-{TOPK}", rows=ROWS, dim=DIM, k=TOPK),
                    cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"]),
                    e2e=dict(value=r["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0)
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    else:
        torch.cuda.set_device(0)
        local_rank = 0
    import shodh_memory_b200 as sg
    dev = torch.device("cuda", local_rank)

    store = sg.ShardedGpuStore(ROWS, DIM, "f16", rank, world, local_rank, exchange=os.environ.get("SG_BENCH_EXCHANGE", "auto"))
    store.fill_synthetic(SEED)
    local_rows = store.hi - store.lo

    # synthetic queries: half fresh unit vectors, half normalize(row_j + noise) for rows of rank 0's shard
    q_host = torch.empty((BATCH, DIM), dtype=torch.float32).pin_memory()
    if rank == 0:
        rng = np.random.default_rng(SEED)
        q = rng.standard_normal((BATCH, DIM)).astype(np.float32)
        picks = rng.integers(0, local_rows, BATCH // 2)
        for i, j in enumerate(picks):
            q[2 * i + 1] = store.local.get_vectors(int(j), 1)[0] + 0.003 * q[2 * i + 1]
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        q_host.copy_(torch.from_numpy(q))
    dq = q_host.to(dev)
    if world > 1:
        dist.broadcast(dq, 0)
        q_host.copy_(dq.cpu())
    q_np = q_host.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    tc_peak_early = peaks()[1]
    sampler = None
    if rank == 0:
        try:
            uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
        except Exception:
            uuid = None
        sampler = ClockSampler(uuid)

    # ---------------- device-resident region: `value` ----------------
    store.local.set_profiling(True)
    for _ in range(warmup):
        store.search_device(dq, TOPK)
    barrier()
    store.local.profile_read()
    launches0 = sg.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]   # per-step marks: median / p95
    e0.record()
    for i in range(steps):
        store.search_device(dq, TOPK)
        marks[i].record()
    e1.record()
    barrier()
    launches = sg.kernel_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    per_step = np.array([(marks[i - 1] if i else e0).elapsed_time(marks[i]) for i in range(steps)])
    scan_ms, sweeps_per_step = store.local.profile_read()
    ms_per_step = ms_total / steps
    value = BATCH * steps / (ms_total / 1e3)

    # ---------------- end-to-end region: host buffers in, host results out ----------------
    def e2e_step():
        if world == 1:
            return store.local.search_batch(q_np, TOPK)      # sg_store_search: the C-ABI call a user makes
        return store.search(q_host, TOPK)
    for _ in range(warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        e2e_step()
    e1.record()
    barrier()
    if world == 1:
        # sg_store_search is synchronous (returns after its D2H completes) and runs on the store's own
        # stream, which torch events on the current stream cannot see: time it with the host clock.
        e2e_ms = (time.perf_counter() - t0) * 1e3
    else:
        e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = BATCH * steps / (e2e_ms / 1e3)
    store.local.profile_read()

    # ---------------- batch-1 latency (configs[2]): top-100, device-resident ----------------
    b1_steps = max(steps, 200)
    for _ in range(warmup):
        store.search_device(dq[:1], B1_K)
    barrier()
    store.local.profile_read()
    e0.record()
    for i in range(b1_steps):
        store.search_device(dq[i % BATCH:i % BATCH + 1], B1_K)
    e1.record()
    barrier()
    b1_ms = max_over_ranks(e0.elapsed_time(e1)) / b1_steps
    b1_scan, _ = store.local.profile_read()
    # ---------------- configs[2] whole: batch-1 top-100 + relevance fusion through the host C ABI ----------------
    fused_ms = None
    if rank == 0 and world == 1:
        rng = np.random.default_rng(SEED + 2)
        now = 1_790_000_000
        n_c = B1_K
        side = dict(entity=np.where(rng.random(n_c) < 0.3, rng.random(n_c), 0.0).astype(np.float32),
                    tag=np.where(rng.random(n_c) < 0.3, rng.random(n_c), 0.0).astype(np.float32),
                    importance=rng.random(n_c).astype(np.float32), momentum=rng.uniform(-1, 1, n_c).astype(np.float32),
                    access_count=np.minimum(rng.geometric(0.3, n_c), 64).astype(np.uint32),
                    graph=rng.random(n_c).astype(np.float32),
                    created_at=(now - rng.integers(0, 72 * 3600, n_c)).astype(np.int64))

        def recall_fused(i):
            _, dist, _ = store.local.search_batch(q_np[i % BATCH:i % BATCH + 1], B1_K)     # sg_store_search
            return sg.fuse_topk(dict(side, semantic=-dist[0]), now, max_results=10)        # sg_fuse_topk
        for i in range(5):
            recall_fused(i)
        t0 = time.perf_counter()
        for i in range(100):
            recall_fused(i)
        fused_ms = (time.perf_counter() - t0) * 1e3 / 100
        store.local.profile_read()
    # ---------------- encoder (configs[4]): MiniLM-L6, random-init weights, 64-token queries ----------------
    enc_res = None
    if rank == 0 and os.environ.get("SG_BENCH_ENCODER", "1") == "1":
        import tempfile
        wpath = os.path.join(tempfile.gettempdir(), f"sg_minilm_random_{os.getpid()}.sgw")
        sg.export_bert_weights(sg.random_minilm_state_dict(SEED), wpath, 6, 384, 12, 1536, 30522, 512)
        enc = sg.GpuMiniLM(wpath, local_rank)
        os.remove(wpath)
        LQ = 64
        rng = np.random.default_rng(SEED + 1)
        ids_np = rng.integers(1000, 30522, (BATCH, LQ)).astype(np.int64)
        ids_np[:, 0], ids_np[:, -1] = 101, 102
        d_ids = torch.from_numpy(ids_np).to(dev)
        d_mask = torch.ones((BATCH, LQ), dtype=torch.int64, device=dev)
        d_emb = torch.empty((BATCH, DIM), dtype=torch.float32, device=dev)
        cs = torch.cuda.current_stream().cuda_stream

        def time_enc(b, reps):
            for _ in range(3):
                enc.encode_ids_device(d_ids.data_ptr(), d_mask.data_ptr(), b, LQ, d_emb.data_ptr(), cs)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                enc.encode_ids_device(d_ids.data_ptr(), d_mask.data_ptr(), b, LQ, d_emb.data_ptr(), cs)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        ms1, msb = time_enc(1, 100), time_enc(BATCH, 20)
        flops_tok = 6 * 2 * (4 * 384 * 384 + 2 * 384 * 1536) + 6 * 4 * LQ * 384
        tf = flops_tok * BATCH * LQ / (msb * 1e-3) / 1e12
        # end-to-end recall for one query: host ids -> H2D -> encoder -> exact scan of this rank's shard -> D2H
        h_ids = torch.from_numpy(ids_np[:1]).pin_memory()
        h_mask = torch.ones((1, LQ), dtype=torch.int64).pin_memory()
        out_i = torch.empty((1, TOPK), dtype=torch.int32, device=dev)
        out_d = torch.empty((1, TOPK), dtype=torch.float32, device=dev)
        out_c = torch.empty(1, dtype=torch.int32, device=dev)
        h_out = torch.empty((1, TOPK), dtype=torch.int32).pin_memory()

        def recall_once():
            d_ids[:1].copy_(h_ids, non_blocking=True)
            d_mask[:1].copy_(h_mask, non_blocking=True)
            enc.encode_ids_device(d_ids.data_ptr(), d_mask.data_ptr(), 1, LQ, d_emb.data_ptr(), cs)
            store.local.search_device(d_emb.data_ptr(), 1, TOPK, out_i.data_ptr(), out_d.data_ptr(), out_c.data_ptr(), cs)
            h_out.copy_(out_i, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        for _ in range(3):
            recall_once()
        t0 = time.perf_counter()
        for _ in range(50):
            recall_once()
        rec_ms = (time.perf_counter() - t0) * 1e3 / 50
        enc_res = dict(model="MiniLM-L6 architecture (6x384, 12 heads, FFN 1536), random-init weights, fp16 tensor-core GEMMs",
                       tokens_per_query=LQ, encode_ms_batch1=ms1, encode_ms_batch256=msb,
                       queries_per_s_batch256=BATCH / (msb * 1e-3),
                       tensor=dict(achieved=tf, peak=tc_peak_early, unit="TFLOP/s", frac=tf / tc_peak_early),
                       recall_e2e_batch1=dict(ms_per_query=rec_ms, rows=int(local_rows),
                                              note="pinned host ids -> H2D -> encoder -> exact top-%d scan of this "
                                                   "rank's shard -> D2H, one stream" % TOPK))
        enc.close()
    tc_stats_val = store.local.tc_stats() if rank == 0 else (0, 0)   # read before the optional legs below
    # ---------------- graph index: VamanaIndex::search's default ANN arm on the device ----------------
    # Opt-out with SG_BENCH_ANN=0.  A separate low-intrinsic-dimension corpus (i.i.d. unit vectors have no neighbourhood
    # structure for a graph to follow): device build (R 32, L 75, alpha 1.2), then batch-256 and batch-1 beam searches with the
    # reference's k = 4 x limit = 40 for top-10 (retrieval.rs:915-916), recall@10 against the exact sweep of the same store.
    ann_res = None
    if rank == 0 and world == 1 and os.environ.get("SG_BENCH_ANN", "1") == "1":
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
    # ---------------- Product Quantization and SPANN: each in its own child process ----------------
    pq_res = spann_res = None
    if rank == 0 and world == 1 and os.environ.get("SG_BENCH_PQ", "1") == "1":
        pq_res = run_leg("pq", local_rank)
    if rank == 0 and world == 1 and os.environ.get("SG_BENCH_SPANN", "1") == "1":
        spann_res = run_leg("spann", local_rank)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, tc_peak, peak_src = peaks()
    alg_bytes = local_rows * DIM * 2                       # one sweep of this rank's shard serves the whole batch
    scan_step_ms = float(np.mean(scan_ms)) if len(scan_ms) else float("nan")
    achieved = alg_bytes / (scan_step_ms * 1e-3) / 1e9
    b1_scan_ms = float(np.mean(b1_scan)) if len(b1_scan) else float("nan")
    b1_achieved = alg_bytes / (b1_scan_ms * 1e-3) / 1e9

    cpu = cpu_reference(2, 1, budget_s=12.0) if world == 1 else None
    cpu_more = None
    if world == 1 and os.environ.get("SG_BENCH_CPU_DETAIL", "1") == "1":
        cpu_more = cpu_detail()
        if enc_res is not None:
            enc_res["cpu_torch_fp32_ms"] = {"L64_2_threads": cpu_encoder_ms(64, 2), "L256_2_threads": cpu_encoder_ms(256, 2),
                                            "L64_8_threads": cpu_encoder_ms(64, min(8, os.cpu_count() or 2)),
                                            "note": "HF BertModel fp32 on this box's host, one query; the reference pads to 256 "
                                                    "and runs ONNX Runtime with 2 intra-op threads (minilm.rs:67,912-921)"}

    # Which roofline binds the sweep?  One sweep must move alg_bytes from HBM AND (tensor path) do
    # 2*rows*dim*queries_per_sweep FLOP on the tensor pipe; at 256 queries per sweep the tensor time exceeds
    # the HBM time, so that is the bound reported (the HBM figure is kept alongside).
    tensor_path = BATCH >= 5 and TOPK <= 24
    q_per_sweep = (256 if BATCH > 128 else 128) if tensor_path else 0
    flops_step = 2.0 * local_rows * DIM * q_per_sweep * max(int(sweeps_per_step), 1)
    t_hbm = alg_bytes / (peak * 1e9)
    t_tc = flops_step / (tc_peak * 1e12) if tensor_path else 0.0
    hbm_view = dict(achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak)
    if tensor_path and t_tc > t_hbm:
        tf = flops_step / (scan_step_ms * 1e-3) / 1e12
        traffic, traffic_src = traffic_from_capture("scan_tc2_kernel", local_rows)
        roofline = dict(bound="tensor", achieved=tf, peak=tc_peak, unit="TFLOP/s", frac=tf / tc_peak, traffic=traffic,
                        traffic_source=traffic_src,
                        kernel="scan_tc2_kernel (tcgen05.mma.cta_group::2, 256 queries x 256 rows per tile)",
                        flops_per_step=flops_step, hbm=hbm_view)
    else:
        traffic, traffic_src = (None, None) if tensor_path else traffic_from_capture("scan_exact_kernel", local_rows)
        roofline = dict(bound="hbm", traffic=traffic, traffic_source=traffic_src,
                        kernel=("scan_tc_kernel (tcgen05 sweep)" if tensor_path else "scan_exact_kernel"), **hbm_view)
    roofline.update(algorithmic_bytes_per_step=alg_bytes, scan_ms_per_step=scan_step_ms,
                    sweep_launches_per_step=int(sweeps_per_step), peak_source=peak_src,
                    note="achieved = work ONE sweep of the shard must do / CUDA-event time of all sweep launches "
                         "(pre-pass included) of the step; padded query rows (zero queries up to 128/256) count as work")

    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=steps, warmup=warmup, ms_per_step=ms_per_step,
        ms_per_step_median=float(np.median(per_step)), ms_per_step_p95=float(np.percentile(per_step, 95)),
        higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload=f"{ROWS}x{DIM} fp16 corpus resident in HBM, batch-{BATCH} f32 queries, exact top-{TOPK} "
                             f"(dist=-dot in the reference's summation order; ids bit-identical to the CPU path)",
                    rows=ROWS, dim=DIM, batch=BATCH, k=TOPK, corpus_dtype="f16",
                    parallelism=f"row-shard x{world}" + (
                        "" if world == 1 else (", fused P2P push + merge kernel over NVLink peer memory" if store.exchange_mode == "p2p"
                                                else ", 1 NCCL all-gather of per-shard top-k + merge kernel")),
                    l2="per-GPU shard %.2f GB >> 126 MB L2: every sweep streams from HBM, no flush needed" % (alg_bytes / 1e9)),
        roofline=roofline,
        e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=BATCH * DIM * 4,
                 d2h_bytes_per_step=BATCH * TOPK * 8 + BATCH * 4, ms_per_step=e2e_ms / steps),
        batch1=dict(value=1e3 / b1_ms, unit=UNIT, k=B1_K, ms_per_query=b1_ms,
                    with_fusion_host_api_ms=fused_ms,   # configs[2]: sg_store_search(top-100) + sg_fuse_topk, host buffers, wall clock
                    roofline=dict(bound="hbm", achieved=b1_achieved, peak=peak, unit="GB/s", frac=b1_achieved / peak,
                                  scan_ms=b1_scan_ms, kernel="scan_exact_kernel",
                                  note="peak = measured COPY bandwidth (read+write); a read-only sweep can slightly exceed it "
                                       "(ncu: 80 % of the 8.2 TB/s DRAM peak, profiles/r01_scan_exact_ncu.md)",
                                  traffic=traffic_from_capture("scan_exact_kernel", local_rows)[0])),
        tensor_path=dict(zip(("queries", "exact_reruns"), tc_stats_val)),
        encoder=enc_res,
        graph_index=ann_res,
        product_quantization=pq_res,
        spann=spann_res,
        gpu_launches=int(launches), clocks=clocks)
    if cpu:
        line["cpu_baseline"] = dict(value=cpu["value"], unit=UNIT, cores=cpu["cores"], kind=cpu["kind"], sample=cpu["sample"])
    if cpu_more:
        line["cpu_baseline"]["detail"] = cpu_more
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
