"""Host-side mirror of the reference's vector-store surface over the C ABI.

`GpuVectorStore` exposes the method set `RetrievalEngine` uses on `VamanaIndex`
(reference src/vector_db/vamana.rs: search :764, add_vector :853, build :200,
mark_deleted :813, is_deleted :823, deleted_count :828, len :976,
extract_all_vectors) with the same argument meaning and error behaviour, so the parity
tests read like the reference's own.  All compute happens in libshodh_gpu.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

_DT = {"f16": capi.SG_DTYPE_F16, "f32": capi.SG_DTYPE_F32}
# DistanceMetric (vamana.rs:38-54) + the similarity.rs ranking sg_top_k_similar uses
_METRIC = {"neg_dot": capi.SG_METRIC_NEG_DOT, "euclidean": capi.SG_METRIC_EUCLIDEAN, "cosine": capi.SG_METRIC_COSINE,
           "cosine_sim": capi.SG_METRIC_COSINE_SIM}


def _ptr(a: np.ndarray, ty):
    return a.ctypes.data_as(C.POINTER(ty))


class GpuVectorStore:
    def __init__(self, dim: int = 384, dtype: str = "f16", device: int = 0, metric: str = "neg_dot"):
        self._h = C.c_void_p()
        self.dim = dim
        self.dtype = dtype
        self.device = device
        self.metric = metric
        capi.check(capi.lib().sg_store_create(C.byref(self._h), device, dim, _DT[dtype], _METRIC[metric]))

    # -- lifecycle --------------------------------------------------------------------
    def close(self) -> None:
        if self._h:
            capi.lib().sg_store_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # -- VamanaIndex surface ------------------------------------------------------------
    def __len__(self) -> int:
        n = C.c_uint64()
        capi.check(capi.lib().sg_store_len(self._h, C.byref(n)))
        return n.value

    def reserve(self, n_rows: int) -> None:
        capi.check(capi.lib().sg_store_reserve(self._h, n_rows))

    def memory_info(self) -> dict:
        """How the row matrix is held: in-place growth (VMM chunks) or the allocate-and-copy fallback."""
        ip, cap, nb, ch = C.c_int(), C.c_uint64(), C.c_uint64(), C.c_uint32()
        capi.check(capi.lib().sg_store_memory_info(self._h, C.byref(ip), C.byref(cap), C.byref(nb), C.byref(ch)))
        return {"in_place": bool(ip.value), "capacity_rows": cap.value, "bytes": nb.value, "chunks": ch.value}

    def add_vector(self, vector) -> int:
        """VamanaIndex::add_vector -> the new id."""
        return self.add_vectors(np.asarray(vector, np.float32)[None, :])

    def add_vectors(self, rows) -> int:
        """VamanaIndex::build / rebuild_from_vectors (bulk) -> id of the first appended row."""
        rows = np.ascontiguousarray(rows, np.float32)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise capi.SgError(capi.SG_ERR_DIM, f"expected [n,{self.dim}] rows, got {rows.shape}")
        first = C.c_uint32()
        capi.check(capi.lib().sg_store_append(self._h, _ptr(rows, C.c_float), rows.shape[0], C.byref(first)))
        return first.value

    build = add_vectors

    def add_vectors_device(self, d_rows: int, n: int, stream: int = 0) -> int:
        """rows already on the device ([n,dim] f32, raw pointer) -> id of the first appended row"""
        first = C.c_uint32()
        capi.check(capi.lib().sg_store_append_device(self._h, d_rows, n, C.byref(first), stream))
        return first.value

    def append_synthetic(self, seed: int, first_row: int, n: int) -> None:
        capi.check(capi.lib().sg_store_append_synthetic(self._h, seed, first_row, n))

    def mark_deleted(self, vector_id: int) -> bool:
        ok = C.c_int()
        capi.check(capi.lib().sg_store_mark_deleted(self._h, vector_id, C.byref(ok)))
        return bool(ok.value)

    def is_deleted(self, vector_id: int) -> bool:
        d = C.c_int()
        capi.check(capi.lib().sg_store_is_deleted(self._h, vector_id, C.byref(d)))
        return bool(d.value)

    def clear_deleted(self) -> None:
        capi.check(capi.lib().sg_store_clear_deleted(self._h))

    def deleted_count(self) -> int:
        n = C.c_uint64()
        capi.check(capi.lib().sg_store_deleted_count(self._h, C.byref(n)))
        return n.value

    def get_vectors(self, first: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dim), np.float32)
        capi.check(capi.lib().sg_store_get_vectors(self._h, first, n, _ptr(out, C.c_float)))
        return out

    def extract_all_vectors(self) -> np.ndarray:
        return self.get_vectors(0, len(self))

    def load_from_file(self, path: str, verify_checksum: bool = True) -> None:
        """VamanaIndex::load_from_file: the reference's "VAMA" v1 file -> this (empty) store."""
        capi.check(capi.lib().sg_store_load_vama(self._h, path.encode(), int(verify_checksum)))

    def save_to_file(self, path: str) -> None:
        """VamanaIndex::save_to_file: rows, deleted ids, graph + medoid -> a "VAMA" v1 file."""
        capi.check(capi.lib().sg_store_save_vama(self._h, path.encode()))

    def set_order(self, order: str) -> None:
        capi.check(capi.lib().sg_store_set_order(
            self._h, capi.SG_ORDER_AVX2 if order == "avx2" else capi.SG_ORDER_SCALAR))

    def set_id_base(self, base: int) -> None:
        capi.check(capi.lib().sg_store_set_id_base(self._h, base))

    def search(self, query, k: int):
        """VamanaIndex::search(query, k) -> list[(id, distance)], distance = -dot, ascending."""
        ids, dist, counts = self.search_batch(np.asarray(query, np.float32)[None, :], k)
        c = int(counts[0])
        return [(int(i), float(d)) for i, d in zip(ids[0, :c], dist[0, :c])]

    def search_batch(self, queries, k: int):
        """b queries at once -> (ids [b,k] u32, dist [b,k] f32, counts [b] u32)."""
        q = np.ascontiguousarray(queries, np.float32)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise capi.SgError(capi.SG_ERR_DIM, f"expected [b,{self.dim}] queries, got {q.shape}")
        b = q.shape[0]
        ids = np.empty((b, k), np.uint32)
        dist = np.empty((b, k), np.float32)
        counts = np.zeros(b, np.uint32)
        capi.check(capi.lib().sg_store_search(self._h, _ptr(q, C.c_float), b, k, _ptr(ids, C.c_uint32),
                                              _ptr(dist, C.c_float), _ptr(counts, C.c_uint32)))
        return ids, dist, counts

    # -- Vamana graph index (VamanaIndex::build / search's ANN arm / add_vector) -----------------------
    def enable_graph(self, max_degree: int = 32, search_list_size: int = 75, alpha: float = 1.2) -> None:
        """An empty store links every appended row like VamanaIndex::add_vector from now on."""
        capi.check(capi.lib().sg_store_enable_graph(self._h, max_degree, search_list_size, alpha))

    def build_graph(self, max_degree: int = 32, search_list_size: int = 75, alpha: float = 1.2, seed: int = 0) -> None:
        """VamanaIndex::build over the rows in the store (VamanaConfig::default: R 32, L 75, alpha 1.2)."""
        capi.check(capi.lib().sg_store_build_graph(self._h, max_degree, search_list_size, alpha, seed))

    def set_graph(self, deg, nbr, medoid: int, max_degree: int) -> None:
        """Upload an adjacency built elsewhere (deg [n] u32, nbr [n, stride] u32)."""
        deg = np.ascontiguousarray(deg, np.uint32)
        nbr = np.ascontiguousarray(nbr, np.uint32)
        capi.check(capi.lib().sg_store_set_graph(self._h, max_degree, medoid, _ptr(deg, C.c_uint32), _ptr(nbr, C.c_uint32),
                                                 nbr.shape[1] if nbr.ndim == 2 else max_degree, deg.size))

    def graph_info(self) -> dict:
        r, m, n = C.c_uint32(), C.c_uint32(), C.c_uint64()
        capi.check(capi.lib().sg_store_graph_info(self._h, C.byref(r), C.byref(m), C.byref(n)))
        return dict(max_degree=r.value, medoid=m.value, graph_len=n.value)

    def get_graph(self):
        """-> (deg [n] u32, nbr [n, max_degree] u32 with 0xFFFFFFFF padding, medoid)"""
        info = self.graph_info()
        n, r = info["graph_len"], info["max_degree"]
        deg = np.zeros(n, np.uint32)
        nbr = np.zeros((n, max(r, 1)), np.uint32)
        capi.check(capi.lib().sg_store_get_graph(self._h, _ptr(deg, C.c_uint32), _ptr(nbr, C.c_uint32), n))
        return deg, nbr, info["medoid"]

    def search_ann(self, query, k: int):
        """VamanaIndex::search, ANN arm -> list[(id, distance)] ascending."""
        ids, dist, counts, _ = self.search_ann_batch(np.asarray(query, np.float32)[None, :], k)
        c = int(counts[0])
        return [(int(i), float(d)) for i, d in zip(ids[0, :c], dist[0, :c])]

    def search_ann_batch(self, queries, k: int):
        """-> (ids [b,k], dist [b,k], counts [b], stats [b,2] = nodes expanded, distance evaluations)."""
        q = np.ascontiguousarray(queries, np.float32)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise capi.SgError(capi.SG_ERR_DIM, f"expected [b,{self.dim}] queries, got {q.shape}")
        b = q.shape[0]
        ids = np.empty((b, k), np.uint32)
        dist = np.empty((b, k), np.float32)
        counts = np.zeros(b, np.uint32)
        stats = np.zeros((b, 2), np.uint64)
        capi.check(capi.lib().sg_store_search_ann(self._h, _ptr(q, C.c_float), b, k, _ptr(ids, C.c_uint32),
                                                  _ptr(dist, C.c_float), _ptr(counts, C.c_uint32), _ptr(stats, C.c_uint64)))
        return ids, dist, counts, stats

    def search_ann_device(self, d_queries: int, b: int, k: int, d_ids: int, d_dist: int, d_counts: int, d_flags: int = 0,
                          stream: int = 0) -> None:
        capi.check(capi.lib().sg_store_search_ann_device(self._h, d_queries, b, k, d_ids, d_dist, d_counts, d_flags, stream))

    def ann_phase_cycles(self, enable: bool = True) -> np.ndarray:
        """profiling hook -> cycles in (pop, list read + visited set, row fetch, scoring, insertion) since the last call"""
        out = np.zeros(8, np.uint64)
        capi.check(capi.lib().sg_store_ann_phase_cycles(self._h, int(enable), _ptr(out, C.c_uint64)))
        return out[:5]

    def ann_stats(self):
        """-> (queries served by the graph search, re-runs with a larger visited table)"""
        a, b = C.c_uint64(), C.c_uint64()
        capi.check(capi.lib().sg_store_ann_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_memory_ids(self, vec2mem) -> None:
        """IdMapping (retrieval.rs:64-140): vec2mem [n,2] u64 = MemoryId (uuid hi, lo) of every vector id."""
        t = np.ascontiguousarray(vec2mem, np.uint64)
        capi.check(capi.lib().sg_store_set_memory_ids(self._h, _ptr(t, C.c_uint64), t.shape[0]))

    def search_ids(self, queries, limit: int):
        """RetrievalEngine::search_ids after the embed step -> (mem ids [b,limit,2] u64, sims [b,limit], counts [b])."""
        q = np.ascontiguousarray(queries, np.float32)
        b = q.shape[0]
        mem = np.empty((b, limit, 2), np.uint64)
        sim = np.empty((b, limit), np.float32)
        cnt = np.zeros(b, np.uint32)
        capi.check(capi.lib().sg_store_search_memory_ids(self._h, _ptr(q, C.c_float), b, limit, _ptr(mem, C.c_uint64),
                                                         _ptr(sim, C.c_float), _ptr(cnt, C.c_uint32)))
        return mem, sim, cnt

    def search_device(self, d_queries: int, b: int, k: int, d_ids: int, d_dist: int, d_counts: int,
                      stream: int = 0) -> None:
        """Raw device pointers (ints), asynchronous on `stream` (a cudaStream_t value)."""
        capi.check(capi.lib().sg_store_search_device(self._h, d_queries, b, k, d_ids, d_dist, d_counts, stream))

    def search_sharded_device(self, d_queries: int, b: int, k: int, peer_ptrs, rank: int, world: int, epoch: int,
                              d_ids: int, d_dist: int, d_counts: int, stream: int = 0) -> None:
        """One rank's whole multi-GPU recall step: local search with the push folded into the result kernels,
        then the wait + merge kernel (sg_store_search_sharded_device).  Outputs hold the GLOBAL top-k."""
        capi.check(capi.lib().sg_store_search_sharded_device(self._h, d_queries, b, k, peer_ptrs, rank, world, epoch,
                                                             d_ids, d_dist, d_counts, stream))

    def set_tc_mode(self, mode: str) -> None:
        """"auto": pick the sweep by shape; "exact": always path A (the comparator of the parity checks)."""
        capi.check(capi.lib().sg_store_set_tc_mode(self._h, {"auto": 0, "exact": 1}[mode]))

    def debug_scores(self, queries, first_row: int, n_rows: int):
        """diagnostics -> (tensor-core scores [b,n_rows], exact reference-order dots [b,n_rows])"""
        q = np.ascontiguousarray(queries, np.float32)
        b = q.shape[0]
        tc = np.empty((n_rows, b), np.float32)          # the tensor-core dump is row-major by corpus row
        ex = np.empty((b, n_rows), np.float32)
        capi.check(capi.lib().sg_debug_tc_scores(self._h, _ptr(q, C.c_float), b, first_row, n_rows, _ptr(tc, C.c_float)))
        capi.check(capi.lib().sg_debug_exact_scores(self._h, _ptr(q, C.c_float), b, first_row, n_rows, _ptr(ex, C.c_float)))
        return tc.T, ex

    # -- profiling hooks for bench.py -----------------------------------------------------
    def set_profiling(self, on: bool) -> None:
        capi.check(capi.lib().sg_store_set_profiling(self._h, int(on)))

    def tc_stats(self):
        """-> (queries served by the tensor-core path, queries that needed the exact re-run)"""
        a, b = C.c_uint64(), C.c_uint64()
        capi.check(capi.lib().sg_store_tc_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def profile_read(self, cap: int = 4096):
        """-> (per-search scan ms array, sweep launches per search)"""
        ms = np.zeros(cap, np.float32)
        n, l = C.c_uint32(), C.c_uint32()
        capi.check(capi.lib().sg_store_profile_read(self._h, _ptr(ms, C.c_float), cap, C.byref(n), C.byref(l)))
        return ms[:min(n.value, cap)].copy(), l.value

    def last_scan_ms(self):
        ms, n = C.c_float(), C.c_uint32()
        capi.check(capi.lib().sg_store_last_scan_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


def top_k_similar(query, candidates, k: int, order: str = "scalar", device: int = 0):
    """similarity.rs::top_k_similar(query, candidates, k) -> (positions [r] u32, scores [r] f32), r = min(k, n)."""
    q = np.ascontiguousarray(query, np.float32).ravel()
    c = np.ascontiguousarray(candidates, np.float32)
    if c.ndim != 2:
        raise capi.SgError(capi.SG_ERR_DIM, f"expected [n,dim] candidates, got {c.shape}")
    n, dim = c.shape
    pos = np.empty(max(k, 1), np.uint32)
    sc = np.empty(max(k, 1), np.float32)
    r = C.c_uint32()
    capi.check(capi.lib().sg_top_k_similar(device, _ptr(q, C.c_float), q.size, _ptr(c, C.c_float), n, dim, k,
                                           capi.SG_ORDER_AVX2 if order == "avx2" else capi.SG_ORDER_SCALAR,
                                           _ptr(pos, C.c_uint32), _ptr(sc, C.c_float), C.byref(r)))
    return pos[:r.value].copy(), sc[:r.value].copy()


def cosine_similarity(a, b, order: str = "scalar", device: int = 0) -> float:
    """similarity.rs::cosine_similarity(a, b) evaluated on the device (one candidate)."""
    _, sc = top_k_similar(a, np.asarray(b, np.float32)[None, :], 1, order, device)
    return float(sc[0])


def vama_inspect(path: str, verify_checksum: bool = True) -> dict:
    n, off = C.c_uint64(), C.c_uint64()
    dim, dc, metric = C.c_uint32(), C.c_uint32(), C.c_uint32()
    capi.check(capi.lib().sg_vama_inspect(path.encode(), int(verify_checksum), C.byref(n), C.byref(dim), C.byref(dc),
                                          C.byref(metric), C.byref(off)))
    return dict(num_vectors=n.value, dimension=dim.value, deleted_count=dc.value, metric=metric.value,
                vectors_offset=off.value)


def merge_topk_device(device: int, d_ids: int, d_dist: int, b: int, n_lists: int, k: int, d_ids_out: int,
                      d_dist_out: int, d_counts_out: int, stream: int = 0, list_stride: int | None = None,
                      query_stride: int | None = None) -> None:
    """Element (q, list l, slot j) at l*list_stride + q*query_stride + j; default layout [b, n_lists, k]."""
    if list_stride is None:
        list_stride, query_stride = k, n_lists * k
    capi.check(capi.lib().sg_merge_topk_device_strided(device, d_ids, d_dist, b, n_lists, k, list_stride,
                                                       query_stride, d_ids_out, d_dist_out, d_counts_out, stream))
