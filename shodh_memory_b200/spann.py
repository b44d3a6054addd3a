"""Host-side mirror of the reference's SpannIndex (src/vector_db/spann.rs) over the C ABI (sg_spann_*).
All compute happens in libshodh_gpu.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

_METRIC = {"neg_dot": 0, "euclidean": 1, "cosine": 2}


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


class GpuSpannIndex:
    def __init__(self, handle, dim: int):
        self._h, self.dim = handle, dim

    @classmethod
    def build(cls, vectors, order, pq_orders, num_partitions: int = 0, num_probes: int = 10, kmeans_iterations: int = 25,
              metric: str = "neg_dot", pq_iterations: int = 20, device: int = 0) -> "GpuSpannIndex":
        """SpannIndex::build; order [n] / pq_orders [dim/8, n] = the shuffles of the partition and PQ k-means."""
        v = np.ascontiguousarray(vectors, np.float32)
        o = np.ascontiguousarray(order, np.uint32)
        po = np.ascontiguousarray(pq_orders, np.uint32)
        n, dim = v.shape
        h = C.c_void_p()
        capi.check(capi.lib().sg_spann_build(C.byref(h), device, dim, num_partitions, num_probes, kmeans_iterations,
                                             _METRIC[metric], _ptr(v, C.c_float), n, _ptr(o, C.c_uint32), pq_iterations,
                                             _ptr(po, C.c_uint32)))
        return cls(h, dim)

    @classmethod
    def from_arrays(cls, centroids, pq_centroids, part_off, ids, codes, num_probes: int = 10, metric: str = "neg_dot",
                    device: int = 0) -> "GpuSpannIndex":
        c = np.ascontiguousarray(centroids, np.float32)
        pc = np.ascontiguousarray(pq_centroids, np.float32)
        off = np.ascontiguousarray(part_off, np.uint64)
        i = np.ascontiguousarray(ids, np.uint32)
        cd = np.ascontiguousarray(codes, np.uint8)
        h = C.c_void_p()
        capi.check(capi.lib().sg_spann_create(C.byref(h), device, c.shape[1], c.shape[0], num_probes, _METRIC[metric],
                                              _ptr(c, C.c_float), _ptr(pc, C.c_float), pc.shape[2], pc.shape[1],
                                              _ptr(off, C.c_uint64), _ptr(i, C.c_uint32), _ptr(cd, C.c_uint8)))
        return cls(h, c.shape[1])

    @classmethod
    def load_from_file(cls, path: str, num_probes: int = 10, verify_checksum: bool = True, device: int = 0) -> "GpuSpannIndex":
        """SpannIndex::load_from_file: the reference's "SPAN" v1 file -> a device index."""
        info = spann_file_inspect(path, verify_checksum)
        h = C.c_void_p()
        capi.check(capi.lib().sg_spann_load_file(C.byref(h), device, path.encode(), num_probes, int(verify_checksum)))
        return cls(h, info["dimension"])

    def save_to_file(self, path: str) -> None:
        """SpannIndex::save_to_file"""
        capi.check(capi.lib().sg_spann_save_file(self._h, path.encode()))

    def close(self) -> None:
        if self._h:
            capi.lib().sg_spann_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        p, n, pr = C.c_uint32(), C.c_uint64(), C.c_uint32()
        capi.check(capi.lib().sg_spann_info(self._h, C.byref(p), C.byref(n), C.byref(pr)))
        return dict(num_partitions=p.value, len=n.value, num_probes=pr.value)

    def __len__(self) -> int:
        return self.info()["len"]

    def arrays(self, pq_shape):
        """-> (centroids [P,dim], pq_centroids pq_shape, part_off [P+1], ids [n], codes [n, m])"""
        inf = self.info()
        cent = np.zeros((inf["num_partitions"], self.dim), np.float32)
        pqc = np.zeros(pq_shape, np.float32)
        off = np.zeros(inf["num_partitions"] + 1, np.uint64)
        ids = np.zeros(max(inf["len"], 1), np.uint32)
        codes = np.zeros((max(inf["len"], 1), pq_shape[0]), np.uint8)
        capi.check(capi.lib().sg_spann_get(self._h, _ptr(cent, C.c_float), _ptr(pqc, C.c_float), _ptr(off, C.c_uint64),
                                           _ptr(ids, C.c_uint32), _ptr(codes, C.c_uint8)))
        return cent, pqc, off, ids[:inf["len"]], codes[:inf["len"]]

    def set_probes(self, num_probes: int) -> None:
        capi.check(capi.lib().sg_spann_set_probes(self._h, num_probes))

    def insert(self, vector_id: int, vector) -> None:
        v = np.ascontiguousarray(vector, np.float32).ravel()
        if v.size != self.dim:
            raise capi.SgError(capi.SG_ERR_DIM, f"expected a {self.dim}-d vector, got {v.size}")
        capi.check(capi.lib().sg_spann_insert(self._h, vector_id, _ptr(v, C.c_float)))

    def search_batch(self, queries, k: int):
        q = np.ascontiguousarray(queries, np.float32)
        if q.ndim != 2 or q.shape[1] != self.dim:                      # spann.rs:587-593: an error, not a panic
            raise capi.SgError(capi.SG_ERR_DIM, f"Query dimension {q.shape[-1]} doesn't match index dimension {self.dim}")
        b = q.shape[0]
        ids = np.empty((b, k), np.uint32)
        dist = np.empty((b, k), np.float32)
        cnt = np.zeros(b, np.uint32)
        capi.check(capi.lib().sg_spann_search(self._h, _ptr(q, C.c_float), b, k, _ptr(ids, C.c_uint32), _ptr(dist, C.c_float),
                                              _ptr(cnt, C.c_uint32)))
        return ids, dist, cnt

    def search(self, query, k: int):
        ids, dist, cnt = self.search_batch(np.asarray(query, np.float32)[None, :], k)
        return [(int(i), float(d)) for i, d in zip(ids[0, :cnt[0]], dist[0, :cnt[0]])]


# ---- the "SPAN" v1 file (spann.rs:10-56): plain host code, no device needed ---------------------------------
def spann_file_inspect(path: str, verify_checksum: bool = True) -> dict:
    n = C.c_uint64()
    v = [C.c_uint32() for _ in range(6)]
    capi.check(capi.lib().sg_spann_file_inspect(path.encode(), int(verify_checksum), C.byref(n), *[C.byref(x) for x in v]))
    keys = ("num_partitions", "dimension", "num_subvectors", "num_centroids", "subvec_dim", "metric")
    return dict(num_vectors=n.value, **{k: x.value for k, x in zip(keys, v)})


def spann_file_read(path: str):
    """-> (centroids [P,dim], pq_centroids [m,kc,ds], part_off [P+1], ids [n], codes [n,m])"""
    i = spann_file_inspect(path, False)
    cent = np.zeros((i["num_partitions"], i["dimension"]), np.float32)
    pqc = np.zeros((i["num_subvectors"], i["num_centroids"], i["subvec_dim"]), np.float32)
    off = np.zeros(i["num_partitions"] + 1, np.uint64)
    ids = np.zeros(max(i["num_vectors"], 1), np.uint32)
    codes = np.zeros((max(i["num_vectors"], 1), i["num_subvectors"]), np.uint8)
    capi.check(capi.lib().sg_spann_file_read(path.encode(), _ptr(cent, C.c_float), _ptr(pqc, C.c_float), _ptr(off, C.c_uint64),
                                             _ptr(ids, C.c_uint32), _ptr(codes, C.c_uint8)))
    return cent, pqc, off, ids[:i["num_vectors"]], codes[:i["num_vectors"]]


def spann_file_write(path: str, centroids, pq_centroids, part_off, ids, codes, metric: int = 0) -> None:
    c = np.ascontiguousarray(centroids, np.float32)
    pc = np.ascontiguousarray(pq_centroids, np.float32)
    off = np.ascontiguousarray(part_off, np.uint64)
    i = np.ascontiguousarray(ids, np.uint32)
    cd = np.ascontiguousarray(codes, np.uint8)
    capi.check(capi.lib().sg_spann_file_write(path.encode(), c.shape[1], c.shape[0], metric, _ptr(c, C.c_float), pc.shape[0],
                                              pc.shape[1], pc.shape[2], _ptr(pc, C.c_float), _ptr(off, C.c_uint64),
                                              _ptr(i, C.c_uint32), _ptr(cd, C.c_uint8)))
