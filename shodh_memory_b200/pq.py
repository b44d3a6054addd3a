"""Host-side mirror of the reference's Product Quantization surface (src/vector_db/pq.rs) over the C ABI:
`GpuProductQuantizer` = ProductQuantizer + CompressedVectorStore.  All compute happens in libshodh_gpu.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def _ptr(a: np.ndarray, ty):
    return a.ctypes.data_as(C.POINTER(ty))


class GpuProductQuantizer:
    def __init__(self, handle, dim: int, device: int):
        self._h = handle
        self.device = device
        self.dim = dim
        d, m, k, n = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        capi.check(capi.lib().sg_pq_info(self._h, C.byref(d), C.byref(m), C.byref(k), C.byref(n)))
        self.num_subvectors, self.num_centroids = m.value, k.value
        self.subvec_dim = dim // m.value

    # -- ProductQuantizer ---------------------------------------------------------------------
    @classmethod
    def from_centroids(cls, centroids, device: int = 0) -> "GpuProductQuantizer":
        """centroids [num_subvectors, num_centroids, subvec_dim] f32 (ProductQuantizer.centroids)"""
        c = np.ascontiguousarray(centroids, np.float32)
        m, k, ds = c.shape
        h = C.c_void_p()
        capi.check(capi.lib().sg_pq_create(C.byref(h), device, m * ds, ds, k, _ptr(c, C.c_float)))
        return cls(h, m * ds, device)

    @classmethod
    def train(cls, vectors, orders, subvec_dim: int = 8, num_centroids: int = 256, kmeans_iterations: int = 20,
              device: int = 0) -> "GpuProductQuantizer":
        """ProductQuantizer::train; orders [num_subvectors, n] = the shuffled index order of each subspace's k-means."""
        v = np.ascontiguousarray(vectors, np.float32)
        o = np.ascontiguousarray(orders, np.uint32)
        n, dim = v.shape
        if o.shape != (dim // subvec_dim, n):
            raise capi.SgError(capi.SG_ERR_DIM, f"expected orders [{dim // subvec_dim},{n}], got {o.shape}")
        h = C.c_void_p()
        capi.check(capi.lib().sg_pq_train(C.byref(h), device, dim, subvec_dim, num_centroids, kmeans_iterations,
                                          _ptr(v, C.c_float), n, _ptr(o, C.c_uint32)))
        return cls(h, dim, device)

    def close(self) -> None:
        if self._h:
            capi.lib().sg_pq_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        n = C.c_uint64()
        capi.check(capi.lib().sg_pq_info(self._h, None, None, None, C.byref(n)))
        return n.value

    @property
    def centroids(self) -> np.ndarray:
        out = np.zeros((self.num_subvectors, self.num_centroids, self.subvec_dim), np.float32)
        capi.check(capi.lib().sg_pq_get_centroids(self._h, _ptr(out, C.c_float)))
        return out

    def encode(self, vectors) -> np.ndarray:
        """ProductQuantizer::encode_batch -> codes [n, num_subvectors] u8"""
        v = np.ascontiguousarray(vectors, np.float32).reshape(-1, self.dim)
        out = np.zeros((v.shape[0], self.num_subvectors), np.uint8)
        capi.check(capi.lib().sg_pq_encode(self._h, _ptr(v, C.c_float), v.shape[0], _ptr(out, C.c_uint8)))
        return out

    def decode(self, codes) -> np.ndarray:
        c = np.ascontiguousarray(codes, np.uint8).reshape(-1, self.num_subvectors)
        out = np.zeros((c.shape[0], self.dim), np.float32)
        capi.check(capi.lib().sg_pq_decode(self._h, _ptr(c, C.c_uint8), c.shape[0], _ptr(out, C.c_float)))
        return out

    def build_distance_table(self, query) -> np.ndarray:
        q = np.ascontiguousarray(query, np.float32).ravel()
        if q.size != self.dim:
            raise capi.SgError(capi.SG_ERR_DIM, f"expected a {self.dim}-d query, got {q.size}")
        out = np.zeros((self.num_subvectors, self.num_centroids), np.float32)
        capi.check(capi.lib().sg_pq_table(self._h, _ptr(q, C.c_float), _ptr(out, C.c_float)))
        return out

    # -- CompressedVectorStore ------------------------------------------------------------------
    def add(self, ids, vectors) -> None:
        i = np.ascontiguousarray(ids, np.uint32).ravel()
        v = np.ascontiguousarray(vectors, np.float32).reshape(-1, self.dim)
        if i.size != v.shape[0]:
            raise capi.SgError(capi.SG_ERR_DIM, "one id per vector")
        capi.check(capi.lib().sg_pq_add(self._h, _ptr(i, C.c_uint32), _ptr(v, C.c_float), i.size))

    def search_batch(self, queries, k: int):
        """CompressedVectorStore::search for b queries -> (ids [b,k], dist [b,k], counts [b])"""
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, self.dim)
        b = q.shape[0]
        ids = np.empty((b, k), np.uint32)
        dist = np.empty((b, k), np.float32)
        cnt = np.zeros(b, np.uint32)
        capi.check(capi.lib().sg_pq_search(self._h, _ptr(q, C.c_float), b, k, _ptr(ids, C.c_uint32), _ptr(dist, C.c_float),
                                           _ptr(cnt, C.c_uint32)))
        return ids, dist, cnt

    def search(self, query, k: int):
        ids, dist, cnt = self.search_batch(np.asarray(query, np.float32)[None, :], k)
        return [(int(i), float(d)) for i, d in zip(ids[0, :cnt[0]], dist[0, :cnt[0]])]
