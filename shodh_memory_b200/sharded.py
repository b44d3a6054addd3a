"""Row-sharded store across the GPUs of one box (BASELINE.json config 4).

One process per GPU (torch.distributed).  GPU g holds the contiguous row block
[g*N/W, (g+1)*N/W) and reports global ids (id_base).  A search is: local exact sweep ->
ONE exchange of the per-shard (id, dist)[b,k] lists (b*k*8 bytes per rank, pure latency on
NVLink/NVSwitch) -> k-way merge on every rank with the same (dist, id) rule.  Product form
("p2p"): sg_store_search_sharded_device -- the kernels that produce a query's final local
result push it into every peer's symmetric buffer themselves, then one wait + merge kernel.
Baseline form ("nccl"): one all_gather_into_tensor + a strided merge straight out of the
receive buffer.

torch is plumbing here (device buffers, streams, the process group); the sweep and the
merge are this package's CUDA kernels.  The packing / layout logic is backend-agnostic so the
gloo CPU tests can exercise it (tests/test_sharded_cpu.py)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import capi
from .vector_store import GpuVectorStore, merge_topk_device


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous row block of `rank`: [rank*N/W, (rank+1)*N/W) (integer division, no gaps)."""
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def gathered_strides(b: int, k: int) -> tuple[int, int, int]:
    """Layout of the all-gather receive buffer [world][2][b][k] (int32 words): returns
    (list_stride, query_stride, dist_offset) for the strided merge."""
    return 2 * b * k, k, b * k


def exchange(local_pack: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """local_pack: int32 [2,b,k] (ids, dist bit patterns) -> gathered int32 [world,2,b,k]."""
    if world == 1:
        return local_pack.unsqueeze(0)
    out = torch.empty(world * local_pack.numel(), dtype=local_pack.dtype, device=local_pack.device)
    dist.all_gather_into_tensor(out, local_pack.reshape(-1), group=group)
    return out.view((world,) + tuple(local_pack.shape))


class PeerExchange:
    """Symmetric (peer-mapped) buffers for the fused exchange+merge kernel, one per (b, k)."""

    def __init__(self, rank: int, world: int, device: int, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem, self.rank, self.world, self.device = symm_mem, rank, world, device
        self.group = group if group is not None else dist.group.WORLD
        self._bufs = {}

    def get(self, b: int, k: int):
        key = (b, k)
        if key not in self._bufs:
            import ctypes as C
            words = int(capi.lib().sg_exchange_buffer_words(self.world, b, k))
            t = self.symm_mem.empty(words, dtype=torch.int32, device=torch.device("cuda", self.device))
            hdl = self.symm_mem.rendezvous(t, self.group)
            t.zero_()
            torch.cuda.synchronize()
            dist.barrier(group=self.group)               # every rank's flags are zero before anyone pushes
            ptrs = (C.c_uint64 * self.world)(*[int(p) for p in hdl.buffer_ptrs])
            self._bufs[key] = dict(t=t, hdl=hdl, ptrs=ptrs, epoch=0)
        return self._bufs[key]


class ShardedGpuStore:
    def __init__(self, n_total: int, dim: int = 384, dtype: str = "f16", rank: int = 0, world: int = 1,
                 device: int = 0, group=None, exchange: str = "auto"):
        """exchange: "p2p" = fused push+merge kernel over peer memory, "nccl" = all-gather + merge kernel
        (the baseline), "auto" = p2p when torch symmetric memory can be set up, else nccl."""
        self.n_total, self.dim, self.rank, self.world, self.device, self.group = n_total, dim, rank, world, device, group
        self.lo, self.hi = shard_range(n_total, rank, world)
        self.local = GpuVectorStore(dim, dtype, device)
        self.local.set_id_base(self.lo)
        self._bufs = {}
        self.peer = None
        self.last_ann_flags = None
        self.exchange_mode = "nccl"
        if world > 1 and exchange in ("auto", "p2p") and world <= 8:
            try:
                self.peer = PeerExchange(rank, world, device, group)
                self.peer.get(1, 1)                      # probe: allocation + rendezvous must work on every rank
                ok = torch.ones(1, device=torch.device("cuda", device))
            except Exception as e:                       # noqa: BLE001 - any failure means "no peer mapping here"
                if exchange == "p2p":
                    raise
                self.peer = None
                ok = torch.zeros(1, device=torch.device("cuda", device))
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # all ranks must agree on the mode
            if ok.item() > 0 and self.peer is not None:
                self.exchange_mode = "p2p"
            else:
                self.peer = None

    def fill_synthetic(self, seed: int) -> None:
        self.local.reserve(self.hi - self.lo)
        self.local.append_synthetic(seed, self.lo, self.hi - self.lo)

    def add_rows(self, rows_global: np.ndarray) -> None:
        """rows_global: the FULL [n_total, dim] matrix (every rank passes the same); keeps its block."""
        self.local.add_vectors(rows_global[self.lo:self.hi])

    def _buffers(self, b: int, k: int):
        key = (b, k)
        if key not in self._bufs:
            dev = torch.device("cuda", self.device)
            self._bufs[key] = dict(
                pack=torch.empty((2, b, k), dtype=torch.int32, device=dev),
                cnt=torch.empty(b, dtype=torch.int32, device=dev),
                ann_flags=torch.zeros(b, dtype=torch.int32, device=dev),
                out_ids=torch.empty((b, k), dtype=torch.int32, device=dev),
                out_dist=torch.empty((b, k), dtype=torch.float32, device=dev),
                out_cnt=torch.empty(b, dtype=torch.int32, device=dev))
        return self._bufs[key]

    def build_graph(self, max_degree: int = 32, search_list_size: int = 75, alpha: float = 1.2, seed: int = 0) -> None:
        """One Vamana graph per shard (VamanaIndex::build over this rank's rows); search_device(..., ann=True) then
        walks every shard's graph and merges the per-shard results exactly like the exact path does."""
        self.local.build_graph(max_degree, search_list_size, alpha, seed + self.rank)

    def search_device(self, d_queries: torch.Tensor, k: int, ann: bool = False):
        """d_queries: f32 [b,dim] on this rank's GPU (same on every rank).  Asynchronous on the
        current torch stream.  Returns (ids int32 [b,k] (bit pattern of u32), dist f32 [b,k], counts).
        ann=True: the local step is the graph search (sg_store_search_ann_device) instead of the exact sweep."""
        b = d_queries.shape[0]
        bufs = self._buffers(b, k)
        stream = torch.cuda.current_stream().cuda_stream
        pack = bufs["pack"]
        if not ann and self.world > 1 and self.peer is not None:
            # product path: search + push (inside the result kernels) + wait/merge, one C-ABI call
            x = self.peer.get(b, k)
            x["epoch"] += 1
            self.local.search_sharded_device(d_queries.data_ptr(), b, k, x["ptrs"], self.rank, self.world, x["epoch"],
                                             bufs["out_ids"].data_ptr(), bufs["out_dist"].data_ptr(),
                                             bufs["out_cnt"].data_ptr(), stream)
            return bufs["out_ids"], bufs["out_dist"], bufs["out_cnt"]
        if ann:
            # flags: bit 0 = a shard's visited table filled up (its list may be incomplete); kept per call in
            # self.last_ann_flags so the caller can re-run those queries through the host form, which retries
            self.local.search_ann_device(d_queries.data_ptr(), b, k, pack[0].data_ptr(), pack[1].data_ptr(),
                                         bufs["cnt"].data_ptr(), bufs["ann_flags"].data_ptr(), stream)
            self.last_ann_flags = bufs["ann_flags"]
        else:
            self.local.search_device(d_queries.data_ptr(), b, k, pack[0].data_ptr(), pack[1].data_ptr(),
                                     bufs["cnt"].data_ptr(), stream)
        if self.world == 1:
            return pack[0], pack[1].view(torch.float32), bufs["cnt"]
        if self.peer is not None:
            x = self.peer.get(b, k)
            x["epoch"] += 1
            capi.check(capi.lib().sg_exchange_merge_topk_device(
                self.device, x["ptrs"], self.rank, self.world, x["epoch"], pack[0].data_ptr(), pack[1].data_ptr(), b, k,
                bufs["out_ids"].data_ptr(), bufs["out_dist"].data_ptr(), bufs["out_cnt"].data_ptr(), stream))
            return bufs["out_ids"], bufs["out_dist"], bufs["out_cnt"]
        gathered = exchange(pack, self.world, self.group)
        ls, qs, doff = gathered_strides(b, k)
        base = gathered.data_ptr()
        merge_topk_device(self.device, base, base + 4 * doff, b, self.world, k, bufs["out_ids"].data_ptr(),
                          bufs["out_dist"].data_ptr(), bufs["out_cnt"].data_ptr(), stream,
                          list_stride=ls, query_stride=qs)
        return bufs["out_ids"], bufs["out_dist"], bufs["out_cnt"]

    def search(self, queries_host: torch.Tensor, k: int):
        """End-to-end form: pinned host queries in, pinned host results out (H2D + D2H inside)."""
        b = queries_host.shape[0]
        key = ("h", b, k)
        if key not in self._bufs:
            dev = torch.device("cuda", self.device)
            self._bufs[key] = dict(dq=torch.empty((b, self.dim), dtype=torch.float32, device=dev),
                                   h_ids=torch.empty((b, k), dtype=torch.int32).pin_memory(),
                                   h_dist=torch.empty((b, k), dtype=torch.float32).pin_memory(),
                                   h_cnt=torch.empty(b, dtype=torch.int32).pin_memory())
        hb = self._bufs[key]
        hb["dq"].copy_(queries_host, non_blocking=True)
        ids, dst, cnt = self.search_device(hb["dq"], k)
        hb["h_ids"].copy_(ids, non_blocking=True)
        hb["h_dist"].copy_(dst, non_blocking=True)
        hb["h_cnt"].copy_(cnt, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return hb["h_ids"], hb["h_dist"], hb["h_cnt"]

    def ann_incomplete(self) -> torch.Tensor | None:
        """After search_device(..., ann=True): bool [b], True where ANY shard flagged its result incomplete
        (visited table full) -- all ranks get the same answer (one all-reduce of b ints)."""
        if self.last_ann_flags is None:
            return None
        f = (self.last_ann_flags & 1).clone()
        if self.world > 1:
            dist.all_reduce(f, op=dist.ReduceOp.MAX, group=self.group)
        return f > 0

    def close(self) -> None:
        self.local.close()
