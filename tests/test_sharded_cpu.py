"""world_size-2 gloo test (CPU) of the N>1 host logic: shard ranges, global ids, the packed
all-gather exchange and the strided layout the merge kernel consumes.  The per-shard search and
the merge are stood in for by the CPU oracle / a numpy restatement of the key order -- only the
HOST-SIDE plumbing of shodh_memory_b200.sharded is under test here (the kernels are -m gpu)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N, DIM, B, K, SEED = 3001, 64, 5, 7, 23821


def _merge_strided_numpy(words, world, b, k, list_stride, query_stride, dist_off, oracle):
    ids_out = np.zeros((b, k), np.uint32)
    dist_out = np.zeros((b, k), np.float32)
    for q in range(b):
        cand = []
        for l in range(world):
            for j in range(k):
                at = l * list_stride + q * query_stride + j
                i = np.uint32(words[at])
                if i == 0xFFFFFFFF:
                    continue
                d = words[at + dist_off:at + dist_off + 1].view(np.float32)[0]
                cand.append((oracle.total_order_key(float(d)), int(i), d))
        cand.sort(key=lambda t: (t[0], t[1]))
        for j, c in enumerate(cand[:k]):
            ids_out[q, j], dist_out[q, j] = c[1], c[2]
    return ids_out, dist_out


def _worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle
    from shodh_memory_b200 import sharded

    corpus = oracle.synth_corpus(SEED, 0, N, DIM)
    queries, _ = oracle.synth_queries(SEED, N, B, DIM)
    lo, hi = sharded.shard_range(N, rank, world)
    ids = np.full((B, K), 0xFFFFFFFF, np.uint32)
    dst = np.full((B, K), np.inf, np.float32)
    for q in range(B):                                       # stand-in for the local GPU sweep
        rid, rdist = oracle.brute_force_search(corpus[lo:hi], queries[q], K)
        ids[q, :len(rid)] = rid + lo                         # id_base
        dst[q, :len(rid)] = rdist
    pack = torch.from_numpy(np.stack([ids.view(np.int32), dst.view(np.int32)]))
    gathered = sharded.exchange(pack, world)                 # the real all-gather, on gloo
    ls, qs, doff = sharded.gathered_strides(B, K)
    words = gathered.numpy().reshape(-1).view(np.uint32)
    got_ids, got_dist = _merge_strided_numpy(words, world, B, K, ls, qs, doff, oracle)
    for q in range(B):
        rid, rdist = oracle.brute_force_search(corpus, queries[q], K)
        assert np.array_equal(got_ids[q], rid), (rank, q)
        assert np.array_equal(got_dist[q].view(np.uint32), rdist.view(np.uint32))
    open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from shodh_memory_b200 import sharded
    for n in (0, 1, 7, 10_000_000, 12345):
        for w in (1, 2, 3, 4, 8):
            edges = [sharded.shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
    assert sharded.shard_range(10_000_000, 3, 8) == (3_750_000, 5_000_000)


def test_two_rank_exchange_and_merge_layout(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
