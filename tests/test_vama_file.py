"""The reference's persisted index file ("VAMA" v1, src/vector_db/vamana_persist.rs:6-34) loads into the GPU
store: parser checks on CPU (sg_vama_inspect), full load + search parity on the GPU.  The writer below is a
test-side restatement of save_to_file (:175-280): header byte offsets :96-111, FNV-1a-64 checksum :155-163."""
import struct

import numpy as np
import pytest

SEED = 23821


def write_vama(path, vectors, deleted=(), neighbors=None, metric=0, corrupt=False, version=1, medoid=0, max_degree=32):
    n, dim = vectors.shape
    rng = np.random.default_rng(0)
    body = bytearray()
    for d in deleted:
        body += struct.pack("<I", d)
    for i in range(n):                                   # graph: u16 count + u32 neighbours (content irrelevant here)
        nb = neighbors[i] if neighbors is not None else rng.integers(0, n, int(rng.integers(0, 5))).tolist()
        body += struct.pack("<H", len(nb)) + struct.pack(f"<{len(nb)}I", *nb)
    off = 64 + len(body)
    pad = (-off) % 64
    body += b"\0" * pad
    body += np.ascontiguousarray(vectors, "<f4").tobytes()
    h = 0xcbf29ce484222325
    for byte in bytes(body):
        h = ((h ^ byte) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    if corrupt:
        h ^= 1
    hdr = bytearray(64)
    hdr[0:4] = b"VAMA"
    struct.pack_into("<I", hdr, 4, version)
    struct.pack_into("<Q", hdr, 8, n)
    struct.pack_into("<I", hdr, 16, dim)
    struct.pack_into("<I", hdr, 20, max_degree)
    struct.pack_into("<I", hdr, 24, medoid)
    hdr[28] = metric
    struct.pack_into("<I", hdr, 29, len(deleted))
    struct.pack_into("<Q", hdr, 33, 0)
    struct.pack_into("<Q", hdr, 41, h)
    with open(path, "wb") as f:
        f.write(bytes(hdr) + bytes(body))
    return 64 + len(body) - vectors.size * 4


def test_vama_parser_on_cpu(sg, oracle, tmp_path):
    vecs = oracle.synth_corpus(SEED, 0, 500, 128)
    p = str(tmp_path / "vamana.idx")
    off = write_vama(p, vecs, deleted=[3, 77])
    info = sg.vama_inspect(p)
    assert info == dict(num_vectors=500, dimension=128, deleted_count=2, metric=0, vectors_offset=off)
    assert off % 64 == 0
    write_vama(p, vecs, corrupt=True)
    with pytest.raises(sg.SgError):                       # vamana_persist.rs:303-312 (checksum corruption test :500-533)
        sg.vama_inspect(p)
    assert sg.vama_inspect(p, verify_checksum=False)["num_vectors"] == 500
    write_vama(p, vecs, version=2)
    with pytest.raises(sg.SgError):
        sg.vama_inspect(p)
    open(p, "wb").write(b"nope")
    with pytest.raises(sg.SgError):
        sg.vama_inspect(p)


@pytest.mark.gpu
def test_vama_load_and_search(sg, oracle, tmp_path):
    n = 20_000
    vecs = oracle.synth_corpus(SEED, 0, n, 384)
    p = str(tmp_path / "vamana.idx")
    deleted = [0, 5, 19_999]
    write_vama(p, vecs, deleted=deleted)
    for dtype in ("f32", "f16"):
        st = sg.GpuVectorStore(384, dtype)
        st.load_from_file(p)
        assert len(st) == n and st.deleted_count() == 3 and st.is_deleted(5)
        seen = vecs if dtype == "f32" else oracle.round_f16(vecs)
        assert np.array_equal(st.get_vectors(100, 50).view(np.uint32), seen[100:150].view(np.uint32))
        q, _ = oracle.synth_queries(SEED, n, 3, 384)
        ids, dist, cnt = st.search_batch(q, 10)
        for i in range(3):
            rid, rdist = oracle.brute_force_search(seen, q[i], 10, deleted=deleted)
            assert np.array_equal(ids[i], rid) and np.array_equal(dist[i].view(np.uint32), rdist.view(np.uint32))
        with pytest.raises(sg.SgError):                   # needs an empty store
            st.load_from_file(p)
        st.close()
    wrong = sg.GpuVectorStore(128, "f16")
    with pytest.raises(sg.SgError):                       # dimension mismatch is an error
        wrong.load_from_file(p)
    wrong.close()
