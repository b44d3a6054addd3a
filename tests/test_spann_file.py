"""The reference's persisted SPANN index file ("SPAN" v1, src/vector_db/spann.rs:10-56): the file functions are
plain host code and are checked here on the CPU against a test-side restatement of save_to_file (:750-888, header
byte order :221-252, FNV-1a-64 checksum :1091-1098).  Loading a file into a device index is in the GPU test below."""
import struct

import numpy as np
import pytest


F = np.float32


def align64(x):
    return (x + 63) & ~63


def write_span(path, cent, pqc, off, ids, codes, metric=0, corrupt=False, version=1, pq_enabled=1):
    P, dim = cent.shape
    m = pqc.shape[0]
    c_off = align64(128)
    cb_off = align64(c_off + P * dim * 4)
    ix_off = align64(cb_off + 12 + m * 256 * 8 * 4)
    d_off = align64(ix_off + P * 12)
    entry = 4 + m
    n = int(off[-1])
    body = bytearray(d_off + n * entry - 128)
    def put(at, b):
        body[at - 128:at - 128 + len(b)] = b
    put(c_off, np.ascontiguousarray(cent, "<f4").tobytes())
    put(cb_off, struct.pack("<III", m, 256, 8) + np.ascontiguousarray(pqc, "<f4").tobytes())
    rel = 0
    for p in range(P):
        cnt = int(off[p + 1] - off[p])
        put(ix_off + p * 12, struct.pack("<QI", rel, cnt))
        for e in range(int(off[p]), int(off[p + 1])):
            put(d_off + rel + (e - int(off[p])) * entry, struct.pack("<I", int(ids[e])) + bytes(codes[e]))
        rel += cnt * entry
    h = 0xcbf29ce484222325
    for byte in bytes(body):
        h = ((h ^ byte) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    if corrupt:
        h ^= 1
    hdr = bytearray(128)
    hdr[0:4] = b"SPAN"
    struct.pack_into("<I", hdr, 4, version)
    struct.pack_into("<Q", hdr, 8, n)
    struct.pack_into("<I", hdr, 16, P)
    struct.pack_into("<I", hdr, 20, dim)
    hdr[24] = pq_enabled
    struct.pack_into("<I", hdr, 25, m)
    hdr[29] = metric
    struct.pack_into("<Q", hdr, 30, h)
    struct.pack_into("<QQQQ", hdr, 38, c_off, cb_off, ix_off, d_off)
    with open(path, "wb") as f:
        f.write(bytes(hdr) + bytes(body))


def small_index(seed=0, P=7, dim=32, n=300):
    rng = np.random.default_rng(seed)
    cent = rng.standard_normal((P, dim)).astype(F)
    pqc = rng.standard_normal((dim // 8, 256, 8)).astype(F)
    sizes = rng.multinomial(n, np.ones(P) / P)
    sizes[2] = 0                                                    # an empty partition
    off = np.zeros(P + 1, np.uint64)
    off[1:] = np.cumsum(sizes)
    n = int(off[-1])
    ids = rng.permutation(n).astype(np.uint32)
    codes = rng.integers(0, 256, (n, dim // 8), dtype=np.uint8)
    return cent, pqc, off, ids, codes


def test_reader_on_a_reference_layout_file(sg, tmp_path):
    cent, pqc, off, ids, codes = small_index()
    p = str(tmp_path / "index.spann")
    write_span(p, cent, pqc, off, ids, codes, metric=1)
    info = sg.spann_file_inspect(p)
    assert info == dict(num_vectors=len(ids), num_partitions=7, dimension=32, num_subvectors=4, num_centroids=256, subvec_dim=8,
                        metric=1)
    c2, q2, o2, i2, d2 = sg.spann_file_read(p)
    assert np.array_equal(c2.view(np.uint32), cent.view(np.uint32)) and np.array_equal(q2.view(np.uint32), pqc.view(np.uint32))
    assert np.array_equal(o2, off) and np.array_equal(i2, ids) and np.array_equal(d2, codes)


def test_writer_is_byte_identical_to_the_restated_save_to_file(sg, tmp_path):
    cent, pqc, off, ids, codes = small_index(seed=3, P=11, dim=64, n=500)
    a, b = str(tmp_path / "a.spann"), str(tmp_path / "b.spann")
    write_span(a, cent, pqc, off, ids, codes, metric=0)
    sg.spann_file_write(b, cent, pqc, off, ids, codes, metric=0)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert sg.spann_file_inspect(b)["num_vectors"] == len(ids)       # its own checksum verifies (verify_index_file)


def test_corruption_and_unsupported_files_are_errors(sg, tmp_path):
    cent, pqc, off, ids, codes = small_index(seed=5)
    p = str(tmp_path / "bad.spann")
    write_span(p, cent, pqc, off, ids, codes, corrupt=True)
    with pytest.raises(sg.SgError, match="Checksum mismatch"):       # spann.rs:904-912
        sg.spann_file_inspect(p)
    assert sg.spann_file_inspect(p, verify_checksum=False)["num_partitions"] == 7
    write_span(p, cent, pqc, off, ids, codes, version=2)
    with pytest.raises(sg.SgError):
        sg.spann_file_inspect(p)
    write_span(p, cent, pqc, off, ids, codes, pq_enabled=0)
    with pytest.raises(sg.SgError, match="PQ"):
        sg.spann_file_inspect(p)
    open(p, "wb").write(b"SPAN" + b"\0" * 20)
    with pytest.raises(sg.SgError):
        sg.spann_file_inspect(p)
    with pytest.raises(sg.SgError):
        sg.spann_file_inspect(str(tmp_path / "missing.spann"))
    write_span(p, cent, pqc, off, ids, codes)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-100])                                    # truncated posting data
    with pytest.raises(sg.SgError):
        sg.spann_file_read(p)


@pytest.mark.gpu
def test_load_search_save_round_trip(sg, oracle, tmp_path):
    """spann.rs:1153-1183 test_spann_save_and_load: save, verify, load, search the loaded index"""
    v = np.random.default_rng(0).random((500, 384), dtype=F)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    ref = oracle.SpannOracle(384, kmeans_iterations=4, pq_iterations=2).build(v.astype(F), seed=1)
    off, ids, codes = ref.csr()
    p = str(tmp_path / "test.spann")
    write_span(p, ref.centroids, ref.pq.centroids, off, ids, codes)    # as the reference would have saved it
    idx = sg.GpuSpannIndex.load_from_file(p)
    assert len(idx) == 500 and idx.info()["num_partitions"] == ref.centroids.shape[0]
    for q in v[:6]:
        ri, rd = ref.search(q, 10)
        assert idx.search(q, 10) == [(int(a), float(b)) for a, b in zip(ri, rd)]
    idx.insert(9000, v[3])
    p2 = str(tmp_path / "resaved.spann")
    idx.save_to_file(p2)
    assert sg.spann_file_inspect(p2)["num_vectors"] == 501
    ref.insert(9000, v[3])
    roff, rids, rcodes = ref.csr()
    _, _, o2, i2, d2 = sg.spann_file_read(p2)
    assert np.array_equal(o2, roff) and np.array_equal(i2, rids) and np.array_equal(d2, rcodes)
    idx.close()
