// persist.cu -- loading the reference's persisted index file ("VAMA" v1) straight into the GPU store.
//
// SURVEY.md 8f item 2.  File format (reference src/vector_db/vamana_persist.rs:6-34, header byte
// offsets :96-111): 64-byte little-endian header { "VAMA", version u32 = 1, num_vectors u64 @8,
// dimension u32 @16, max_degree u32 @20, medoid u32 @24, distance_metric u8 @28, deleted_count u32
// @29, incremental_inserts u64 @33, checksum u64 @41 }, then deleted ids [u32; deleted_count], then
// the graph (per node: neighbour_count u16 + [u32; count]), then -- aligned to 64 bytes -- the
// vectors, row-major [[f32; dimension]; num_vectors].  checksum = FNV-1a-64 over every byte after
// the header (:155-163, :269, :303-312).  The vectors section is exactly the dense f32 matrix
// sg_store_append takes; the graph section (with the header's medoid and max_degree) is uploaded with
// sg_store_set_graph so the index is searchable by sg_store_search_ann exactly as the reference
// would search it (the exact sweep needs no graph).  This is host-side file plumbing; no arithmetic
// of the hot path happens here.
#include <cstdio>
#include <cstring>
#include <vector>

#include "sg_store_internal.h"

namespace {

struct VamaInfo {
    uint64_t num_vectors = 0;
    uint32_t dimension = 0, max_degree = 0, medoid = 0, deleted_count = 0;
    uint8_t metric = 0;
    uint64_t incremental_inserts = 0, checksum = 0, vectors_offset = 0, file_size = 0;
};

uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

// parses header + walks the graph section to find the vectors; optionally verifies the checksum
int vama_parse(FILE *f, VamaInfo &vi, std::vector<uint32_t> *deleted, int verify, int *checksum_ok,
               std::vector<uint32_t> *deg_out = nullptr, std::vector<uint32_t> *nbr_out = nullptr)
{
    uint8_t h[64];
    if (fseek(f, 0, SEEK_END) != 0) return SG_ERR_INVALID;
    vi.file_size = (uint64_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    if (vi.file_size < 64 || fread(h, 1, 64, f) != 64 || memcmp(h, "VAMA", 4) != 0) {
        sg::set_error("not a VAMA file (bad magic or shorter than the 64-byte header)");
        return SG_ERR_INVALID;
    }
    if (rd32(h + 4) != 1) {
        sg::set_error("unsupported VAMA version %u", rd32(h + 4));
        return SG_ERR_UNSUPPORTED;
    }
    vi.num_vectors = rd64(h + 8);
    vi.dimension = rd32(h + 16);
    vi.max_degree = rd32(h + 20);
    vi.medoid = rd32(h + 24);
    vi.metric = h[28];
    vi.deleted_count = rd32(h + 29);
    vi.incremental_inserts = rd64(h + 33);
    vi.checksum = rd64(h + 41);
    uint64_t off = 64 + (uint64_t)vi.deleted_count * 4;
    if (off > vi.file_size) {
        sg::set_error("VAMA file truncated in the deleted-ids section");
        return SG_ERR_INVALID;
    }
    // plausibility BEFORE any allocation sized by header fields (a corrupt header must give SG_ERR_INVALID, not a
    // std::bad_alloc thrown through the C ABI): every node costs >= 2 bytes of graph and dimension * 4 bytes of vector
    {
        const uint64_t rest = vi.file_size - off;
        const unsigned __int128 need_bytes = (unsigned __int128)vi.num_vectors * (2ull + (uint64_t)vi.dimension * 4ull);
        if (vi.dimension == 0 || vi.num_vectors > 0xFFFFFFFEull || need_bytes > (unsigned __int128)rest) {
            sg::set_error("VAMA header implausible for a %llu-byte file: %llu vectors x dim %u", (unsigned long long)vi.file_size,
                          (unsigned long long)vi.num_vectors, vi.dimension);
            return SG_ERR_INVALID;
        }
    }
    if (deleted) {
        deleted->resize(vi.deleted_count);
        if (vi.deleted_count && fread(deleted->data(), 4, vi.deleted_count, f) != vi.deleted_count) return SG_ERR_INVALID;
    }
    if (deg_out) deg_out->assign(vi.num_vectors, 0);
    if (nbr_out) {
        if (vi.max_degree == 0 || vi.max_degree > 64) {
            sg::set_error("VAMA graph: max_degree %u outside 1..64", vi.max_degree);
            return SG_ERR_UNSUPPORTED;
        }
        nbr_out->assign(vi.num_vectors * (size_t)vi.max_degree, 0xFFFFFFFFu);
    }
    // graph: variable-length records; buffered walk
    fseek(f, (long)off, SEEK_SET);
    std::vector<uint8_t> buf(1 << 20);
    size_t have = 0, pos = 0;
    uint64_t file_pos = off;
    auto need = [&](size_t n) -> bool {
        if (have - pos >= n) return true;
        memmove(buf.data(), buf.data() + pos, have - pos);
        have -= pos;
        pos = 0;
        size_t got = fread(buf.data() + have, 1, buf.size() - have, f);
        have += got;
        return have - pos >= n;
    };
    for (uint64_t node = 0; node < vi.num_vectors; ++node) {
        if (!need(2)) {
            sg::set_error("VAMA file truncated in the graph section (node %llu)", (unsigned long long)node);
            return SG_ERR_INVALID;
        }
        uint16_t cnt;
        memcpy(&cnt, buf.data() + pos, 2);
        pos += 2;
        file_pos += 2;
        if (nbr_out && cnt > vi.max_degree) {
            sg::set_error("VAMA graph: node %llu lists %u neighbours, max_degree is %u", (unsigned long long)node, cnt,
                          vi.max_degree);
            return SG_ERR_INVALID;
        }
        if (deg_out) (*deg_out)[node] = cnt;
        size_t skip = (size_t)cnt * 4, copied = 0;
        while (skip) {
            if (!need(1)) {
                sg::set_error("VAMA file truncated in the graph section (node %llu)", (unsigned long long)node);
                return SG_ERR_INVALID;
            }
            size_t take = have - pos < skip ? have - pos : skip;
            if (nbr_out) memcpy(reinterpret_cast<uint8_t *>(nbr_out->data() + node * vi.max_degree) + copied, buf.data() + pos, take);
            copied += take;
            pos += take;
            file_pos += take;
            skip -= take;
        }
    }
    vi.vectors_offset = (file_pos + 63) & ~63ull;
    const uint64_t want = vi.vectors_offset + vi.num_vectors * (uint64_t)vi.dimension * 4;
    if (want > vi.file_size) {
        sg::set_error("VAMA file truncated in the vectors section (%llu < %llu bytes)", (unsigned long long)vi.file_size,
                      (unsigned long long)want);
        return SG_ERR_INVALID;
    }
    if (verify) {
        fseek(f, 64, SEEK_SET);
        uint64_t hash = 0xcbf29ce484222325ull;
        size_t got;
        while ((got = fread(buf.data(), 1, buf.size(), f)) > 0)
            for (size_t i = 0; i < got; ++i) {
                hash ^= buf[i];
                hash *= 0x100000001b3ull;
            }
        if (checksum_ok) *checksum_ok = hash == vi.checksum;
        if (hash != vi.checksum) {
            sg::set_error("VAMA checksum mismatch: stored %llx, computed %llx", (unsigned long long)vi.checksum,
                          (unsigned long long)hash);
            return SG_ERR_INVALID;
        }
    } else if (checksum_ok) {
        *checksum_ok = -1;
    }
    return SG_OK;
}

}  // namespace

extern "C" {

static int vama_inspect_impl(const char *path, int verify_checksum, uint64_t *num_vectors_out, uint32_t *dimension_out,
                             uint32_t *deleted_count_out, uint32_t *metric_out, uint64_t *vectors_offset_out);

int sg_vama_inspect(const char *path, int verify_checksum, uint64_t *num_vectors_out, uint32_t *dimension_out,
                    uint32_t *deleted_count_out, uint32_t *metric_out, uint64_t *vectors_offset_out)
{
    return sg::abi_guard("sg_vama_inspect", [&] {
        return vama_inspect_impl(path, verify_checksum, num_vectors_out, dimension_out, deleted_count_out, metric_out,
                                 vectors_offset_out);
    });
}

static int vama_inspect_impl(const char *path, int verify_checksum, uint64_t *num_vectors_out, uint32_t *dimension_out,
                    uint32_t *deleted_count_out, uint32_t *metric_out, uint64_t *vectors_offset_out)
{
    if (!path) return SG_ERR_INVALID;
    FILE *f = fopen(path, "rb");
    if (!f) {
        sg::set_error("cannot open %s", path);
        return SG_ERR_INVALID;
    }
    VamaInfo vi;
    int rc = vama_parse(f, vi, nullptr, verify_checksum, nullptr);
    fclose(f);
    if (rc != SG_OK) return rc;
    if (num_vectors_out) *num_vectors_out = vi.num_vectors;
    if (dimension_out) *dimension_out = vi.dimension;
    if (deleted_count_out) *deleted_count_out = vi.deleted_count;
    if (metric_out) *metric_out = vi.metric;
    if (vectors_offset_out) *vectors_offset_out = vi.vectors_offset;
    return SG_OK;
}

static int load_vama_impl(sg_store *s, const char *path, int verify_checksum);

int sg_store_load_vama(sg_store *s, const char *path, int verify_checksum)
{
    if (!s || !path) return SG_ERR_INVALID;
    uint64_t have = 0;
    int rc = sg_store_len(s, &have);
    if (rc != SG_OK) return rc;
    if (have != 0) {
        sg::set_error("sg_store_load_vama needs an empty store (ids must equal the file's row numbers)");
        return SG_ERR_INVALID;
    }
    rc = sg::abi_guard("sg_store_load_vama", [&] { return load_vama_impl(s, path, verify_checksum); });
    if (rc != SG_OK && sg_store_len(s, &have) == SG_OK && have != 0) {
        // a failure part-way must not leave the (formerly empty) store half populated; the error message stays
        std::lock_guard<std::mutex> g(s->mu);
        sg::store_truncate_locked(s);
    }
    return rc;
}

static int load_vama_impl(sg_store *s, const char *path, int verify_checksum)
{
    int rc = SG_OK;
    FILE *f = fopen(path, "rb");
    if (!f) {
        sg::set_error("cannot open %s", path);
        return SG_ERR_INVALID;
    }
    VamaInfo vi;
    std::vector<uint32_t> deleted, deg, nbr;
    rc = vama_parse(f, vi, &deleted, verify_checksum, nullptr, &deg, &nbr);
    if (rc != SG_OK) {
        fclose(f);
        return rc;
    }
    uint32_t store_dim = 0;
    sg_store_dim(s, &store_dim);
    if (store_dim != vi.dimension) {
        fclose(f);
        sg::set_error("dimension mismatch: the file holds %u-d vectors, the store is %u-d", vi.dimension, store_dim);
        return SG_ERR_DIM;
    }
    int store_metric = 0;
    sg_store_metric(s, &store_metric);
    if ((int)vi.metric != store_metric) {      // DistanceMetric as u8: 0 NormalizedDotProduct, 1 Euclidean, 2 Cosine
        fclose(f);
        sg::set_error("VAMA file uses distance metric %u, the store was created with metric %d", vi.metric, store_metric);
        return SG_ERR_INVALID;
    }
    // upload the vectors section in 32 MB pieces through the normal append path (rounds to the store dtype)
    rc = sg_store_reserve(s, vi.num_vectors);
    const size_t dim = vi.dimension;
    const size_t rows_per = dim ? ((size_t)32 << 20) / (dim * 4) : 0;
    std::vector<float> chunk;
    if (rc == SG_OK && rows_per == 0) rc = SG_ERR_DIM;
    if (rc == SG_OK) {
        chunk.resize(rows_per * dim);
        fseek(f, (long)vi.vectors_offset, SEEK_SET);
        uint64_t done = 0;
        while (rc == SG_OK && done < vi.num_vectors) {
            const size_t n = vi.num_vectors - done < rows_per ? (size_t)(vi.num_vectors - done) : rows_per;
            if (fread(chunk.data(), 4, n * dim, f) != n * dim) {
                sg::set_error("short read in the vectors section");
                rc = SG_ERR_INVALID;
                break;
            }
            rc = sg_store_append(s, chunk.data(), n, nullptr);   // returns SG_ERR_DIM-like errors from the store
            done += n;
        }
    }
    fclose(f);
    if (rc != SG_OK) return rc;
    for (uint32_t id : deleted) {
        int ok = 0;
        rc = sg_store_mark_deleted(s, id, &ok);
        if (rc != SG_OK) return rc;
    }
    // the graph as the reference built it: searched as-is by sg_store_search_ann
    bool any_edge = false;
    for (uint32_t d : deg) any_edge |= d != 0;
    if (vi.num_vectors && (any_edge || vi.num_vectors == 1))
        rc = sg_store_set_graph(s, vi.max_degree, vi.medoid, deg.data(), nbr.data(), vi.max_degree, vi.num_vectors);
    return rc;
}


int sg_store_save_vama(sg_store *s, const char *path)
{
    if (!s || !path) return SG_ERR_INVALID;
    uint64_t n = 0, graph_len = 0, n_deleted = 0;
    uint32_t dim = 0, max_degree = 0, medoid = 0;
    int metric = 0;
    int rc = sg_store_len(s, &n);
    if (rc == SG_OK) rc = sg_store_dim(s, &dim);
    if (rc == SG_OK) rc = sg_store_metric(s, &metric);
    if (rc == SG_OK) rc = sg_store_graph_info(s, &max_degree, &medoid, &graph_len);
    if (rc == SG_OK) rc = sg_store_deleted_count(s, &n_deleted);
    if (rc != SG_OK) return rc;
    if (metric > 2) {
        sg::set_error("metric %d has no DistanceMetric value in the VAMA header", metric);
        return SG_ERR_UNSUPPORTED;
    }
    std::vector<uint32_t> deg(n, 0), nbr;
    if (graph_len) {
        if (graph_len != n) {
            sg::set_error("the graph covers %llu of %llu rows", (unsigned long long)graph_len, (unsigned long long)n);
            return SG_ERR_INVALID;
        }
        nbr.resize((size_t)n * max_degree);
        if ((rc = sg_store_get_graph(s, deg.data(), nbr.data(), n)) != SG_OK) return rc;
    } else {
        max_degree = 32;   // VamanaConfig::default; every list is empty
    }
    FILE *f = fopen(path, "wb+");
    if (!f) {
        sg::set_error("cannot create %s", path);
        return SG_ERR_INVALID;
    }
    uint64_t hash = 0xcbf29ce484222325ull;
    uint64_t off = 64;
    bool io_ok = true;
    auto put = [&](const void *p, size_t bytes) {
        const uint8_t *b = static_cast<const uint8_t *>(p);
        for (size_t i = 0; i < bytes; ++i) {
            hash ^= b[i];
            hash *= 0x100000001b3ull;
        }
        io_ok = io_ok && fwrite(p, 1, bytes, f) == bytes;
        off += bytes;
    };
    uint8_t h[64];
    memset(h, 0, sizeof(h));
    io_ok = fwrite(h, 1, 64, f) == 64;                      // header placeholder, rewritten with the checksum
    uint32_t n_del_written = 0;
    for (uint64_t id = 0; id < n && n_deleted; ++id) {      // deleted ids (ascending; the reference's set order is arbitrary)
        int d = 0;
        if ((rc = sg_store_is_deleted(s, (uint32_t)id, &d)) != SG_OK) break;
        if (d) {
            const uint32_t v = (uint32_t)id;
            put(&v, 4);
            ++n_del_written;
        }
    }
    for (uint64_t i = 0; i < n && rc == SG_OK; ++i) {       // graph: u16 count + u32 ids
        const uint16_t c = (uint16_t)deg[i];
        put(&c, 2);
        if (c) put(nbr.data() + i * max_degree, (size_t)c * 4);
    }
    const uint8_t zeros[64] = {0};
    if (off % 64) put(zeros, 64 - off % 64);                // vectors start 64-byte aligned (:212-215)
    const size_t rows_per = dim ? ((size_t)32 << 20) / ((size_t)dim * 4) : 0;
    std::vector<float> chunk(rows_per * dim);
    for (uint64_t done = 0; done < n && rc == SG_OK && rows_per;) {
        const size_t m = n - done < rows_per ? (size_t)(n - done) : rows_per;
        rc = sg_store_get_vectors(s, done, m, chunk.data());
        if (rc == SG_OK) put(chunk.data(), m * dim * 4);
        done += m;
    }
    if (rc == SG_OK) {
        memcpy(h, "VAMA", 4);
        const uint32_t version = 1;
        const uint64_t inc = 0;
        memcpy(h + 4, &version, 4);
        memcpy(h + 8, &n, 8);
        memcpy(h + 16, &dim, 4);
        memcpy(h + 20, &max_degree, 4);
        memcpy(h + 24, &medoid, 4);
        h[28] = (uint8_t)metric;
        memcpy(h + 29, &n_del_written, 4);
        memcpy(h + 33, &inc, 8);
        memcpy(h + 41, &hash, 8);
        io_ok = io_ok && fseek(f, 0, SEEK_SET) == 0 && fwrite(h, 1, 64, f) == 64;
    }
    io_ok = (fclose(f) == 0) && io_ok;
    if (rc == SG_OK && !io_ok) {
        sg::set_error("writing %s failed", path);
        rc = SG_ERR_INVALID;
    }
    return rc;
}

}  // extern "C"
