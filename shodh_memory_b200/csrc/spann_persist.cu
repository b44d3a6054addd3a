// spann_persist.cu -- the reference's persisted SPANN index file ("SPAN" v1) <-> the GPU index.
//
// File format (reference src/vector_db/spann.rs:10-56, header fields in to_bytes order :221-252): 128-byte
// little-endian header { "SPAN", version u32 = 1, num_vectors u64 @8, num_partitions u32 @16, dimension u32 @20,
// pq_enabled u8 @24, pq_subvectors u32 @25, distance_metric u8 @29, checksum u64 @30, centroids_offset u64 @38,
// codebook_offset u64 @46, posting_index_offset u64 @54, posting_data_offset u64 @62 }, then -- each section
// aligned to 64 bytes -- centroids [[f32; dim]; P], the PQ codebook { num_subvectors u32, num_centroids u32,
// subvec_dim u32, [[f32; subvec_dim]; num_centroids] x num_subvectors }, the posting index [(offset u64 relative to
// posting_data_offset, count u32); P] and the posting data (per entry: vector_id u32 + pq_codes [u8; num_subvectors]).
// checksum = FNV-1a-64 over every byte after the header (:1091-1098).  save_to_file :750-888, load_from_file :891-1003.
// The file functions below are plain host code (no device needed: tests/test_spann_file.py runs them on the CPU);
// sg_spann_load_file / sg_spann_save_file compose them with sg_spann_create / sg_spann_get.
#include <cstdio>
#include <cstring>
#include <vector>

#include "sg_store_internal.h"

namespace {

constexpr uint64_t SPAN_HEADER = 128;

struct SpanInfo {
    uint64_t num_vectors = 0, checksum = 0, centroids_off = 0, codebook_off = 0, index_off = 0, data_off = 0, file_size = 0;
    uint32_t n_part = 0, dim = 0, pq_subvectors = 0, m = 0, kc = 0, ds = 0;
    uint8_t pq_enabled = 0, metric = 0;
};

uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
uint64_t align64(uint64_t x) { return (x + 63) & ~63ull; }

uint64_t fnv1a(uint64_t h, const void *p, size_t n)
{
    const uint8_t *b = static_cast<const uint8_t *>(p);
    for (size_t i = 0; i < n; ++i) {
        h ^= b[i];
        h *= 0x100000001b3ull;
    }
    return h;
}

int span_parse_header(FILE *f, SpanInfo &si, int verify)
{
    uint8_t h[SPAN_HEADER];
    if (fseek(f, 0, SEEK_END) != 0) return SG_ERR_INVALID;
    si.file_size = (uint64_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    if (si.file_size < SPAN_HEADER || fread(h, 1, SPAN_HEADER, f) != SPAN_HEADER || memcmp(h, "SPAN", 4) != 0) {
        sg::set_error("not a SPAN file (bad magic or shorter than the 128-byte header)");
        return SG_ERR_INVALID;
    }
    if (rd32(h + 4) != 1) {
        sg::set_error("unsupported SPAN version %u", rd32(h + 4));
        return SG_ERR_UNSUPPORTED;
    }
    si.num_vectors = rd64(h + 8);
    si.n_part = rd32(h + 16);
    si.dim = rd32(h + 20);
    si.pq_enabled = h[24];
    si.pq_subvectors = rd32(h + 25);
    si.metric = h[29];
    si.checksum = rd64(h + 30);
    si.centroids_off = rd64(h + 38);
    si.codebook_off = rd64(h + 46);
    si.index_off = rd64(h + 54);
    si.data_off = rd64(h + 62);
    if (si.pq_enabled != 1) {
        sg::set_error("SPAN file without PQ: posting lists store only PQ codes, such an index cannot be searched (spann.rs:369-379)");
        return SG_ERR_UNSUPPORTED;
    }
    // every section must lie inside the file; products are formed in 128 bits so a hostile header cannot wrap them
    auto fits = [&](uint64_t off, unsigned __int128 bytes) { return off <= si.file_size && bytes <= (unsigned __int128)(si.file_size - off); };
    if (!fits(si.centroids_off, (unsigned __int128)si.n_part * si.dim * 4) || !fits(si.codebook_off, 12) ||
        !fits(si.index_off, (unsigned __int128)si.n_part * 12) || si.data_off > si.file_size || si.n_part > (1u << 24) ||
        si.dim == 0 || si.dim > 65536) {
        sg::set_error("SPAN file truncated (section offsets beyond %llu bytes)", (unsigned long long)si.file_size);
        return SG_ERR_INVALID;
    }
    uint8_t cb[12];
    fseek(f, (long)si.codebook_off, SEEK_SET);
    if (fread(cb, 1, 12, f) != 12) return SG_ERR_INVALID;
    si.m = rd32(cb);
    si.kc = rd32(cb + 4);
    si.ds = rd32(cb + 8);
    if (si.m == 0 || si.ds == 0 || si.kc == 0 || si.kc > 256 || (uint64_t)si.m * si.ds != si.dim ||
        !fits(si.codebook_off + 12, (unsigned __int128)si.m * si.kc * si.ds * 4) ||
        // a posting entry is a 4-byte id + m code bytes: num_vectors of them must fit behind data_off
        !fits(si.data_off, (unsigned __int128)si.num_vectors * (4ull + si.m)) || si.num_vectors > 0xFFFFFFFEull) {
        sg::set_error("SPAN codebook section is inconsistent (%u subvectors x %u centroids x %u dims for dimension %u)", si.m,
                      si.kc, si.ds, si.dim);
        return SG_ERR_INVALID;
    }
    if (verify) {
        std::vector<uint8_t> buf(1 << 20);
        fseek(f, (long)SPAN_HEADER, SEEK_SET);
        uint64_t hash = 0xcbf29ce484222325ull;
        size_t got;
        while ((got = fread(buf.data(), 1, buf.size(), f)) > 0) hash = fnv1a(hash, buf.data(), got);
        if (hash != si.checksum) {
            sg::set_error("Checksum mismatch: stored=%llu, computed=%llu", (unsigned long long)si.checksum, (unsigned long long)hash);
            return SG_ERR_INVALID;
        }
    }
    return SG_OK;
}

}  // namespace

extern "C" {

static int sg_spann_file_inspect_impl(const char *path, int verify_checksum, uint64_t *num_vectors_out, uint32_t *num_partitions_out,
                          uint32_t *dimension_out, uint32_t *num_subvectors_out, uint32_t *num_centroids_out,
                          uint32_t *subvec_dim_out, uint32_t *metric_out);
int sg_spann_file_inspect(const char *path, int verify_checksum, uint64_t *num_vectors_out, uint32_t *num_partitions_out,
                          uint32_t *dimension_out, uint32_t *num_subvectors_out, uint32_t *num_centroids_out,
                          uint32_t *subvec_dim_out, uint32_t *metric_out)
{
    return sg::abi_guard("sg_spann_file_inspect", [&] { return sg_spann_file_inspect_impl(path, verify_checksum, num_vectors_out, num_partitions_out, dimension_out, num_subvectors_out, num_centroids_out, subvec_dim_out, metric_out); });
}

static int sg_spann_file_inspect_impl(const char *path, int verify_checksum, uint64_t *num_vectors_out, uint32_t *num_partitions_out,
                          uint32_t *dimension_out, uint32_t *num_subvectors_out, uint32_t *num_centroids_out,
                          uint32_t *subvec_dim_out, uint32_t *metric_out)
{
    if (!path) return SG_ERR_INVALID;
    FILE *f = fopen(path, "rb");
    if (!f) {
        sg::set_error("Index file not found: %s", path);
        return SG_ERR_INVALID;
    }
    SpanInfo si;
    const int rc = span_parse_header(f, si, verify_checksum);
    fclose(f);
    if (rc != SG_OK) return rc;
    if (num_vectors_out) *num_vectors_out = si.num_vectors;
    if (num_partitions_out) *num_partitions_out = si.n_part;
    if (dimension_out) *dimension_out = si.dim;
    if (num_subvectors_out) *num_subvectors_out = si.m;
    if (num_centroids_out) *num_centroids_out = si.kc;
    if (subvec_dim_out) *subvec_dim_out = si.ds;
    if (metric_out) *metric_out = si.metric;
    return SG_OK;
}

static int sg_spann_file_read_impl(const char *path, float *centroids_out, float *pq_centroids_out, uint64_t *part_off_out, uint32_t *ids_out,
                       uint8_t *codes_out);
int sg_spann_file_read(const char *path, float *centroids_out, float *pq_centroids_out, uint64_t *part_off_out, uint32_t *ids_out,
                       uint8_t *codes_out)
{
    return sg::abi_guard("sg_spann_file_read", [&] { return sg_spann_file_read_impl(path, centroids_out, pq_centroids_out, part_off_out, ids_out, codes_out); });
}

static int sg_spann_file_read_impl(const char *path, float *centroids_out, float *pq_centroids_out, uint64_t *part_off_out, uint32_t *ids_out,
                       uint8_t *codes_out)
{
    if (!path || !centroids_out || !pq_centroids_out || !part_off_out) return SG_ERR_INVALID;
    FILE *f = fopen(path, "rb");
    if (!f) {
        sg::set_error("Index file not found: %s", path);
        return SG_ERR_INVALID;
    }
    SpanInfo si;
    int rc = span_parse_header(f, si, 0);
    if (rc != SG_OK) {
        fclose(f);
        return rc;
    }
    bool ok = true;
    fseek(f, (long)si.centroids_off, SEEK_SET);
    ok = ok && fread(centroids_out, 4, (size_t)si.n_part * si.dim, f) == (size_t)si.n_part * si.dim;
    fseek(f, (long)si.codebook_off + 12, SEEK_SET);
    ok = ok && fread(pq_centroids_out, 4, (size_t)si.m * si.kc * si.ds, f) == (size_t)si.m * si.kc * si.ds;
    std::vector<uint8_t> index((size_t)si.n_part * 12);
    fseek(f, (long)si.index_off, SEEK_SET);
    ok = ok && fread(index.data(), 1, index.size(), f) == index.size();
    const size_t entry = 4 + (size_t)si.m;
    uint64_t at = 0;
    std::vector<uint8_t> buf;
    for (uint32_t p = 0; ok && p < si.n_part; ++p) {
        const uint64_t rel = rd64(index.data() + (size_t)p * 12);
        const uint32_t cnt = rd32(index.data() + (size_t)p * 12 + 8);
        part_off_out[p] = at;
        if (cnt == 0) continue;
        if (si.data_off + rel + (uint64_t)cnt * entry > si.file_size || at + cnt > si.num_vectors || !ids_out || !codes_out) {
            sg::set_error("SPAN posting list %u is out of bounds (offset %llu, %u entries)", p, (unsigned long long)rel, cnt);
            fclose(f);
            return SG_ERR_INVALID;
        }
        buf.resize((size_t)cnt * entry);
        fseek(f, (long)(si.data_off + rel), SEEK_SET);
        ok = fread(buf.data(), 1, buf.size(), f) == buf.size();
        for (uint32_t e = 0; ok && e < cnt; ++e) {
            memcpy(ids_out + at + e, buf.data() + (size_t)e * entry, 4);
            memcpy(codes_out + (at + e) * si.m, buf.data() + (size_t)e * entry + 4, si.m);
        }
        at += cnt;
    }
    fclose(f);
    if (!ok) {
        sg::set_error("short read in the SPAN file");
        return SG_ERR_INVALID;
    }
    part_off_out[si.n_part] = at;
    if (at != si.num_vectors) {
        sg::set_error("SPAN posting lists hold %llu entries, the header says %llu", (unsigned long long)at,
                      (unsigned long long)si.num_vectors);
        return SG_ERR_INVALID;
    }
    return SG_OK;
}

static int sg_spann_file_write_impl(const char *path, uint32_t dim, uint32_t num_partitions, uint32_t metric, const float *centroids,
                        uint32_t num_subvectors, uint32_t num_centroids, uint32_t subvec_dim, const float *pq_centroids,
                        const uint64_t *part_off, const uint32_t *ids, const uint8_t *codes);
int sg_spann_file_write(const char *path, uint32_t dim, uint32_t num_partitions, uint32_t metric, const float *centroids,
                        uint32_t num_subvectors, uint32_t num_centroids, uint32_t subvec_dim, const float *pq_centroids,
                        const uint64_t *part_off, const uint32_t *ids, const uint8_t *codes)
{
    return sg::abi_guard("sg_spann_file_write", [&] { return sg_spann_file_write_impl(path, dim, num_partitions, metric, centroids, num_subvectors, num_centroids, subvec_dim, pq_centroids, part_off, ids, codes); });
}

static int sg_spann_file_write_impl(const char *path, uint32_t dim, uint32_t num_partitions, uint32_t metric, const float *centroids,
                        uint32_t num_subvectors, uint32_t num_centroids, uint32_t subvec_dim, const float *pq_centroids,
                        const uint64_t *part_off, const uint32_t *ids, const uint8_t *codes)
{
    if (!path || !centroids || !pq_centroids || !part_off) return SG_ERR_INVALID;
    if (num_partitions == 0) {
        sg::set_error("Cannot save empty index");                               // spann.rs:757-759
        return SG_ERR_INVALID;
    }
    if (num_centroids != 256 || subvec_dim != 8 || num_subvectors * subvec_dim != dim || metric > 2) {
        sg::set_error("the SPAN v1 layout fixes 256 centroids of 8 dims per subvector (got %u x %u for dimension %u)", num_centroids,
                      subvec_dim, dim);
        return SG_ERR_UNSUPPORTED;
    }
    const uint64_t n = part_off[num_partitions];
    if (n && (!ids || !codes)) return SG_ERR_INVALID;
    const uint64_t centroids_off = align64(SPAN_HEADER), centroids_size = (uint64_t)num_partitions * dim * 4;
    const uint64_t codebook_off = align64(centroids_off + centroids_size);
    const uint64_t codebook_size = 12 + (uint64_t)num_subvectors * 256 * 8 * 4;
    const uint64_t index_off = align64(codebook_off + codebook_size);
    const uint64_t data_off = align64(index_off + (uint64_t)num_partitions * 12);
    const size_t entry = 4 + (size_t)num_subvectors;
    FILE *f = fopen(path, "wb+");
    if (!f) {
        sg::set_error("cannot create %s", path);
        return SG_ERR_INVALID;
    }
    uint64_t hash = 0xcbf29ce484222325ull, off = SPAN_HEADER;
    bool ok = true;
    auto put = [&](const void *p, size_t bytes) {
        hash = fnv1a(hash, p, bytes);
        ok = ok && fwrite(p, 1, bytes, f) == bytes;
        off += bytes;
    };
    auto pad_to = [&](uint64_t target) {
        static const uint8_t zeros[64] = {0};
        while (off < target) put(zeros, (size_t)(target - off < 64 ? target - off : 64));
    };
    uint8_t h[SPAN_HEADER];
    memset(h, 0, sizeof(h));
    ok = fwrite(h, 1, SPAN_HEADER, f) == SPAN_HEADER;
    pad_to(centroids_off);
    put(centroids, (size_t)centroids_size);
    pad_to(codebook_off);
    const uint32_t cb[3] = {num_subvectors, 256u, 8u};
    put(cb, 12);
    put(pq_centroids, (size_t)num_subvectors * 256 * 8 * 4);
    pad_to(index_off);
    uint64_t rel = 0;
    for (uint32_t p = 0; p < num_partitions; ++p) {
        const uint32_t cnt = (uint32_t)(part_off[p + 1] - part_off[p]);
        uint8_t ie[12];
        memcpy(ie, &rel, 8);
        memcpy(ie + 8, &cnt, 4);
        put(ie, 12);
        rel += (uint64_t)cnt * entry;
    }
    pad_to(data_off);
    std::vector<uint8_t> buf;
    for (uint32_t p = 0; p < num_partitions; ++p) {
        const uint64_t lo = part_off[p], hi = part_off[p + 1];
        buf.resize((size_t)(hi - lo) * entry);
        for (uint64_t e = lo; e < hi; ++e) {
            memcpy(buf.data() + (size_t)(e - lo) * entry, ids + e, 4);
            memcpy(buf.data() + (size_t)(e - lo) * entry + 4, codes + e * num_subvectors, num_subvectors);
        }
        if (!buf.empty()) put(buf.data(), buf.size());
    }
    memcpy(h, "SPAN", 4);
    const uint32_t version = 1;
    memcpy(h + 4, &version, 4);
    memcpy(h + 8, &n, 8);
    memcpy(h + 16, &num_partitions, 4);
    memcpy(h + 20, &dim, 4);
    h[24] = 1;
    memcpy(h + 25, &num_subvectors, 4);
    h[29] = (uint8_t)metric;
    memcpy(h + 30, &hash, 8);
    memcpy(h + 38, &centroids_off, 8);
    memcpy(h + 46, &codebook_off, 8);
    memcpy(h + 54, &index_off, 8);
    memcpy(h + 62, &data_off, 8);
    ok = ok && fseek(f, 0, SEEK_SET) == 0 && fwrite(h, 1, SPAN_HEADER, f) == SPAN_HEADER;
    ok = (fclose(f) == 0) && ok;
    if (!ok) {
        sg::set_error("writing %s failed", path);
        return SG_ERR_INVALID;
    }
    return SG_OK;
}

static int sg_spann_load_file_impl(sg_spann **out, int device, const char *path, uint32_t num_probes, int verify_checksum);
int sg_spann_load_file(sg_spann **out, int device, const char *path, uint32_t num_probes, int verify_checksum)
{
    return sg::abi_guard("sg_spann_load_file", [&] { return sg_spann_load_file_impl(out, device, path, num_probes, verify_checksum); });
}

static int sg_spann_load_file_impl(sg_spann **out, int device, const char *path, uint32_t num_probes, int verify_checksum)
{
    if (!out || !path) return SG_ERR_INVALID;
    *out = nullptr;
    uint64_t n = 0;
    uint32_t P = 0, dim = 0, m = 0, kc = 0, ds = 0, metric = 0;
    int rc = sg_spann_file_inspect(path, verify_checksum, &n, &P, &dim, &m, &kc, &ds, &metric);
    if (rc != SG_OK) return rc;
    std::vector<float> cent((size_t)P * dim), pqc((size_t)m * kc * ds);
    std::vector<uint64_t> off((size_t)P + 1);
    std::vector<uint32_t> ids((size_t)(n ? n : 1));
    std::vector<uint8_t> codes((size_t)(n ? n : 1) * m);
    if ((rc = sg_spann_file_read(path, cent.data(), pqc.data(), off.data(), ids.data(), codes.data())) != SG_OK) return rc;
    return sg_spann_create(out, device, dim, P, num_probes ? num_probes : 10, (int)metric, cent.data(), pqc.data(), ds, kc,
                           off.data(), ids.data(), codes.data());
}

static int sg_spann_save_file_impl(sg_spann *s, const char *path);
int sg_spann_save_file(sg_spann *s, const char *path)
{
    return sg::abi_guard("sg_spann_save_file", [&] { return sg_spann_save_file_impl(s, path); });
}

static int sg_spann_save_file_impl(sg_spann *s, const char *path)
{
    if (!s || !path) return SG_ERR_INVALID;
    uint32_t P = 0;
    uint64_t n = 0;
    int rc = sg_spann_info(s, &P, &n, nullptr);
    if (rc != SG_OK) return rc;
    uint32_t dim = 0, m = 0, kc = 0, ds = 0;
    int metric = 0;
    if ((rc = sg_spann_shape(s, &dim, &m, &kc, &ds, &metric)) != SG_OK) return rc;
    std::vector<float> cent((size_t)P * dim), pqc((size_t)m * kc * ds);
    std::vector<uint64_t> off((size_t)P + 1);
    std::vector<uint32_t> ids((size_t)(n ? n : 1));
    std::vector<uint8_t> codes((size_t)(n ? n : 1) * m);
    if ((rc = sg_spann_get(s, cent.data(), pqc.data(), off.data(), ids.data(), codes.data())) != SG_OK) return rc;
    return sg_spann_file_write(path, dim, P, (uint32_t)metric, cent.data(), m, kc, ds, pqc.data(), off.data(), ids.data(),
                               codes.data());
}

}  // extern "C"
