// spann.cu -- the SPANN partitioned index on the device (the layer above pq.cu).
//
// Reference: src/vector_db/spann.rs -- SpannIndex::{build :363-463, kmeans_cluster :466-542,
// find_nearest_centroid :545-558, compute_distance :562-572, search :574-693, insert :1006-1051}.
// An index is: P partition centroids (k-means over the full vectors, P = ceil(sqrt(n))), a product quantizer
// (pq.rs), and per partition a posting list of (vector id, 48-byte PQ code).  A search ranks the partitions by
// compute_distance(query, centroid) with ties by index, probes the first num_probes, scores every entry of
// those lists with the ADC table and keeps the k smallest (distance, id).
//
// spann_search_kernel: ONE CTA PER QUERY.  At n = 10M: P = 3163 centroids (4.9 MB, L2-resident) and
// num_probes x n / P ~ 32k entries (1.6 MB of codes) per query -- small enough that one CTA does the whole
// query without a second launch: (1) ADC table into shared memory (48 x 256 squared distances, pq.rs order),
// (2) one thread per centroid walks the 384 products left to right (`Iterator::sum`) -> u64 keys, bitonic
// sort in shared memory, (3) the probed lists stream through the same per-warp threshold / append / compact
// top-k as the other scans, (4) CTA merge, results written in final order.  Every f32 operation is the
// reference's, so probes, distances and ids are bit-identical to the CPU path.
//
// Build and insert are orchestrated on the host over device kernels (they are not the hot path): the
// k-means assignment runs on the device; the centroid update adds a cluster's members in VECTOR ORDER (the
// reference's `+=` order), which the host arranges with a stable counting sort of the assignments.
// Parity: tests/test_spann_gpu.py (centroids of both k-means, posting lists, probes, distances, ids == oracle).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "sg_internal.h"

using namespace sg;

struct sg_spann {
    int device = 0;
    uint32_t dim = 0, n_part = 0, num_probes = 10;
    int metric = SG_METRIC_NEG_DOT;
    sg_pq *pq = nullptr;                       // owns the quantizer (codebooks + encode)
    uint32_t m = 0, kc = 0, ds = 0;
    float *d_centroids = nullptr;              // [P][dim]
    float *d_pq_centroids = nullptr;           // [m][kc][ds] (copy for the search kernel)
    // posting lists: host mirror per partition (insert appends here), device CSR rebuilt when dirty
    std::vector<std::vector<uint32_t>> h_ids;
    std::vector<std::vector<uint8_t>> h_codes;
    bool dirty = true;
    uint64_t n = 0;
    uint64_t *d_off = nullptr;                 // [P+1]
    uint32_t *d_ids = nullptr;
    uint8_t *d_codes = nullptr;
    size_t ids_cap = 0, codes_cap = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    float *d_q = nullptr; size_t q_cap = 0;
    uint64_t *d_pkeys = nullptr; size_t pkeys_cap = 0;   // [b][p_pow2] partition keys of the batch in flight
    uint32_t *d_rids = nullptr; float *d_rdist = nullptr; uint32_t *d_rcnt = nullptr; size_t res_cap = 0, cnt_cap = 0;
};

namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr uint32_t SPANN_MAX_PART = 4096;

// spann.rs:562-572: Euclidean = sum (x-y)^2, otherwise 1 - sum x*y; f32 sums left to right
SG_DEVICE_INLINE float spann_distance(const float *a, const float *b, uint32_t dim, int metric)
{
    float s = 0.0f;
    if (metric == SG_METRIC_EUCLIDEAN) {
        for (uint32_t i = 0; i < dim; ++i) {
            const float d = __fsub_rn(a[i], b[i]);
            s = __fadd_rn(s, __fmul_rn(d, d));
        }
        return s;
    }
    for (uint32_t i = 0; i < dim; ++i) s = __fadd_rn(s, __fmul_rn(a[i], b[i]));
    return __fsub_rn(1.0f, s);
}

SG_DEVICE_INLINE float pq_sq_l2(const float *a, const float *b, uint32_t n)
{
    float s = 0.0f;
    for (uint32_t i = 0; i < n; ++i) {
        const float d = __fsub_rn(a[i], b[i]);
        s = __fadd_rn(s, __fmul_rn(d, d));
    }
    return s;
}

// find_nearest_centroid (:545-558) for n vectors: one thread per vector, centroids read through L2
__global__ void spann_assign_kernel(const float *vecs, uint64_t n, uint32_t dim, const float *centroids, uint32_t k, int metric,
                                    uint32_t *assign)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *v = vecs + i * dim;
    uint32_t best = 0;
    float best_d = 3.402823466e+38f;
    for (uint32_t c = 0; c < k; ++c) {
        const float d = spann_distance(v, centroids + (size_t)c * dim, dim, metric);
        if (d < best_d) {
            best_d = d;
            best = c;
        }
    }
    assign[i] = best;
}

// centroid update (:500-520): members [moff[c], moff[c+1]) of cluster c are listed in vector order; one thread
// per (cluster, component) adds them in that order; empty clusters keep their centroid
__global__ void spann_update_kernel(const float *vecs, uint32_t dim, uint32_t k, const uint64_t *moff, const uint32_t *members,
                                    float *centroids)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)k * dim) return;
    const uint32_t c = (uint32_t)(t / dim), j = (uint32_t)(t % dim);
    const uint64_t lo = moff[c], hi = moff[c + 1];
    if (hi == lo) return;
    float acc = 0.0f;
    for (uint64_t e = lo; e < hi; ++e) acc = __fadd_rn(acc, vecs[(size_t)members[e] * dim + j]);
    centroids[t] = __fdiv_rn(acc, (float)(hi - lo));
}

// ---- partition ranking for a whole batch (spann.rs:597-609) -------------------------------------------------
// compute_distance(query, centroid) for every (query, centroid) pair: the reference's sum is strictly left to right,
// so one thread owns one pair's chain -- but the OPERANDS are staged: a CTA loads 32 centroids into shared memory
// with coalesced reads (row pitch dim + 1: thread-per-row walks are conflict-free), the queries come in groups of
// 16 (two per thread: two independent chains), and each warp's queries are shared-memory broadcasts.  (The first
// version had every query's CTA walk all P centroids straight from global memory, one row per thread: 1536-byte
// strides, a 32-byte sector per 4 bytes used -- ~39 MB of L2 traffic per query, 0.9 ms per query.)
// Output: keys[q][p_pow2] = (order_key32(distance) << 32 | p), KEY_SENTINEL padded.
constexpr uint32_t SP_CT = 32;      // centroids per CTA
constexpr uint32_t SP_QG = 16;      // queries per pass (8 warps x 2 chains)
__global__ void __launch_bounds__(256)
spann_rank_kernel(const float *__restrict__ queries, uint32_t b, uint32_t dim, const float *__restrict__ centroids, uint32_t n_part,
                  uint32_t p_pow2, int metric, uint64_t *__restrict__ keys)
{
    extern __shared__ __align__(16) float sp_sm[];
    float *cent = sp_sm;                                  // [SP_CT][dim + 1]
    float *qs = sp_sm + (size_t)SP_CT * (dim + 1);        // [SP_QG][dim]
    const uint32_t c0 = blockIdx.x * SP_CT;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < SP_CT * dim; i += blockDim.x) {
        const uint32_t c = i / dim, d = i - c * dim;
        cent[(size_t)c * (dim + 1) + d] = c0 + c < n_part ? centroids[(size_t)(c0 + c) * dim + d] : 0.0f;
    }
    const uint32_t p = c0 + lane;
    const float *cr = cent + (size_t)lane * (dim + 1);
    for (uint32_t q0 = 0; q0 < b; q0 += SP_QG) {
        __syncthreads();                                  // previous pass done with qs (and, first time, cent is complete)
        for (uint32_t i = threadIdx.x; i < SP_QG * dim; i += blockDim.x) {
            const uint32_t q = i / dim, d = i - q * dim;
            qs[i] = q0 + q < b ? queries[(size_t)(q0 + q) * dim + d] : 0.0f;
        }
        __syncthreads();
        const float *qa = qs + (size_t)(2 * warp) * dim, *qb = qa + dim;
        float sa = 0.0f, sb = 0.0f;                       // two independent left-to-right chains
        // operands are fetched 8 at a time ahead of the two add chains (the chains are the only true dependency:
        // without the explicit batching every step exposed a full shared-memory round trip, ~100 cycles per element)
        const uint32_t dim8 = dim & ~7u;
        if (metric == SG_METRIC_EUCLIDEAN) {
            for (uint32_t i0 = 0; i0 < dim8; i0 += 8) {
                float x[8], ya[8], yb[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    x[u] = cr[i0 + u];
                    ya[u] = qa[i0 + u];
                    yb[u] = qb[i0 + u];
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float da = __fsub_rn(ya[u], x[u]), db = __fsub_rn(yb[u], x[u]);
                    sa = __fadd_rn(sa, __fmul_rn(da, da));
                    sb = __fadd_rn(sb, __fmul_rn(db, db));
                }
            }
            for (uint32_t i = dim8; i < dim; ++i) {
                const float da = __fsub_rn(qa[i], cr[i]), db = __fsub_rn(qb[i], cr[i]);
                sa = __fadd_rn(sa, __fmul_rn(da, da));
                sb = __fadd_rn(sb, __fmul_rn(db, db));
            }
        } else {
            for (uint32_t i0 = 0; i0 < dim8; i0 += 8) {
                float x[8], ya[8], yb[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    x[u] = cr[i0 + u];
                    ya[u] = qa[i0 + u];
                    yb[u] = qb[i0 + u];
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    sa = __fadd_rn(sa, __fmul_rn(ya[u], x[u]));
                    sb = __fadd_rn(sb, __fmul_rn(yb[u], x[u]));
                }
            }
            for (uint32_t i = dim8; i < dim; ++i) {
                sa = __fadd_rn(sa, __fmul_rn(qa[i], cr[i]));
                sb = __fadd_rn(sb, __fmul_rn(qb[i], cr[i]));
            }
            sa = __fsub_rn(1.0f, sa);
            sb = __fsub_rn(1.0f, sb);
        }
        if (p < n_part) {
            const uint32_t qi = q0 + 2 * warp;
            if (sa != sa) sa = __uint_as_float(0x7FC00000u);
            if (sb != sb) sb = __uint_as_float(0x7FC00000u);
            if (qi < b) keys[(size_t)qi * p_pow2 + p] = make_key(order_key32(sa), p);
            if (qi + 1 < b) keys[(size_t)(qi + 1) * p_pow2 + p] = make_key(order_key32(sb), p);
        }
    }
    // the pad of every query's key row (the sort wants a power of two)
    if (blockIdx.x == 0)
        for (uint32_t q = warp; q < b; q += blockDim.x / 32)
            for (uint32_t i = n_part + lane; i < p_pow2; i += 32) keys[(size_t)q * p_pow2 + i] = KEY_SENTINEL;
}

struct SpannSearchArgs {
    const float *queries;          // [b][dim]
    uint32_t dim, n_part, num_probes;
    int metric;
    const uint64_t *pkeys;         // [b][p_pow2] partition keys of the batch (spann_rank_kernel)
    const float *pq_centroids;     // [m][kc][ds]
    uint32_t m, kc, ds;
    const uint64_t *off;           // [P+1]
    const uint32_t *ids;
    const uint8_t *codes;          // [n][m]
    uint32_t topk, cap, p_pow2;
    uint32_t *ids_out; float *dist_out; uint32_t *counts_out;   // [b][topk], [b][topk], [b]
};

SG_DEVICE_INLINE void sp_compact(uint64_t *buf, uint32_t &cnt, uint32_t &tau, uint32_t k, uint32_t lane)
{
    uint32_t mm = next_pow2(cnt < 2 ? 2 : cnt);
    for (uint32_t i = cnt + lane; i < mm; i += 32) buf[i] = KEY_SENTINEL;
    __syncwarp();
    warp_bitonic_sort(buf, mm, lane);
    if (cnt > k) cnt = k;
    if (cnt == k) tau = (uint32_t)(buf[k - 1] >> 32);
}

__host__ __device__ inline size_t sp_align16(size_t x) { return (x + 15) & ~(size_t)15; }
// shared memory layout: region0 (ADC table, later the merge buffer) | partition keys | query | candidates | counts
__host__ __device__ inline size_t sp_region0(uint32_t m, uint32_t kc, uint32_t W, uint32_t topk)
{
    uint32_t M = 2;
    while (M < W * topk) M <<= 1;
    const size_t t = (size_t)m * kc * 4, mb = (size_t)M * 8;
    return sp_align16(t > mb ? t : mb);
}
size_t sp_smem(uint32_t m, uint32_t kc, uint32_t W, uint32_t topk, uint32_t p_pow2, uint32_t dim, uint32_t cap)
{
    return sp_region0(m, kc, W, topk) + (size_t)p_pow2 * 8 + sp_align16((size_t)dim * 4) + (size_t)W * cap * 8 + (size_t)W * 4 + 16;
}

__global__ void __launch_bounds__(256, 1)
spann_search_kernel(const SpannSearchArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const uint32_t q = blockIdx.x, m = a.m, kc = a.kc, C = a.cap, topk = a.topk;
    float *lut = reinterpret_cast<float *>(smem);
    uint64_t *pkeys = reinterpret_cast<uint64_t *>(smem + sp_region0(m, kc, W, topk));
    float *qs = reinterpret_cast<float *>(pkeys + a.p_pow2);
    uint64_t *cand = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(qs) + sp_align16((size_t)a.dim * 4));
    uint32_t *cnt_s = reinterpret_cast<uint32_t *>(cand + (size_t)W * C);

    for (uint32_t d = threadIdx.x; d < a.dim; d += blockDim.x) qs[d] = a.queries[(size_t)q * a.dim + d];
    __syncthreads();
    // (1) build_distance_table (pq.rs:329-353)
    for (uint32_t i = threadIdx.x; i < m * kc; i += blockDim.x) {
        const uint32_t s = i / kc;
        lut[i] = pq_sq_l2(qs + (size_t)s * a.ds, a.pq_centroids + (size_t)i * a.ds, a.ds);
    }
    // (2) partition ranking (:597-609): (distance total_cmp, index) keys come from spann_rank_kernel
    for (uint32_t p = threadIdx.x; p < a.p_pow2; p += blockDim.x) pkeys[p] = a.pkeys[(size_t)q * a.p_pow2 + p];
    __syncthreads();
    block_bitonic_sort(pkeys, a.p_pow2);
    const uint32_t n_probe = a.num_probes < a.n_part ? a.num_probes : a.n_part;

    // (3) ADC over the probed posting lists, in probe order (:626-650)
    uint32_t cnt = 0, tau = 0xFFFFFFFFu;
    uint64_t *buf = cand + (size_t)warp * C;
    const uint32_t words = m / 4;
    const uint32_t *codes32 = reinterpret_cast<const uint32_t *>(a.codes);
    for (uint32_t t = 0; t < n_probe; ++t) {
        const uint32_t p = (uint32_t)pkeys[t];
        const uint64_t lo = a.off[p], hi = a.off[p + 1];
        for (uint64_t base = lo + (uint64_t)warp * 32; base < hi; base += (uint64_t)W * 32) {
            const uint64_t e = base + lane;
            const bool live = e < hi;
            float total = 0.0f;
            bool bad = false;
            uint32_t id = 0;
            if (live) {
                id = a.ids[e];
                auto add4 = [&](uint32_t four, uint32_t w) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t code = (four >> (8 * u)) & 0xFFu;
                        bad |= code >= kc;
                        total = __fadd_rn(total, lut[(w * 4 + u) * kc + (code < kc ? code : 0)]);
                    }
                };
                if ((m & 15u) == 0) {             // 16-byte loads: a lane's 48-byte code is 3 of them (word loads touch every
                    const uint4 *c4 = reinterpret_cast<const uint4 *>(a.codes + e * m);   // sector 8 times with no L1 to absorb it)
                    for (uint32_t w4 = 0; w4 < m / 16; ++w4) {
                        const uint4 v = __ldg(c4 + w4);
                        add4(v.x, 4 * w4);
                        add4(v.y, 4 * w4 + 1);
                        add4(v.z, 4 * w4 + 2);
                        add4(v.w, 4 * w4 + 3);
                    }
                } else {
                    const uint32_t *cw = codes32 + e * words;
                    for (uint32_t w = 0; w < words; ++w) add4(__ldg(cw + w), w);
                }
            }
            const float d = bad ? 3.402823466e+38f : total;
            const uint32_t key32 = order_key32(d != d ? __uint_as_float(0x7FC00000u) : d);
            if (cnt + 32 > C) sp_compact(buf, cnt, tau, topk, lane);
            const bool hit = live && key32 <= tau;
            const uint32_t mk = __ballot_sync(FULL, hit);
            if (mk) {
                if (hit) buf[cnt + __popc(mk & ((1u << lane) - 1))] = make_key(key32, id);
                cnt += __popc(mk);
                __syncwarp();
            }
        }
    }
    sp_compact(buf, cnt, tau, topk, lane);
    if (lane == 0) cnt_s[warp] = cnt;
    __syncthreads();
    // (4) CTA merge (the table is no longer needed: sort in its place), results in (distance, id) order
    uint64_t *mbuf = reinterpret_cast<uint64_t *>(lut);
    const uint32_t M = next_pow2(W * topk < 2 ? 2 : W * topk);
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
        const uint32_t w = i / topk, r = i - w * topk;
        mbuf[i] = (w < W && r < cnt_s[w]) ? cand[(size_t)w * C + r] : KEY_SENTINEL;
    }
    __syncthreads();
    block_bitonic_sort(mbuf, M);
    uint32_t found = 0;
    for (uint32_t w = 0; w < W; ++w) found += cnt_s[w];
    if (found > topk) found = topk;
    for (uint32_t i = threadIdx.x; i < topk; i += blockDim.x) {
        const uint64_t key = mbuf[i];
        const bool ok = i < found && key != KEY_SENTINEL;
        a.ids_out[(size_t)q * topk + i] = ok ? (uint32_t)key : 0xFFFFFFFFu;
        a.dist_out[(size_t)q * topk + i] = ok ? key32_to_dist((uint32_t)(key >> 32)) : __uint_as_float(0x7F800000u);
    }
    if (threadIdx.x == 0) a.counts_out[q] = found;
}

template <typename T>
int sp_grow(T *&ptr, size_t &cap, size_t need)
{
    if (need <= cap) return SG_OK;
    size_t ncap = need + need / 2;
    T *np = nullptr;
    SG_CUDA_TRY(cudaMalloc(&np, ncap * sizeof(T)));
    if (ptr) cudaFree(ptr);
    ptr = np;
    cap = ncap;
    return SG_OK;
}

uint32_t pow2_ge(uint32_t x)
{
    uint32_t p = 2;
    while (p < x) p <<= 1;
    return p;
}

// device CSR from the host posting lists
int sp_upload_lists(sg_spann *s)
{
    if (!s->dirty) return SG_OK;
    const uint32_t P = s->n_part;
    std::vector<uint64_t> off(P + 1, 0);
    for (uint32_t p = 0; p < P; ++p) off[p + 1] = off[p] + s->h_ids[p].size();
    const uint64_t n = off[P];
    int rc = sp_grow(s->d_ids, s->ids_cap, (size_t)(n ? n : 1));
    if (rc == SG_OK) rc = sp_grow(s->d_codes, s->codes_cap, (size_t)(n ? n : 1) * s->m);
    if (rc != SG_OK) return rc;
    if (!s->d_off) SG_CUDA_TRY(cudaMalloc(&s->d_off, (size_t)(P + 1) * 8));
    SG_CUDA_TRY(cudaMemcpy(s->d_off, off.data(), (size_t)(P + 1) * 8, cudaMemcpyHostToDevice));
    for (uint32_t p = 0; p < P; ++p) {
        const size_t cntp = s->h_ids[p].size();
        if (!cntp) continue;
        SG_CUDA_TRY(cudaMemcpy(s->d_ids + off[p], s->h_ids[p].data(), cntp * 4, cudaMemcpyHostToDevice));
        SG_CUDA_TRY(cudaMemcpy(s->d_codes + off[p] * s->m, s->h_codes[p].data(), cntp * s->m, cudaMemcpyHostToDevice));
    }
    s->n = n;
    s->dirty = false;
    return SG_OK;
}

int sp_alloc(sg_spann **out, int device, uint32_t dim, uint32_t n_part, uint32_t num_probes, int metric)
{
    if (n_part == 0 || n_part > SPANN_MAX_PART) {
        set_error("num_partitions=%u outside 1..%u", n_part, SPANN_MAX_PART);
        return SG_ERR_UNSUPPORTED;
    }
    if (metric < SG_METRIC_NEG_DOT || metric > SG_METRIC_COSINE) {
        set_error("unknown metric %d", metric);
        return SG_ERR_INVALID;
    }
    {
        int ndev = 0;
        cudaError_t ce = cudaGetDeviceCount(&ndev);
        if (ce != cudaSuccess || ndev == 0) {
            set_error("no CUDA device available (%s); libshodh_gpu has no CPU fallback",
                      ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce));
            return SG_ERR_NO_DEVICE;
        }
        if (device < 0 || device >= ndev) {
            set_error("device ordinal %d out of range (0..%d)", device, ndev - 1);
            return SG_ERR_INVALID;
        }
    }
    sg_spann *s = new (std::nothrow) sg_spann();
    if (!s) return SG_ERR_OOM;
    s->device = device;
    s->dim = dim;
    s->n_part = n_part;
    s->num_probes = num_probes ? num_probes : 10;
    s->metric = metric;
    s->h_ids.resize(n_part);
    s->h_codes.resize(n_part);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_centroids, (size_t)n_part * dim * 4);
    if (e != cudaSuccess) {
        set_error("SPANN allocation failed: %s", cudaGetErrorString(e));
        if (s->stream) cudaStreamDestroy(s->stream);
        delete s;
        return SG_ERR_CUDA;
    }
    *out = s;
    return SG_OK;
}

// adopt a trained quantizer: keep the handle (encode) and a device copy of its codebooks (search kernel)
int sp_attach_pq(sg_spann *s, sg_pq *pq)
{
    uint32_t dim = 0, m = 0, kc = 0;
    int rc = sg_pq_info(pq, &dim, &m, &kc, nullptr);
    if (rc != SG_OK) return rc;
    if (dim != s->dim) {
        set_error("the quantizer is %u-d, the index %u-d", dim, s->dim);
        return SG_ERR_DIM;
    }
    s->pq = pq;
    s->m = m;
    s->kc = kc;
    s->ds = dim / m;
    std::vector<float> c((size_t)m * kc * s->ds);
    if ((rc = sg_pq_get_centroids(pq, c.data())) != SG_OK) return rc;
    SG_CUDA_TRY(cudaSetDevice(s->device));
    SG_CUDA_TRY(cudaMalloc(&s->d_pq_centroids, c.size() * 4));
    SG_CUDA_TRY(cudaMemcpy(s->d_pq_centroids, c.data(), c.size() * 4, cudaMemcpyHostToDevice));
    return SG_OK;
}

}  // namespace

extern "C" {

void sg_spann_destroy(sg_spann *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->pq) sg_pq_destroy(s->pq);
    cudaFree(s->d_centroids);
    cudaFree(s->d_pq_centroids);
    cudaFree(s->d_off);
    cudaFree(s->d_ids);
    cudaFree(s->d_codes);
    cudaFree(s->d_q);
    cudaFree(s->d_pkeys);
    cudaFree(s->d_rids);
    cudaFree(s->d_rdist);
    cudaFree(s->d_rcnt);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int sg_spann_build(sg_spann **out, int device, uint32_t dim, uint32_t num_partitions, uint32_t num_probes,
                   uint32_t kmeans_iterations, int metric, const float *vectors, uint64_t n, const uint32_t *order,
                   uint32_t pq_iterations, const uint32_t *pq_orders)
{
    if (!out || !order || !pq_orders || (n && !vectors)) return SG_ERR_INVALID;
    *out = nullptr;
    if (n == 0) {
        set_error("Cannot build index from empty vectors");                   // spann.rs:364-366
        return SG_ERR_INVALID;
    }
    const uint32_t P = num_partitions ? num_partitions : (uint32_t)std::max<double>(std::ceil(std::sqrt((double)n)), 1.0);   // :136-139
    sg_spann *s = nullptr;
    int rc = sp_alloc(&s, device, dim, P, num_probes, metric);
    if (rc != SG_OK) return rc;
    float *d_vecs = nullptr;
    uint32_t *d_assign = nullptr, *d_members = nullptr;
    uint64_t *d_moff = nullptr;
    auto fail = [&](int code) {
        cudaFree(d_vecs);
        cudaFree(d_assign);
        cudaFree(d_members);
        cudaFree(d_moff);
        sg_spann_destroy(s);
        return code;
    };
    cudaError_t e = cudaMalloc(&d_vecs, (size_t)n * dim * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_assign, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_members, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_moff, (size_t)(P + 1) * 8);
    if (e == cudaSuccess) e = cudaMemcpy(d_vecs, vectors, (size_t)n * dim * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("SPANN build allocation failed: %s", cudaGetErrorString(e));
        return fail(e == cudaErrorMemoryAllocation ? SG_ERR_OOM : SG_ERR_CUDA);
    }
    // Step 1 (:402, kmeans_cluster): initial centroids = the first P vectors of the shuffled order
    {
        std::vector<float> init((size_t)P * dim);
        for (uint32_t c = 0; c < P; ++c) {
            const uint64_t src = order[c < n ? c : c % n];
            memcpy(init.data() + (size_t)c * dim, vectors + src * dim, (size_t)dim * 4);
        }
        e = cudaMemcpy(s->d_centroids, init.data(), init.size() * 4, cudaMemcpyHostToDevice);
    }
    std::vector<uint32_t> assign(n, 0), prev(n, 0), members(n);
    std::vector<uint64_t> moff(P + 1);
    auto run_assign = [&]() -> cudaError_t {
        spann_assign_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s->stream>>>(d_vecs, n, dim, s->d_centroids, P, metric, d_assign);
        count_launch();
        cudaError_t r = cudaGetLastError();
        if (r == cudaSuccess) r = cudaMemcpyAsync(assign.data(), d_assign, (size_t)n * 4, cudaMemcpyDeviceToHost, s->stream);
        if (r == cudaSuccess) r = cudaStreamSynchronize(s->stream);
        return r;
    };
    for (uint32_t it = 0; it < kmeans_iterations && e == cudaSuccess; ++it) {
        if ((e = run_assign()) != cudaSuccess) break;
        uint64_t changed = 0;
        for (uint64_t i = 0; i < n; ++i) changed += assign[i] != prev[i];
        prev = assign;
        // members of every cluster in vector order (stable counting sort), then the in-order sums on the device
        std::fill(moff.begin(), moff.end(), 0);
        for (uint64_t i = 0; i < n; ++i) ++moff[assign[i] + 1];
        for (uint32_t c = 0; c < P; ++c) moff[c + 1] += moff[c];
        {
            std::vector<uint64_t> cur(moff.begin(), moff.end() - 1);
            for (uint64_t i = 0; i < n; ++i) members[cur[assign[i]]++] = (uint32_t)i;
        }
        e = cudaMemcpyAsync(d_moff, moff.data(), (size_t)(P + 1) * 8, cudaMemcpyHostToDevice, s->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_members, members.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s->stream);
        if (e == cudaSuccess) {
            const uint64_t tot = (uint64_t)P * dim;
            spann_update_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, s->stream>>>(d_vecs, dim, P, d_moff, d_members, s->d_centroids);
            count_launch();
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (changed == 0) break;                                                   // :535-538
    }
    // Step 3 (:412-432): every vector goes to its nearest FINAL centroid, entries in vector order
    if (e == cudaSuccess) e = run_assign();
    if (e != cudaSuccess) {
        set_error("SPANN build failed: %s", cudaGetErrorString(e));
        return fail(SG_ERR_CUDA);
    }
    cudaFree(d_vecs);
    cudaFree(d_assign);
    cudaFree(d_members);
    cudaFree(d_moff);
    d_vecs = nullptr;
    d_assign = d_members = nullptr;
    d_moff = nullptr;
    // Step 2 (:405-410): the quantizer, trained on all vectors; then the codes
    sg_pq *pq = nullptr;
    rc = sg_pq_train(&pq, device, dim, 8, 256, pq_iterations, vectors, n, pq_orders);
    if (rc == SG_OK) rc = sp_attach_pq(s, pq);
    if (rc != SG_OK) {
        if (pq && !s->pq) sg_pq_destroy(pq);
        return fail(rc);
    }
    std::vector<uint8_t> codes((size_t)n * s->m);
    if ((rc = sg_pq_encode(pq, vectors, n, codes.data())) != SG_OK) return fail(rc);
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t p = assign[i];
        s->h_ids[p].push_back((uint32_t)i);
        s->h_codes[p].insert(s->h_codes[p].end(), codes.begin() + i * s->m, codes.begin() + (i + 1) * s->m);
    }
    s->dirty = true;
    if ((rc = sp_upload_lists(s)) != SG_OK) return fail(rc);
    *out = s;
    return SG_OK;
}

int sg_spann_create(sg_spann **out, int device, uint32_t dim, uint32_t num_partitions, uint32_t num_probes, int metric,
                    const float *centroids, const float *pq_centroids, uint32_t subvec_dim, uint32_t num_centroids,
                    const uint64_t *part_off, const uint32_t *ids, const uint8_t *codes)
{
    if (!out || !centroids || !pq_centroids || !part_off) return SG_ERR_INVALID;
    *out = nullptr;
    sg_spann *s = nullptr;
    int rc = sp_alloc(&s, device, dim, num_partitions, num_probes, metric);
    if (rc != SG_OK) return rc;
    sg_pq *pq = nullptr;
    rc = sg_pq_create(&pq, device, dim, subvec_dim, num_centroids, pq_centroids);
    if (rc == SG_OK) rc = sp_attach_pq(s, pq);
    if (rc != SG_OK) {
        if (pq && !s->pq) sg_pq_destroy(pq);
        sg_spann_destroy(s);
        return rc;
    }
    cudaError_t e = cudaMemcpy(s->d_centroids, centroids, (size_t)num_partitions * dim * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("centroid upload failed: %s", cudaGetErrorString(e));
        sg_spann_destroy(s);
        return SG_ERR_CUDA;
    }
    for (uint32_t p = 0; p < num_partitions; ++p) {
        const uint64_t lo = part_off[p], hi = part_off[p + 1];
        if (hi < lo || (hi > lo && (!ids || !codes))) {
            set_error("bad posting-list offsets at partition %u", p);
            sg_spann_destroy(s);
            return SG_ERR_INVALID;
        }
        s->h_ids[p].assign(ids + lo, ids + hi);
        s->h_codes[p].assign(codes + lo * s->m, codes + hi * s->m);
    }
    s->dirty = true;
    if ((rc = sp_upload_lists(s)) != SG_OK) {
        sg_spann_destroy(s);
        return rc;
    }
    *out = s;
    return SG_OK;
}

int sg_spann_info(sg_spann *s, uint32_t *num_partitions_out, uint64_t *len_out, uint32_t *num_probes_out)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    uint64_t n = 0;
    for (const auto &l : s->h_ids) n += l.size();
    if (num_partitions_out) *num_partitions_out = s->n_part;
    if (len_out) *len_out = n;
    if (num_probes_out) *num_probes_out = s->num_probes;
    return SG_OK;
}

int sg_spann_shape(sg_spann *s, uint32_t *dim_out, uint32_t *num_subvectors_out, uint32_t *num_centroids_out,
                   uint32_t *subvec_dim_out, int *metric_out)
{
    if (!s) return SG_ERR_INVALID;
    if (dim_out) *dim_out = s->dim;
    if (num_subvectors_out) *num_subvectors_out = s->m;
    if (num_centroids_out) *num_centroids_out = s->kc;
    if (subvec_dim_out) *subvec_dim_out = s->ds;
    if (metric_out) *metric_out = s->metric;
    return SG_OK;
}

int sg_spann_get(sg_spann *s, float *centroids_out, float *pq_centroids_out, uint64_t *part_off_out, uint32_t *ids_out,
                 uint8_t *codes_out)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    if (centroids_out) SG_CUDA_TRY(cudaMemcpy(centroids_out, s->d_centroids, (size_t)s->n_part * s->dim * 4, cudaMemcpyDeviceToHost));
    if (pq_centroids_out)
        SG_CUDA_TRY(cudaMemcpy(pq_centroids_out, s->d_pq_centroids, (size_t)s->m * s->kc * s->ds * 4, cudaMemcpyDeviceToHost));
    uint64_t at = 0;
    for (uint32_t p = 0; p < s->n_part; ++p) {
        if (part_off_out) part_off_out[p] = at;
        const size_t c = s->h_ids[p].size();
        if (ids_out && c) memcpy(ids_out + at, s->h_ids[p].data(), c * 4);
        if (codes_out && c) memcpy(codes_out + at * s->m, s->h_codes[p].data(), c * s->m);
        at += c;
    }
    if (part_off_out) part_off_out[s->n_part] = at;
    return SG_OK;
}

int sg_spann_insert(sg_spann *s, uint32_t vector_id, const float *vector)
{
    if (!s || !vector) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    int rc = sp_grow(s->d_q, s->q_cap, (size_t)s->dim);
    if (rc != SG_OK) return rc;
    if ((rc = sp_grow(s->d_rcnt, s->cnt_cap, (size_t)1)) != SG_OK) return rc;
    SG_CUDA_TRY(cudaMemcpyAsync(s->d_q, vector, (size_t)s->dim * 4, cudaMemcpyHostToDevice, s->stream));
    spann_assign_kernel<<<1, 128, 0, s->stream>>>(s->d_q, 1, s->dim, s->d_centroids, s->n_part, s->metric, s->d_rcnt);
    count_launch();
    SG_CUDA_TRY(cudaGetLastError());
    uint32_t p = 0;
    SG_CUDA_TRY(cudaMemcpyAsync(&p, s->d_rcnt, 4, cudaMemcpyDeviceToHost, s->stream));
    SG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    std::vector<uint8_t> code(s->m);
    if ((rc = sg_pq_encode(s->pq, vector, 1, code.data())) != SG_OK) return rc;
    s->h_ids[p].push_back(vector_id);
    s->h_codes[p].insert(s->h_codes[p].end(), code.begin(), code.end());
    s->dirty = true;
    return SG_OK;
}

int sg_spann_search(sg_spann *s, const float *queries, uint32_t b, uint32_t topk, uint32_t *ids_out, float *dist_out,
                    uint32_t *counts_out)
{
    if (!s || (b && (!queries || !ids_out || !dist_out || !counts_out))) return SG_ERR_INVALID;
    if (b == 0) return SG_OK;
    if (topk == 0 || topk > 512) {
        set_error("k=%u outside the supported range 1..512", topk);
        return topk == 0 ? SG_ERR_INVALID : SG_ERR_UNSUPPORTED;
    }
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    int rc = sp_upload_lists(s);
    if (rc != SG_OK) return rc;
    uint32_t kp = 32;
    while (kp < topk) kp <<= 1;
    const uint32_t cap = 2 * kp, p_pow2 = pow2_ge(s->n_part);
    uint32_t W = 8;
    while (W > 1 && sp_smem(s->m, s->kc, W, topk, p_pow2, s->dim, cap) > 220 * 1024) W >>= 1;
    const size_t smem = sp_smem(s->m, s->kc, W, topk, p_pow2, s->dim, cap);
    if (smem > 220 * 1024) {
        set_error("no SPANN search configuration fits shared memory (k=%u)", topk);
        return SG_ERR_UNSUPPORTED;
    }
    if ((rc = sp_grow(s->d_q, s->q_cap, (size_t)b * s->dim)) != SG_OK) return rc;
    if ((size_t)b * topk > s->res_cap) {
        size_t c1 = s->res_cap, c2 = s->res_cap;
        if ((rc = sp_grow(s->d_rids, c1, (size_t)b * topk)) != SG_OK) return rc;
        if ((rc = sp_grow(s->d_rdist, c2, (size_t)b * topk)) != SG_OK) return rc;
        s->res_cap = c1 < c2 ? c1 : c2;
    }
    if ((rc = sp_grow(s->d_rcnt, s->cnt_cap, (size_t)b)) != SG_OK) return rc;
    if ((rc = sp_grow(s->d_pkeys, s->pkeys_cap, (size_t)b * p_pow2)) != SG_OK) return rc;
    cudaStream_t st = s->stream;
    SG_CUDA_TRY(cudaMemcpyAsync(s->d_q, queries, (size_t)b * s->dim * 4, cudaMemcpyHostToDevice, st));
    {   // partition ranking of the whole batch
        const size_t rsm = ((size_t)SP_CT * (s->dim + 1) + (size_t)SP_QG * s->dim) * 4;
        static DevAttr da_rank;
        if (rsm > 48 * 1024) SG_CUDA_TRY(ensure_smem(spann_rank_kernel, da_rank, rsm));
        spann_rank_kernel<<<(s->n_part + SP_CT - 1) / SP_CT, 256, rsm, st>>>(s->d_q, b, s->dim, s->d_centroids, s->n_part, p_pow2,
                                                                            s->metric, s->d_pkeys);
        count_launch();
        SG_CUDA_TRY(cudaGetLastError());
    }
    SpannSearchArgs a;
    a.queries = s->d_q;
    a.dim = s->dim;
    a.n_part = s->n_part;
    a.num_probes = s->num_probes;
    a.metric = s->metric;
    a.pkeys = s->d_pkeys;
    a.pq_centroids = s->d_pq_centroids;
    a.m = s->m;
    a.kc = s->kc;
    a.ds = s->ds;
    a.off = s->d_off;
    a.ids = s->d_ids;
    a.codes = s->d_codes;
    a.topk = topk;
    a.cap = cap;
    a.p_pow2 = p_pow2;
    a.ids_out = s->d_rids;
    a.dist_out = s->d_rdist;
    a.counts_out = s->d_rcnt;
    static DevAttr da_search;
    SG_CUDA_TRY(ensure_smem(spann_search_kernel, da_search, smem));
    spann_search_kernel<<<b, W * 32, smem, st>>>(a);
    count_launch();
    SG_CUDA_TRY(cudaGetLastError());
    SG_CUDA_TRY(cudaMemcpyAsync(ids_out, s->d_rids, (size_t)b * topk * 4, cudaMemcpyDeviceToHost, st));
    SG_CUDA_TRY(cudaMemcpyAsync(dist_out, s->d_rdist, (size_t)b * topk * 4, cudaMemcpyDeviceToHost, st));
    SG_CUDA_TRY(cudaMemcpyAsync(counts_out, s->d_rcnt, (size_t)b * 4, cudaMemcpyDeviceToHost, st));
    SG_CUDA_TRY(cudaStreamSynchronize(st));
    return SG_OK;
}

int sg_spann_set_probes(sg_spann *s, uint32_t num_probes)
{
    if (!s || num_probes == 0) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    s->num_probes = num_probes;
    return SG_OK;
}

}  // extern "C"
