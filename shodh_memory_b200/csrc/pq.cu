// pq.cu -- Product Quantization on the device: the ADC scan over 48-byte codes.
//
// Reference: src/vector_db/pq.rs -- ProductQuantizer::{train/fit/kmeans :94-215, encode :220-257, decode
// :260-290, build_distance_table :329-353, distance_with_table :358-369} and
// CompressedVectorStore::{add :428-432, search :449-466}.  A 384-d vector becomes 48 one-byte codes (one of
// 256 centroids per 8-d subspace); a search builds a 48 x 256 table of squared L2 distances from the query's
// subvectors to every centroid and then scores every stored vector with 48 table look-ups:
//     total = 0.0; for m in 0..48 { total += table[m][code[m]] }      (f32, in subvector order)
// sorted by (distance total_cmp, id), truncated to k.  All of it is first-party arithmetic with a fixed order,
// so the device reproduces it bit for bit: squared_l2_distance is a left-to-right f32 sum of (x-y)*(x-y)
// (`powi(2)` is a multiply), k-means assigns with a strict `<` (first minimum wins) and averages in vector
// order, the scan adds its 48 terms in order.
//
// pq_scan_kernel.  Bound: HBM for the code stream (48 B per vector = 1/16 of the fp16 row) against 48
// shared-memory gathers per vector.  The query's table (48 KB) sits in shared memory; one thread owns one
// vector, reads its 48 bytes word by word (consecutive threads read consecutive vectors: a warp covers 1536
// contiguous bytes, every line is used whole out of L1), adds the 48 entries in order, and feeds the same per-warp
// threshold / append / bitonic-compact top-k as the exact sweep.  Algorithmic bytes per scan = n x (m + 4).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "sg_internal.h"

using namespace sg;

struct sg_pq {
    int device = 0, sm_count = 148;
    uint32_t dim = 0, ds = 0, m = 0, k = 0;
    float *d_centroids = nullptr;              // [m][k][ds]
    uint8_t *d_codes = nullptr;                // [cap][m]
    uint32_t *d_ids = nullptr;                 // [cap]
    uint64_t n = 0, cap = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    float *d_rows = nullptr; size_t rows_cap = 0;        // staging for encode / queries (floats)
    float *d_table = nullptr; size_t table_cap = 0;      // [b][m][k]
    uint64_t *d_lists = nullptr; size_t lists_cap = 0;
    uint64_t *d_lists2 = nullptr; size_t lists2_cap = 0;
    uint32_t *d_rids = nullptr; float *d_rdist = nullptr; uint32_t *d_rcnt = nullptr; size_t res_cap = 0, cnt_cap = 0;
    uint8_t *d_tmp_codes = nullptr; size_t tmp_codes_cap = 0;
};

namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;

// pq.rs:393-401: (x - y).powi(2) summed left to right
SG_DEVICE_INLINE float sq_l2(const float *a, const float *b, uint32_t n)
{
    float s = 0.0f;
    for (uint32_t i = 0; i < n; ++i) {
        const float d = __fsub_rn(a[i], b[i]);
        s = __fadd_rn(s, __fmul_rn(d, d));
    }
    return s;
}

// encode (:220-257): grid (row tiles of 256, subspace); the subspace's centroids sit in shared memory
__global__ void __launch_bounds__(256)
pq_encode_kernel(const float *rows, uint64_t n, uint32_t dim, uint32_t ds, uint32_t m, uint32_t k, const float *centroids,
                 uint8_t *codes)
{
    extern __shared__ float cent[];            // [k][ds]
    const uint32_t s = blockIdx.y;
    for (uint32_t i = threadIdx.x; i < k * ds; i += blockDim.x) cent[i] = centroids[(size_t)s * k * ds + i];
    __syncthreads();
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    float sub[16];
    for (uint32_t j = 0; j < ds; ++j) sub[j] = rows[row * dim + (size_t)s * ds + j];
    uint32_t best = 0;
    float best_d = 3.402823466e+38f;
    for (uint32_t c = 0; c < k; ++c) {
        const float d = sq_l2(sub, cent + (size_t)c * ds, ds);
        if (d < best_d) {
            best_d = d;
            best = c;
        }
    }
    codes[row * m + s] = (uint8_t)best;
}

// decode (:260-290): one thread per (row, subspace)
__global__ void pq_decode_kernel(const uint8_t *codes, uint64_t n, uint32_t ds, uint32_t m, uint32_t k, const float *centroids,
                                 float *rows)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * m) return;
    const uint32_t s = (uint32_t)(t % m);
    uint32_t c = codes[t];
    if (c >= k) c = k - 1;                     // the host rejects out-of-range codes before launching
    for (uint32_t j = 0; j < ds; ++j) rows[t * ds + j] = centroids[((size_t)s * k + c) * ds + j];
}

// build_distance_table (:329-353): grid (query, subspace), thread = centroid
__global__ void pq_table_kernel(const float *queries, uint32_t dim, uint32_t ds, uint32_t m, uint32_t k, const float *centroids,
                                float *table)
{
    const uint32_t q = blockIdx.x, s = blockIdx.y, c = threadIdx.x;
    __shared__ float sub[16];
    if (threadIdx.x < ds) sub[threadIdx.x] = queries[(size_t)q * dim + (size_t)s * ds + threadIdx.x];
    __syncthreads();
    if (c < k) table[((size_t)q * m + s) * k + c] = sq_l2(sub, centroids + ((size_t)s * k + c) * ds, ds);
}

// k-means assignment step (:173-186): one thread per (vector, subspace); centroids of the subspace in shared memory
__global__ void __launch_bounds__(256)
pq_assign_kernel(const float *vecs, uint64_t n, uint32_t dim, uint32_t ds, uint32_t k, const float *centroids, uint32_t *assign)
{
    extern __shared__ float cent[];
    const uint32_t s = blockIdx.y;
    for (uint32_t i = threadIdx.x; i < k * ds; i += blockDim.x) cent[i] = centroids[(size_t)s * k * ds + i];
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float sub[16];
    for (uint32_t j = 0; j < ds; ++j) sub[j] = vecs[i * dim + (size_t)s * ds + j];
    uint32_t best = 0;
    float best_d = 3.402823466e+38f;
    for (uint32_t c = 0; c < k; ++c) {
        const float d = sq_l2(sub, cent + (size_t)c * ds, ds);
        if (d < best_d) {
            best_d = d;
            best = c;
        }
    }
    assign[(size_t)s * n + i] = best;
}

// k-means update step (:188-210): the members of a cluster are added in VECTOR ORDER (that is the reference's
// order), so one thread owns one (subspace, centroid, component) and walks the assignment list; empty clusters
// keep their centroid.
__global__ void pq_update_kernel(const float *vecs, uint64_t n, uint32_t dim, uint32_t ds, uint32_t k, const uint32_t *assign,
                                 float *centroids, uint32_t m)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)m * k * ds) return;
    const uint32_t j = (uint32_t)(t % ds), c = (uint32_t)((t / ds) % k), s = (uint32_t)(t / ((uint64_t)ds * k));
    const uint32_t *as = assign + (size_t)s * n;
    float acc = 0.0f;
    uint64_t cnt = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (as[i] == c) {
            acc = __fadd_rn(acc, vecs[i * dim + (size_t)s * ds + j]);
            ++cnt;
        }
    if (cnt) centroids[t] = __fdiv_rn(acc, (float)cnt);
}

// the same update for large training sets: members of (subspace s, cluster c) are members[s*n + moff[s*(k+1)+c] ..
// moff[s*(k+1)+c+1]), listed in vector order (a stable counting sort of the assignments, arranged by the host), so a
// thread adds only its own cluster's members -- n*dim work in total instead of n per thread -- in the same order
__global__ void pq_update_members_kernel(const float *vecs, uint64_t n, uint32_t dim, uint32_t ds, uint32_t k, const uint64_t *moff,
                                         const uint32_t *members, float *centroids, uint32_t m)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)m * k * ds) return;
    const uint32_t j = (uint32_t)(t % ds), c = (uint32_t)((t / ds) % k), s = (uint32_t)(t / ((uint64_t)ds * k));
    const uint64_t lo = moff[(size_t)s * (k + 1) + c], hi = moff[(size_t)s * (k + 1) + c + 1];
    if (hi == lo) return;
    const uint32_t *mem = members + (size_t)s * n;
    float acc = 0.0f;
    for (uint64_t e = lo; e < hi; ++e) acc = __fadd_rn(acc, vecs[(size_t)mem[e] * dim + (size_t)s * ds + j]);
    centroids[t] = __fdiv_rn(acc, (float)(hi - lo));
}

// initial centroids (:160-170): the first k entries of the shuffled order (wrapping when n < k)
__global__ void pq_init_kernel(const float *vecs, uint64_t n, uint32_t dim, uint32_t ds, uint32_t k, const uint32_t *order,
                               float *centroids, uint32_t m)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)m * k * ds) return;
    const uint32_t j = (uint32_t)(t % ds), c = (uint32_t)((t / ds) % k), s = (uint32_t)(t / ((uint64_t)ds * k));
    const uint64_t src = order[(size_t)s * n + (c < n ? c : c % n)];
    centroids[t] = vecs[src * dim + (size_t)s * ds + j];
}

// ---- the ADC scan ----------------------------------------------------------------------------------------
// first shared-memory region: the queries' tables during the sweep, the W x topk merge buffer after it
__host__ __device__ inline size_t pq_region0_bytes(uint32_t qb, uint32_t m, uint32_t k, uint32_t W, uint32_t topk)
{
    uint32_t M = 2;
    while (M < W * topk) M <<= 1;
    const size_t t = (size_t)qb * m * k * 4, mb = (size_t)M * 8;
    return ((t > mb ? t : mb) + 15) & ~(size_t)15;
}

struct PqScanArgs {
    const uint8_t *codes;      // [n][m]
    const uint32_t *ids;       // [n]
    uint64_t n;
    uint32_t m, k;             // codes per vector, centroids per subspace
    const float *table;        // [QB][m][k] of this launch's queries
    uint32_t nq;               // live queries (<= QB)
    uint32_t topk, cap;        // results per query; per-(warp, query) candidate capacity (pow2 >= topk + 32)
    uint64_t *out_lists;       // [nq][gridDim.x][topk] sorted keys, KEY_SENTINEL padded
};

SG_DEVICE_INLINE void pq_compact(uint64_t *buf, uint32_t &cnt, uint32_t &tau, uint32_t k, uint32_t lane)
{
    uint32_t mm = next_pow2(cnt < 2 ? 2 : cnt);
    for (uint32_t i = cnt + lane; i < mm; i += 32) buf[i] = KEY_SENTINEL;
    __syncwarp();
    warp_bitonic_sort(buf, mm, lane);
    if (cnt > k) cnt = k;
    if (cnt == k) tau = (uint32_t)(buf[k - 1] >> 32);
}

template <int QB>
__global__ void __launch_bounds__(256, 1)
pq_scan_kernel(const PqScanArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const uint32_t m = a.m, k = a.k, C = a.cap;
    float *lut = reinterpret_cast<float *>(smem);                                   // [QB][m*k]; later the CTA's merge buffer
    uint64_t *cand = reinterpret_cast<uint64_t *>(smem + pq_region0_bytes(QB, m, k, W, a.topk));   // [W][QB][C]
    uint32_t *cnt_s = reinterpret_cast<uint32_t *>(cand + (size_t)W * QB * C);      // [W][QB]
    for (uint32_t i = threadIdx.x; i < (uint32_t)QB * m * k; i += blockDim.x)
        lut[i] = i < a.nq * m * k ? a.table[i] : 0.0f;
    __syncthreads();

    uint32_t cnt[QB], tau[QB];
#pragma unroll
    for (int q = 0; q < QB; ++q) {
        cnt[q] = 0;
        tau[q] = 0xFFFFFFFFu;
    }
    uint64_t *my_cand = cand + (size_t)warp * QB * C;
    const uint64_t n_tiles = (a.n + 31) / 32;
    const uint64_t G = (uint64_t)gridDim.x * W;
    const uint32_t words = m / 4;
    const uint32_t *codes32 = reinterpret_cast<const uint32_t *>(a.codes);
    for (uint64_t tile = (uint64_t)blockIdx.x * W + warp; tile < n_tiles; tile += G) {
        const uint64_t row = tile * 32 + lane;
        const bool live = row < a.n;
        float total[QB];
        bool bad = false;
#pragma unroll
        for (int q = 0; q < QB; ++q) total[q] = 0.0f;
        uint32_t id = 0;
        if (live) {
            id = a.ids[row];
            const uint32_t *cw = codes32 + row * words;
            for (uint32_t w = 0; w < words; ++w) {
                const uint32_t four = __ldg(cw + w);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint32_t code = (four >> (8 * t)) & 0xFFu;
                    const uint32_t s = w * 4 + t;
                    bad |= code >= k;
                    const uint32_t at = s * k + (code < k ? code : 0);
#pragma unroll
                    for (int q = 0; q < QB; ++q) total[q] = __fadd_rn(total[q], lut[(size_t)q * m * k + at]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < QB; ++q) {
            const float d = bad ? 3.402823466e+38f : total[q];             // distance_with_table: f32::MAX on a corrupt code
            const uint32_t key32 = order_key32(d != d ? __uint_as_float(0x7FC00000u) : d);
            uint64_t *buf = my_cand + (size_t)q * C;
            if (cnt[q] + 32 > C) pq_compact(buf, cnt[q], tau[q], a.topk, lane);
            const bool hit = live && (q < (int)a.nq) && key32 <= tau[q];
            const uint32_t mk = __ballot_sync(FULL, hit);
            if (mk) {
                if (hit) buf[cnt[q] + __popc(mk & ((1u << lane) - 1))] = make_key(key32, id);
                cnt[q] += __popc(mk);
                __syncwarp();
            }
        }
    }
#pragma unroll
    for (int q = 0; q < QB; ++q) {
        pq_compact(my_cand + (size_t)q * C, cnt[q], tau[q], a.topk, lane);
        if (lane == 0) cnt_s[warp * QB + q] = cnt[q];
    }
    __syncthreads();
    // per-CTA merge of the W warp lists: the table is no longer needed, sort in its place
    uint64_t *mbuf = reinterpret_cast<uint64_t *>(lut);
    const uint32_t topk = a.topk;
    const uint32_t M = next_pow2(W * topk < 2 ? 2 : W * topk);
    for (uint32_t q = 0; q < a.nq; ++q) {
        for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
            const uint32_t w = i / topk, r = i - w * topk;
            uint64_t key = KEY_SENTINEL;
            if (w < W && r < cnt_s[w * QB + q]) key = cand[((size_t)w * QB + q) * C + r];
            mbuf[i] = key;
        }
        __syncthreads();
        block_bitonic_sort(mbuf, M);
        uint64_t *out = a.out_lists + ((size_t)q * gridDim.x + blockIdx.x) * topk;
        for (uint32_t i = threadIdx.x; i < topk; i += blockDim.x) out[i] = mbuf[i];
        __syncthreads();
    }
}

// Four queries per sweep with an INTERLEAVED table: lut4[(s * k + code)] is a float4 holding the entry of all
// four queries, so ONE 16-byte shared-memory gather serves four look-ups (the scan is bound by shared-memory
// gathers, not by HBM: 48 x QB look-ups per 52 bytes streamed -- see DESIGN.md 3.5d).  Each query still adds its
// 48 terms in subvector order (`total += table[m][code[m]]`, pq.rs:358-369): bit-identical to the CPU path.
// 16-byte gathers also conflict less than 4-byte ones (a quarter-warp of 8 lanes shares 8 bank groups).
__global__ void __launch_bounds__(512, 1)
pq_scan4_kernel(const PqScanArgs a)
{
    constexpr int QB = 4;
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const uint32_t m = a.m, k = a.k, C = a.cap;
    float4 *lut4 = reinterpret_cast<float4 *>(smem);                                 // [m*k]; later the CTA's merge buffer
    uint64_t *cand = reinterpret_cast<uint64_t *>(smem + pq_region0_bytes(QB, m, k, W, a.topk));   // [W][QB][C]
    uint32_t *cnt_s = reinterpret_cast<uint32_t *>(cand + (size_t)W * QB * C);      // [W][QB]
    for (uint32_t i = threadIdx.x; i < m * k; i += blockDim.x) {
        float4 v;
        v.x = a.table[i];
        v.y = a.nq > 1 ? a.table[(size_t)m * k + i] : 0.0f;
        v.z = a.nq > 2 ? a.table[(size_t)2 * m * k + i] : 0.0f;
        v.w = a.nq > 3 ? a.table[(size_t)3 * m * k + i] : 0.0f;
        lut4[i] = v;
    }
    __syncthreads();

    uint32_t cnt[QB], tau[QB];
#pragma unroll
    for (int q = 0; q < QB; ++q) {
        cnt[q] = 0;
        tau[q] = 0xFFFFFFFFu;
    }
    uint64_t *my_cand = cand + (size_t)warp * QB * C;
    const uint64_t n_tiles = (a.n + 31) / 32;
    const uint64_t G = (uint64_t)gridDim.x * W;
    const uint32_t words = m / 4;
    const uint32_t *codes32 = reinterpret_cast<const uint32_t *>(a.codes);
    for (uint64_t tile = (uint64_t)blockIdx.x * W + warp; tile < n_tiles; tile += G) {
        const uint64_t row = tile * 32 + lane;
        const bool live = row < a.n;
        float total[QB] = {0.0f, 0.0f, 0.0f, 0.0f};
        bool bad = false;
        uint32_t id = 0;
        if (live) {
            id = a.ids[row];
            auto add4 = [&](uint32_t four, uint32_t w) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint32_t code = (four >> (8 * t)) & 0xFFu;
                    const uint32_t s = w * 4 + t;
                    bad |= code >= k;
                    const float4 v = lut4[s * k + (code < k ? code : 0)];
                    total[0] = __fadd_rn(total[0], v.x);
                    total[1] = __fadd_rn(total[1], v.y);
                    total[2] = __fadd_rn(total[2], v.z);
                    total[3] = __fadd_rn(total[3], v.w);
                }
            };
            if ((m & 15u) == 0) {                 // 16-byte code loads: 3 per 48-byte code instead of 12 word loads that
                const uint4 *c4 = reinterpret_cast<const uint4 *>(a.codes + row * m);   // each touch every sector of the warp's span
                for (uint32_t w4 = 0; w4 < m / 16; ++w4) {
                    const uint4 v = __ldg(c4 + w4);
                    add4(v.x, 4 * w4);
                    add4(v.y, 4 * w4 + 1);
                    add4(v.z, 4 * w4 + 2);
                    add4(v.w, 4 * w4 + 3);
                }
            } else {
                const uint32_t *cw = codes32 + row * words;
                for (uint32_t w = 0; w < words; ++w) add4(__ldg(cw + w), w);
            }
        }
#pragma unroll
        for (int q = 0; q < QB; ++q) {
            const float d = bad ? 3.402823466e+38f : total[q];             // distance_with_table: f32::MAX on a corrupt code
            const uint32_t key32 = order_key32(d != d ? __uint_as_float(0x7FC00000u) : d);
            uint64_t *buf = my_cand + (size_t)q * C;
            if (cnt[q] + 32 > C) pq_compact(buf, cnt[q], tau[q], a.topk, lane);
            const bool hit = live && (q < (int)a.nq) && key32 <= tau[q];
            const uint32_t mk = __ballot_sync(FULL, hit);
            if (mk) {
                if (hit) buf[cnt[q] + __popc(mk & ((1u << lane) - 1))] = make_key(key32, id);
                cnt[q] += __popc(mk);
                __syncwarp();
            }
        }
    }
#pragma unroll
    for (int q = 0; q < QB; ++q) {
        pq_compact(my_cand + (size_t)q * C, cnt[q], tau[q], a.topk, lane);
        if (lane == 0) cnt_s[warp * QB + q] = cnt[q];
    }
    __syncthreads();
    uint64_t *mbuf = reinterpret_cast<uint64_t *>(lut4);       // the table is no longer needed: merge in its place
    const uint32_t topk = a.topk;
    const uint32_t M = next_pow2(W * topk < 2 ? 2 : W * topk);
    for (uint32_t q = 0; q < a.nq; ++q) {
        for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
            const uint32_t w = i / topk, r = i - w * topk;
            uint64_t key = KEY_SENTINEL;
            if (w < W && r < cnt_s[w * QB + q]) key = cand[((size_t)w * QB + q) * C + r];
            mbuf[i] = key;
        }
        __syncthreads();
        block_bitonic_sort(mbuf, M);
        uint64_t *out = a.out_lists + ((size_t)q * gridDim.x + blockIdx.x) * topk;
        for (uint32_t i = threadIdx.x; i < topk; i += blockDim.x) out[i] = mbuf[i];
        __syncthreads();
    }
}

size_t pq_scan_smem(uint32_t qb, uint32_t m, uint32_t k, uint32_t W, uint32_t cap, uint32_t topk)
{
    return pq_region0_bytes(qb, m, k, W, topk) + (size_t)W * qb * cap * 8 + (size_t)W * qb * 4 + 16;
}

template <typename T>
int pq_grow(T *&ptr, size_t &cap, size_t need, bool keep, size_t keep_elems, cudaStream_t st)
{
    if (need <= cap) return SG_OK;
    size_t ncap = need + need / 2;
    T *np = nullptr;
    SG_CUDA_TRY(cudaMalloc(&np, ncap * sizeof(T)));
    if (keep && ptr && keep_elems) {
        cudaError_t e = cudaMemcpyAsync(np, ptr, keep_elems * sizeof(T), cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            cudaFree(np);
            set_error("device copy during growth failed: %s", cudaGetErrorString(e));
            return SG_ERR_CUDA;
        }
    }
    if (ptr) cudaFree(ptr);
    ptr = np;
    cap = ncap;
    return SG_OK;
}

int pq_check_shape(uint32_t dim, uint32_t ds, uint32_t num_centroids)
{
    if (dim == 0 || ds == 0 || ds > 16 || dim % ds != 0 || (dim / ds) % 4 != 0) {
        set_error("PQ shape dim=%u subvec_dim=%u unsupported: subvec_dim <= 16, dim %% subvec_dim == 0, dim / subvec_dim a multiple of 4",
                  dim, ds);
        return SG_ERR_UNSUPPORTED;
    }
    if (num_centroids == 0 || num_centroids > 256) {
        set_error("num_centroids=%u outside 1..256 (codes are one byte, pq.rs:17)", num_centroids);
        return SG_ERR_UNSUPPORTED;
    }
    return SG_OK;
}

int pq_device(int device, int *sms)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s); libshodh_gpu has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return SG_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        set_error("device ordinal %d out of range (0..%d)", device, n - 1);
        return SG_ERR_INVALID;
    }
    cudaDeviceProp p;
    SG_CUDA_TRY(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        set_error("device %d is sm_%d%d; libshodh_gpu is built for sm_100a (B200) only", device, p.major, p.minor);
        return SG_ERR_NO_DEVICE;
    }
    *sms = p.multiProcessorCount;
    return SG_OK;
}

int pq_alloc(sg_pq **out, int device, uint32_t dim, uint32_t ds, uint32_t k)
{
    int sms = 0;
    int rc = pq_device(device, &sms);
    if (rc != SG_OK) return rc;
    SG_CUDA_TRY(cudaSetDevice(device));
    sg_pq *p = new (std::nothrow) sg_pq();
    if (!p) return SG_ERR_OOM;
    p->device = device;
    p->sm_count = sms;
    p->dim = dim;
    p->ds = ds;
    p->m = dim / ds;
    p->k = k;
    cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_centroids, (size_t)p->m * k * ds * 4);
    if (e != cudaSuccess) {
        set_error("PQ allocation failed: %s", cudaGetErrorString(e));
        if (p->stream) cudaStreamDestroy(p->stream);
        delete p;
        return SG_ERR_CUDA;
    }
    *out = p;
    return SG_OK;
}

// rows (host f32 [n][dim]) -> codes on the device at d_codes_out, in pieces through the staging buffer
int pq_encode_to(sg_pq *p, const float *rows, uint64_t n, uint8_t *d_codes_out)
{
    const uint64_t piece = ((size_t)64 << 20) / ((size_t)p->dim * 4) ? ((size_t)64 << 20) / ((size_t)p->dim * 4) : 1;
    int rc = pq_grow(p->d_rows, p->rows_cap, (size_t)(n < piece ? n : piece) * p->dim, false, 0, p->stream);
    if (rc != SG_OK) return rc;
    for (uint64_t done = 0; done < n; done += piece) {
        const uint64_t cnt = n - done < piece ? n - done : piece;
        SG_CUDA_TRY(cudaMemcpyAsync(p->d_rows, rows + done * p->dim, (size_t)cnt * p->dim * 4, cudaMemcpyHostToDevice, p->stream));
        dim3 grid((unsigned)((cnt + 255) / 256), p->m);
        pq_encode_kernel<<<grid, 256, (size_t)p->k * p->ds * 4, p->stream>>>(p->d_rows, cnt, p->dim, p->ds, p->m, p->k,
                                                                             p->d_centroids, d_codes_out + done * p->m);
        count_launch();
        SG_CUDA_TRY(cudaGetLastError());
        SG_CUDA_TRY(cudaStreamSynchronize(p->stream));
    }
    return SG_OK;
}

}  // namespace

extern "C" {

int sg_pq_create(sg_pq **out, int device, uint32_t dim, uint32_t subvec_dim, uint32_t num_centroids, const float *centroids)
{
    if (!out || !centroids) return SG_ERR_INVALID;
    *out = nullptr;
    int rc = pq_check_shape(dim, subvec_dim, num_centroids);
    if (rc != SG_OK) return rc;
    sg_pq *p = nullptr;
    if ((rc = pq_alloc(&p, device, dim, subvec_dim, num_centroids)) != SG_OK) return rc;
    cudaError_t e = cudaMemcpy(p->d_centroids, centroids, (size_t)p->m * p->k * p->ds * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("centroid upload failed: %s", cudaGetErrorString(e));
        sg_pq_destroy(p);
        return SG_ERR_CUDA;
    }
    *out = p;
    return SG_OK;
}

int sg_pq_train(sg_pq **out, int device, uint32_t dim, uint32_t subvec_dim, uint32_t num_centroids, uint32_t iterations,
                const float *vectors, uint64_t n, const uint32_t *orders)
{
    if (!out || !vectors || !orders) return SG_ERR_INVALID;
    *out = nullptr;
    if (n == 0) {
        set_error("No training vectors provided");          // pq.rs:95-97
        return SG_ERR_INVALID;
    }
    int rc = pq_check_shape(dim, subvec_dim, num_centroids);
    if (rc != SG_OK) return rc;
    const uint32_t k = (uint64_t)num_centroids < n ? num_centroids : (uint32_t)n;     // :118
    sg_pq *p = nullptr;
    if ((rc = pq_alloc(&p, device, dim, subvec_dim, k)) != SG_OK) return rc;
    const uint32_t m = p->m, ds = subvec_dim;
    float *d_vecs = nullptr;
    uint32_t *d_order = nullptr, *d_assign = nullptr;
    cudaError_t e = cudaMalloc(&d_vecs, (size_t)n * dim * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_order, (size_t)m * n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_assign, (size_t)m * n * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_vecs, vectors, (size_t)n * dim * 4, cudaMemcpyHostToDevice, p->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_order, orders, (size_t)m * n * 4, cudaMemcpyHostToDevice, p->stream);
    if (e == cudaSuccess) {
        const uint64_t tot = (uint64_t)m * k * ds;
        pq_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, p->stream>>>(d_vecs, n, dim, ds, k, d_order, p->d_centroids, m);
        count_launch();
        // small training sets: every thread walks the whole assignment list (no host round trip); large ones
        // (n >= SG_PQ_MEMBERS_MIN_N, default 20000): per-cluster member lists in vector order, arranged by the host
        const char *thr = getenv("SG_PQ_MEMBERS_MIN_N");
        const uint64_t members_min_n = thr ? (uint64_t)atoll(thr) : 20000;
        const bool use_members = n >= members_min_n;
        uint32_t *d_members = nullptr;
        uint64_t *d_moff = nullptr;
        std::vector<uint32_t> h_assign, h_members;
        std::vector<uint64_t> h_moff, cur;
        if (use_members) {
            e = cudaMalloc(&d_members, (size_t)m * n * 4);
            if (e == cudaSuccess) e = cudaMalloc(&d_moff, (size_t)m * (k + 1) * 8);
            try {                                  // no exception may cross the C ABI (and the device buffers must be freed)
                h_assign.resize((size_t)m * n);
                h_members.resize((size_t)m * n);
                h_moff.resize((size_t)m * (k + 1));
                cur.resize(k);
            } catch (const std::bad_alloc &) {
                e = cudaErrorMemoryAllocation;
            }
        }
        for (uint32_t it = 0; it < iterations && e == cudaSuccess; ++it) {
            dim3 grid((unsigned)((n + 255) / 256), m);
            pq_assign_kernel<<<grid, 256, (size_t)k * ds * 4, p->stream>>>(d_vecs, n, dim, ds, k, p->d_centroids, d_assign);
            count_launch();
            if (!use_members) {
                pq_update_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, p->stream>>>(d_vecs, n, dim, ds, k, d_assign, p->d_centroids, m);
                count_launch();
                continue;
            }
            e = cudaMemcpyAsync(h_assign.data(), d_assign, (size_t)m * n * 4, cudaMemcpyDeviceToHost, p->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
            if (e != cudaSuccess) break;
            for (uint32_t s2 = 0; s2 < m; ++s2) {                       // stable counting sort of subspace s2's assignments
                const uint32_t *as = h_assign.data() + (size_t)s2 * n;
                uint64_t *off = h_moff.data() + (size_t)s2 * (k + 1);
                std::fill(off, off + k + 1, 0);
                for (uint64_t i = 0; i < n; ++i) ++off[as[i] + 1];
                for (uint32_t c = 0; c < k; ++c) off[c + 1] += off[c];
                std::copy(off, off + k, cur.begin());
                uint32_t *mem = h_members.data() + (size_t)s2 * n;
                for (uint64_t i = 0; i < n; ++i) mem[cur[as[i]]++] = (uint32_t)i;
            }
            e = cudaMemcpyAsync(d_moff, h_moff.data(), h_moff.size() * 8, cudaMemcpyHostToDevice, p->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_members, h_members.data(), (size_t)m * n * 4, cudaMemcpyHostToDevice, p->stream);
            if (e == cudaSuccess) {
                pq_update_members_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, p->stream>>>(d_vecs, n, dim, ds, k, d_moff, d_members,
                                                                                           p->d_centroids, m);
                count_launch();
                e = cudaStreamSynchronize(p->stream);
            }
        }
        cudaFree(d_members);
        cudaFree(d_moff);
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    cudaFree(d_vecs);
    cudaFree(d_order);
    cudaFree(d_assign);
    if (e != cudaSuccess) {
        set_error("PQ training failed: %s", cudaGetErrorString(e));
        sg_pq_destroy(p);
        return e == cudaErrorMemoryAllocation ? SG_ERR_OOM : SG_ERR_CUDA;
    }
    *out = p;
    return SG_OK;
}

void sg_pq_destroy(sg_pq *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    cudaFree(p->d_centroids);
    cudaFree(p->d_codes);
    cudaFree(p->d_ids);
    cudaFree(p->d_rows);
    cudaFree(p->d_table);
    cudaFree(p->d_lists);
    cudaFree(p->d_lists2);
    cudaFree(p->d_rids);
    cudaFree(p->d_rdist);
    cudaFree(p->d_rcnt);
    cudaFree(p->d_tmp_codes);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

int sg_pq_info(sg_pq *p, uint32_t *dim_out, uint32_t *num_subvectors_out, uint32_t *num_centroids_out, uint64_t *len_out)
{
    if (!p) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(p->mu);
    if (dim_out) *dim_out = p->dim;
    if (num_subvectors_out) *num_subvectors_out = p->m;
    if (num_centroids_out) *num_centroids_out = p->k;
    if (len_out) *len_out = p->n;
    return SG_OK;
}

int sg_pq_get_centroids(sg_pq *p, float *centroids_out)
{
    if (!p || !centroids_out) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(p->mu);
    SG_CUDA_TRY(cudaSetDevice(p->device));
    SG_CUDA_TRY(cudaMemcpy(centroids_out, p->d_centroids, (size_t)p->m * p->k * p->ds * 4, cudaMemcpyDeviceToHost));
    return SG_OK;
}

int sg_pq_encode(sg_pq *p, const float *rows, uint64_t n, uint8_t *codes_out)
{
    if (!p || (n && (!rows || !codes_out))) return SG_ERR_INVALID;
    if (n == 0) return SG_OK;
    std::lock_guard<std::mutex> g(p->mu);
    SG_CUDA_TRY(cudaSetDevice(p->device));
    int rc = pq_grow(p->d_tmp_codes, p->tmp_codes_cap, (size_t)n * p->m, false, 0, p->stream);
    if (rc != SG_OK) return rc;
    if ((rc = pq_encode_to(p, rows, n, p->d_tmp_codes)) != SG_OK) return rc;
    SG_CUDA_TRY(cudaMemcpy(codes_out, p->d_tmp_codes, (size_t)n * p->m, cudaMemcpyDeviceToHost));
    return SG_OK;
}

int sg_pq_decode(sg_pq *p, const uint8_t *codes, uint64_t n, float *rows_out)
{
    if (!p || (n && (!codes || !rows_out))) return SG_ERR_INVALID;
    if (n == 0) return SG_OK;
    for (uint64_t i = 0; i < n * p->m; ++i)
        if (codes[i] >= p->k) {
            set_error("PQ code %u out of bounds for subspace %llu with %u centroids (data corruption?)", codes[i],
                      (unsigned long long)(i % p->m), p->k);                        // pq.rs:277-284
            return SG_ERR_INVALID;
        }
    std::lock_guard<std::mutex> g(p->mu);
    SG_CUDA_TRY(cudaSetDevice(p->device));
    int rc = pq_grow(p->d_tmp_codes, p->tmp_codes_cap, (size_t)n * p->m, false, 0, p->stream);
    if (rc != SG_OK) return rc;
    if ((rc = pq_grow(p->d_rows, p->rows_cap, (size_t)n * p->dim, false, 0, p->stream)) != SG_OK) return rc;
    SG_CUDA_TRY(cudaMemcpyAsync(p->d_tmp_codes, codes, (size_t)n * p->m, cudaMemcpyHostToDevice, p->stream));
    const uint64_t tot = n * p->m;
    pq_decode_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, p->stream>>>(p->d_tmp_codes, n, p->ds, p->m, p->k, p->d_centroids,
                                                                           p->d_rows);
    count_launch();
    SG_CUDA_TRY(cudaGetLastError());
    SG_CUDA_TRY(cudaMemcpyAsync(rows_out, p->d_rows, (size_t)n * p->dim * 4, cudaMemcpyDeviceToHost, p->stream));
    SG_CUDA_TRY(cudaStreamSynchronize(p->stream));
    return SG_OK;
}

int sg_pq_add(sg_pq *p, const uint32_t *ids, const float *rows, uint64_t n)
{
    if (!p || (n && (!ids || !rows))) return SG_ERR_INVALID;
    if (n == 0) return SG_OK;
    std::lock_guard<std::mutex> g(p->mu);
    SG_CUDA_TRY(cudaSetDevice(p->device));
    size_t c1 = (size_t)p->cap * p->m, c2 = p->cap;
    int rc = pq_grow(p->d_codes, c1, (size_t)(p->n + n) * p->m, true, (size_t)p->n * p->m, p->stream);
    if (rc != SG_OK) return rc;
    if ((rc = pq_grow(p->d_ids, c2, (size_t)(p->n + n), true, (size_t)p->n, p->stream)) != SG_OK) return rc;
    p->cap = c1 / p->m < c2 ? c1 / p->m : c2;
    if ((rc = pq_encode_to(p, rows, n, p->d_codes + p->n * p->m)) != SG_OK) return rc;
    SG_CUDA_TRY(cudaMemcpy(p->d_ids + p->n, ids, (size_t)n * 4, cudaMemcpyHostToDevice));
    p->n += n;
    return SG_OK;
}

int sg_pq_table(sg_pq *p, const float *query, float *table_out)
{
    if (!p || !query || !table_out) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(p->mu);
    SG_CUDA_TRY(cudaSetDevice(p->device));
    int rc = pq_grow(p->d_rows, p->rows_cap, (size_t)p->dim, false, 0, p->stream);
    if (rc != SG_OK) return rc;
    if ((rc = pq_grow(p->d_table, p->table_cap, (size_t)p->m * p->k, false, 0, p->stream)) != SG_OK) return rc;
    SG_CUDA_TRY(cudaMemcpyAsync(p->d_rows, query, (size_t)p->dim * 4, cudaMemcpyHostToDevice, p->stream));
    pq_table_kernel<<<dim3(1, p->m), 256, 0, p->stream>>>(p->d_rows, p->dim, p->ds, p->m, p->k, p->d_centroids, p->d_table);
    count_launch();
    SG_CUDA_TRY(cudaGetLastError());
    SG_CUDA_TRY(cudaMemcpyAsync(table_out, p->d_table, (size_t)p->m * p->k * 4, cudaMemcpyDeviceToHost, p->stream));
    SG_CUDA_TRY(cudaStreamSynchronize(p->stream));
    return SG_OK;
}

int sg_pq_search(sg_pq *p, const float *queries, uint32_t b, uint32_t topk, uint32_t *ids_out, float *dist_out,
                 uint32_t *counts_out)
{
    if (!p || (b && (!queries || !ids_out || !dist_out || !counts_out))) return SG_ERR_INVALID;
    if (b == 0) return SG_OK;
    if (topk == 0 || topk > 1024) {
        set_error("k=%u outside the supported range 1..1024", topk);
        return topk == 0 ? SG_ERR_INVALID : SG_ERR_UNSUPPORTED;
    }
    std::lock_guard<std::mutex> g(p->mu);
    SG_CUDA_TRY(cudaSetDevice(p->device));
    cudaStream_t st = p->stream;
    const uint32_t m = p->m, k = p->k;
    // plan: warps per CTA and queries per sweep that fit 227 KB next to the 4*m*k-byte table(s)
    uint32_t kp = 32;
    while (kp < topk) kp <<= 1;
    const uint32_t cap = 2 * kp;
    uint32_t W = 8, QB = b >= 2 ? 2 : 1;
    auto fits = [&](uint32_t qb, uint32_t w) { return pq_scan_smem(qb, m, k, w, cap, topk) <= 220 * 1024; };
    // three or more queries: four per sweep with the interleaved table (one 16-byte gather = four look-ups), as many
    // warps as fit beside the 16*m*k-byte table
    const char *q4 = getenv("SG_PQ_QB4");
    if (b >= 3 && !(q4 && atoi(q4) == 0)) {
        for (uint32_t w4 : {16u, 12u, 8u, 4u}) {
            if (fits(4, w4)) {
                QB = 4;
                W = w4;
                break;
            }
        }
    }
    while (!fits(QB, W)) {
        if (W > 1) W >>= 1;
        else if (QB > 1) { QB = 1; W = 8; }
        else {
            set_error("no PQ scan configuration fits shared memory (m=%u k=%u topk=%u)", m, k, topk);
            return SG_ERR_UNSUPPORTED;
        }
    }
    const uint64_t n_tiles = (p->n + 31) / 32;
    uint64_t gr = (n_tiles + W - 1) / W;
    const uint32_t grid = (uint32_t)(gr < (uint64_t)p->sm_count ? (gr ? gr : 1) : p->sm_count);
    int rc = pq_grow(p->d_rows, p->rows_cap, (size_t)b * p->dim, false, 0, st);
    if (rc == SG_OK) rc = pq_grow(p->d_table, p->table_cap, (size_t)b * m * k, false, 0, st);
    if (rc == SG_OK) rc = pq_grow(p->d_lists, p->lists_cap, (size_t)b * grid * topk, false, 0, st);
    const uint32_t groups = merge_groups(grid, topk);
    if (rc == SG_OK && groups > 1) rc = pq_grow(p->d_lists2, p->lists2_cap, (size_t)b * groups * topk, false, 0, st);
    if (rc == SG_OK && (size_t)b * topk > p->res_cap) {
        size_t c1 = p->res_cap, c2 = p->res_cap;
        rc = pq_grow(p->d_rids, c1, (size_t)b * topk, false, 0, st);
        if (rc == SG_OK) rc = pq_grow(p->d_rdist, c2, (size_t)b * topk, false, 0, st);
        p->res_cap = c1 < c2 ? c1 : c2;
    }
    if (rc == SG_OK) rc = pq_grow(p->d_rcnt, p->cnt_cap, (size_t)b, false, 0, st);
    if (rc != SG_OK) return rc;
    SG_CUDA_TRY(cudaMemcpyAsync(p->d_rows, queries, (size_t)b * p->dim * 4, cudaMemcpyHostToDevice, st));
    pq_table_kernel<<<dim3(b, m), 256, 0, st>>>(p->d_rows, p->dim, p->ds, m, k, p->d_centroids, p->d_table);
    count_launch();
    SG_CUDA_TRY(cudaGetLastError());
    if (p->n == 0) {
        SG_CUDA_TRY(launch_merge_keys(nullptr, b, 0, topk, 0, nullptr, p->d_rids, p->d_rdist, p->d_rcnt, st));
    } else {
        const size_t smem = pq_scan_smem(QB, m, k, W, cap, topk);
        static DevAttr da4, da2, da1;
        for (uint32_t done = 0; done < b; done += QB) {
            PqScanArgs a;
            a.codes = p->d_codes;
            a.ids = p->d_ids;
            a.n = p->n;
            a.m = m;
            a.k = k;
            a.table = p->d_table + (size_t)done * m * k;
            a.nq = b - done < QB ? b - done : QB;
            a.topk = topk;
            a.cap = cap;
            a.out_lists = p->d_lists + (size_t)done * grid * topk;
            if (QB == 4) {
                SG_CUDA_TRY(ensure_smem(pq_scan4_kernel, da4, smem));
                pq_scan4_kernel<<<grid, W * 32, smem, st>>>(a);
            } else if (QB == 2) {
                SG_CUDA_TRY(ensure_smem(pq_scan_kernel<2>, da2, smem));
                pq_scan_kernel<2><<<grid, W * 32, smem, st>>>(a);
            } else {
                SG_CUDA_TRY(ensure_smem(pq_scan_kernel<1>, da1, smem));
                pq_scan_kernel<1><<<grid, W * 32, smem, st>>>(a);
            }
            count_launch();
            SG_CUDA_TRY(cudaGetLastError());
        }
        SG_CUDA_TRY(launch_merge_keys(p->d_lists, b, grid, topk, 0, groups > 1 ? p->d_lists2 : nullptr, p->d_rids, p->d_rdist,
                                      p->d_rcnt, st));
    }
    SG_CUDA_TRY(cudaMemcpyAsync(ids_out, p->d_rids, (size_t)b * topk * 4, cudaMemcpyDeviceToHost, st));
    SG_CUDA_TRY(cudaMemcpyAsync(dist_out, p->d_rdist, (size_t)b * topk * 4, cudaMemcpyDeviceToHost, st));
    SG_CUDA_TRY(cudaMemcpyAsync(counts_out, p->d_rcnt, (size_t)b * 4, cudaMemcpyDeviceToHost, st));
    SG_CUDA_TRY(cudaStreamSynchronize(st));
    return SG_OK;
}

}  // extern "C"
