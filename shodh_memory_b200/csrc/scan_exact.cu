// scan_exact.cu -- the exact-order corpus sweep ("path A").
//
// Replaces the row loop of VamanaIndex::brute_force_search
// (reference src/vector_db/vamana.rs:1167-1188) for 1..4 queries per sweep:
//   for every live row: dist = -dot(query,row)   [distance_inline.rs:479-481]
// with the dot product evaluated in EXACTLY the reference's summation order
// (scalar build: distance_inline.rs:157-173; avx2 build: :67-111), so the
// distances -- and therefore the (dist, id) ranking -- are bit-identical to the
// CPU path with no re-scoring step.
//
// B200 mapping.  The sweep is HBM-bound (768 B per fp16 row, ~40 warp-instructions
// of fp32 work per row against a budget of ~130 issue slots per row per SM at
// 6.5 TB/s), so the design is about streaming:
//   * the corpus is read with 2-D TMA tiles (cp.async.bulk.tensor, SWIZZLE_128B):
//     one box = 32 rows x 128 B; a row tile (32 rows) is nbox boxes;
//   * every warp is its own producer AND consumer: it owns a private ring of
//     `ring` box slots in shared memory, waits on the slot's mbarrier, consumes
//     the box (one row per lane -- the 128B swizzle makes the 16-byte
//     lane-strided LDS conflict-free), then lane 0 re-arms the slot with the
//     next box.  No block-wide barrier exists in the main loop, ~24 KB per warp
//     is always in flight;
//   * one thread owns one row, so the reference's strictly sequential
//     `sum += ((p0+p1)+p2)+p3` chain needs no cross-lane reduction at all;
//   * top-k: per-warp threshold + append buffer in shared memory, compacted by
//     an in-smem bitonic sort when full (O(log) times per sweep); per-CTA merge
//     at the end; one sorted list of k keys per (query, CTA) goes to HBM.
// Algorithmic bytes per sweep = n_rows * dim * sizeof(elem)  (SURVEY.md 8d).
#include "sg_internal.h"
#include "sg_dist.cuh"

namespace sg {

namespace {

constexpr uint32_t BOX_BYTES = 4096;  // 32 rows x 128 B
constexpr uint32_t ROWS_PER_TILE = 32;

template <int QB>
struct WarpTopK {
    uint32_t cnt[QB];
    uint32_t tau[QB];
};


// sort the warp's candidate buffer, keep the best k, refresh the threshold
SG_DEVICE_INLINE void compact(uint64_t *buf, uint32_t &cnt, uint32_t &tau, uint32_t k, uint32_t lane)
{
    uint32_t m = next_pow2(cnt < 2 ? 2 : cnt);
    for (uint32_t i = cnt + lane; i < m; i += 32) buf[i] = KEY_SENTINEL;
    __syncwarp();
    warp_bitonic_sort(buf, m, lane);
    if (cnt > k) cnt = k;
    if (cnt == k) tau = (uint32_t)(buf[k - 1] >> 32);
}

template <int QB, bool F16, int ORDER, int METRIC>
__global__ void __launch_bounds__(512, 1)
scan_exact_kernel(const __grid_constant__ CUtensorMap tmap, const ScanArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t W = blockDim.x >> 5;
    const uint32_t ring = a.ring;
    const uint32_t nbox = a.nbox;
    const uint32_t dim = a.dim;
    const uint32_t C = a.cap;

    // ---- shared memory carve-up (host computes the same layout: scan_smem_bytes) ----
    uint8_t *ring_base = smem;                                     // [W][ring][4096], 1024-aligned
    float *q_s = reinterpret_cast<float *>(smem + (size_t)W * ring * BOX_BYTES);  // [QB][dim]
    uint64_t *cand = reinterpret_cast<uint64_t *>(q_s + (size_t)QB * dim);        // [W][QB][C]
    uint32_t *cnt_s = reinterpret_cast<uint32_t *>(cand + (size_t)W * QB * C);    // [W][QB]
    uint64_t *bars = reinterpret_cast<uint64_t *>(cnt_s + ((W * QB + 1) & ~1u));  // [W][ring]
    float *qn_s = reinterpret_cast<float *>(bars + (size_t)W * ring);             // [QB] |query|^2 (cosine metrics)

    // "flagged" mode (device-side conditional re-run of uncertified tensor-path queries): the number
    // of queries and their indices live in device memory; with zero flagged queries every CTA exits here.
    uint32_t n_total = a.nq;
    if (a.flag_count) {
        pdl_wait();                              // flagged launches are PDL launches: wait for the certification results
        n_total = *a.flag_count;
        if (n_total == 0) return;
    }
    const uint32_t n_groups = a.flag_count ? (n_total + QB - 1) / QB : 1;
    const uint32_t my_bars = smem_u32(bars + (size_t)warp * ring);
    const uint32_t my_ring = smem_u32(ring_base + (size_t)warp * ring * BOX_BYTES);
    if (lane == 0) {
        for (uint32_t s = 0; s < ring; ++s) mbar_init(my_bars + 8 * s, 1);
        mbar_fence_init();
        if (warp == 0) tma_prefetch_desc(&tmap);
    }
    __syncthreads();

    // ---- this warp's tile sequence: tiles gw, gw+G, gw+2G, ... ----
    const uint64_t n_tiles = (a.n_rows + ROWS_PER_TILE - 1) / ROWS_PER_TILE;
    const uint64_t G = (uint64_t)gridDim.x * W;
    const uint64_t gw = (uint64_t)blockIdx.x * W + warp;
    const uint64_t my_tiles = gw < n_tiles ? (n_tiles - gw + G - 1) / G : 0;
    const uint64_t my_boxes = my_tiles * nbox;
    const int32_t box_elems = F16 ? 64 : 32;

    uint32_t slot = 0, phase = 0, issue_slot = 0;          // ring position: carried across groups
    for (uint32_t grp = 0; grp < n_groups; ++grp) {
    const uint32_t nq_grp = a.flag_count ? (n_total - grp * QB < (uint32_t)QB ? n_total - grp * QB : (uint32_t)QB) : a.nq;
    for (uint32_t i = threadIdx.x; i < (uint32_t)QB * dim; i += blockDim.x) {
        const uint32_t q = i / dim, d = i - q * dim;
        float v = 0.0f;
        if (q < nq_grp) {
            const uint32_t qi = a.flag_list ? a.flag_list[grp * QB + q] : q;
            v = a.queries[(size_t)qi * dim + d];
        }
        q_s[i] = v;
    }
    __syncthreads();
    if constexpr (METRIC == SG_METRIC_COSINE || METRIC == SG_METRIC_COSINE_SIM) {
        // |query|^2 once per sweep, in the reference's order (l2_norm_squared_*_inline :366-431 == dot(a,a))
        if (threadIdx.x < (uint32_t)QB) qn_s[threadIdx.x] = norm2_smem<ORDER>(q_s + (size_t)threadIdx.x * dim, dim);
        __syncthreads();
    }
    // producer state (lane 0 only uses it, but keep it warp-uniform)
    uint64_t issued = 0;
    uint64_t issue_tile = gw;
    uint32_t issue_j = 0;
    auto issue_one = [&]() {
        if (lane == 0) {
            uint32_t bar = my_bars + 8 * issue_slot;
            mbar_arrive_expect_tx(bar, BOX_BYTES);
            tma_load_2d(my_ring + issue_slot * BOX_BYTES, &tmap, bar, (int32_t)issue_j * box_elems,
                        (int32_t)(issue_tile * ROWS_PER_TILE));
        }
        ++issued;
        if (++issue_slot == ring) issue_slot = 0;
        if (++issue_j == nbox) {
            issue_j = 0;
            issue_tile += G;
        }
    };
    {
        uint64_t pro = my_boxes < ring ? my_boxes : ring;
        for (uint64_t s = 0; s < pro; ++s) issue_one();
    }

    WarpTopK<QB> tk;
#pragma unroll
    for (int q = 0; q < QB; ++q) {
        tk.cnt[q] = 0;
        tk.tau[q] = 0xFFFFFFFFu;
    }
    uint64_t *my_cand = cand + (size_t)warp * QB * C;

    const uint32_t sw = (lane & 7) << 4;  // 128B-swizzle XOR term for this lane's row

    for (uint64_t it = 0; it < my_tiles; ++it) {
        const uint64_t tile = gw + it * G;
        uint32_t del_word = 0;
        if (a.deleted) del_word = __ldg(a.deleted + tile);

        constexpr int NACC = (ORDER == SG_ORDER_AVX2) ? 8 : 1;
        constexpr bool ROWNORM = (METRIC == SG_METRIC_COSINE || METRIC == SG_METRIC_COSINE_SIM);
        float acc[QB][NACC];
        float racc[NACC];      // |row|^2 chain (cosine metrics only)
#pragma unroll
        for (int t = 0; t < NACC; ++t) racc[t] = 0.0f;
#pragma unroll
        for (int q = 0; q < QB; ++q)
#pragma unroll
            for (int t = 0; t < NACC; ++t) acc[q][t] = 0.0f;

        for (uint32_t j = 0; j < nbox; ++j) {
            mbar_wait(my_bars + 8 * slot, phase);
            const uint32_t row_addr = my_ring + slot * BOX_BYTES + lane * 128;
            const float *qj = q_s + j * box_elems;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                             : "r"(row_addr + (((uint32_t)c << 4) ^ sw)));
                if constexpr (F16) {
                    float x[8];
                    {
                        float2 f;
                        f = __half22float2(*reinterpret_cast<__half2 *>(&v.x)); x[0] = f.x; x[1] = f.y;
                        f = __half22float2(*reinterpret_cast<__half2 *>(&v.y)); x[2] = f.x; x[3] = f.y;
                        f = __half22float2(*reinterpret_cast<__half2 *>(&v.z)); x[4] = f.x; x[5] = f.y;
                        f = __half22float2(*reinterpret_cast<__half2 *>(&v.w)); x[6] = f.x; x[7] = f.y;
                    }
                    if constexpr (ROWNORM) {
                        accum4<ORDER, SG_METRIC_NEG_DOT>(racc, make_float4(x[0], x[1], x[2], x[3]), x[0], x[1], x[2], x[3], 0);
                        accum4<ORDER, SG_METRIC_NEG_DOT>(racc, make_float4(x[4], x[5], x[6], x[7]), x[4], x[5], x[6], x[7], 4);
                    }
#pragma unroll
                    for (int q = 0; q < QB; ++q) {
                        const float4 qa = *reinterpret_cast<const float4 *>(qj + q * dim + c * 8);
                        const float4 qb = *reinterpret_cast<const float4 *>(qj + q * dim + c * 8 + 4);
                        accum4<ORDER, METRIC>(acc[q], qa, x[0], x[1], x[2], x[3], 0);
                        accum4<ORDER, METRIC>(acc[q], qb, x[4], x[5], x[6], x[7], 4);
                    }
                } else {
                    const float x0 = __uint_as_float(v.x), x1 = __uint_as_float(v.y);
                    const float x2 = __uint_as_float(v.z), x3 = __uint_as_float(v.w);
                    const int o = (c & 1) * 4;
                    if constexpr (ROWNORM) accum4<ORDER, SG_METRIC_NEG_DOT>(racc, make_float4(x0, x1, x2, x3), x0, x1, x2, x3, o);
#pragma unroll
                    for (int q = 0; q < QB; ++q) {
                        const float4 qa = *reinterpret_cast<const float4 *>(qj + q * dim + c * 4);
                        accum4<ORDER, METRIC>(acc[q], qa, x0, x1, x2, x3, o);
                    }
                }
            }
            // slot consumed by all lanes -> lane 0 refills it with the box `ring` ahead
            __syncwarp();
            if (issued < my_boxes) issue_one();
            if (++slot == ring) {
                slot = 0;
                phase ^= 1;
            }
        }

        // ---- one exact distance per lane; filter against the warp threshold ----
        const uint64_t row = tile * ROWS_PER_TILE + lane;
        const bool live = row < a.n_rows && !((del_word >> lane) & 1u);
#pragma unroll
        for (int q = 0; q < QB; ++q) {
            float qn2 = 0.0f, rn2 = 0.0f;
            if constexpr (ROWNORM) {
                qn2 = qn_s[q];
                rn2 = hsum<ORDER>(racc);
            }
            const uint32_t key32 = order_key32(finish_distance<METRIC>(hsum<ORDER>(acc[q]), qn2, rn2));
            uint64_t *buf = my_cand + (size_t)q * C;
            if (tk.cnt[q] + 32 > C) compact(buf, tk.cnt[q], tk.tau[q], a.k, lane);
            const bool hit = live && (q < (int)nq_grp) && key32 <= tk.tau[q];
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
            if (m) {
                if (hit) buf[tk.cnt[q] + __popc(m & ((1u << lane) - 1))] = make_key(key32, (uint32_t)row);
                tk.cnt[q] += __popc(m);
                __syncwarp();
            }
        }
    }

    // ---- per-warp final lists, then a per-CTA merge to one sorted list per query ----
#pragma unroll
    for (int q = 0; q < QB; ++q) {
        compact(my_cand + (size_t)q * C, tk.cnt[q], tk.tau[q], a.k, lane);
        if (lane == 0) cnt_s[warp * QB + q] = tk.cnt[q];
    }
    __syncthreads();  // every TMA box issued by this CTA has been consumed: the ring is free

    uint64_t *mbuf = reinterpret_cast<uint64_t *>(ring_base);
    const uint32_t k = a.k;
    const uint32_t M = next_pow2(W * k < 2 ? 2 : W * k);
    for (uint32_t q = 0; q < nq_grp; ++q) {
        for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
            uint32_t w = i / k, r = i - w * k;
            uint64_t key = KEY_SENTINEL;
            if (w < W && r < cnt_s[w * QB + q]) key = cand[((size_t)w * QB + q) * C + r];
            mbuf[i] = key;
        }
        __syncthreads();
        block_bitonic_sort(mbuf, M);
        uint64_t *out = a.out_lists + ((size_t)(grp * QB + q) * gridDim.x + blockIdx.x) * k;
        for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) out[i] = mbuf[i];
        __syncthreads();
    }
    }  // group loop
}

template <int QB, bool F16, int ORDER, int METRIC>
cudaError_t launch_t(const CUtensorMap &tmap, const ScanArgs &a, dim3 grid, dim3 block, size_t smem,
                     cudaStream_t st)
{
    auto kern = scan_exact_kernel<QB, F16, ORDER, METRIC>;
    static DevAttr da;                        // per instantiation: raise the opt-in limit only when it grows
    {
        cudaError_t e = ensure_smem(kern, da, smem);
        if (e != cudaSuccess) return e;
    }
    cudaError_t le;
    if (a.flag_count) {
        le = launch_pdl(kern, grid, block, smem, st, tmap, a);
    } else {
        kern<<<grid, block, smem, st>>>(tmap, a);
        le = cudaGetLastError();
    }
    count_launch();
    return le;
}

}  // namespace

size_t scan_smem_bytes(uint32_t W, uint32_t ring, uint32_t qb, uint32_t dim, uint32_t cap)
{
    size_t s = (size_t)W * ring * BOX_BYTES;
    s += (size_t)qb * dim * 4;
    s += (size_t)W * qb * cap * 8;
    s += (size_t)((W * qb + 1) & ~1u) * 4;
    s += (size_t)W * ring * 8;
    s += 16;  // |query|^2 per query of the sweep
    return s;
}

template <int QB, int METRIC>
static cudaError_t launch_qm(const CUtensorMap &tmap, const ScanArgs &a, bool f16, int order, dim3 g, dim3 b, size_t smem,
                             cudaStream_t st)
{
    if (f16) return order == SG_ORDER_AVX2 ? launch_t<QB, true, SG_ORDER_AVX2, METRIC>(tmap, a, g, b, smem, st)
                                           : launch_t<QB, true, SG_ORDER_SCALAR, METRIC>(tmap, a, g, b, smem, st);
    return order == SG_ORDER_AVX2 ? launch_t<QB, false, SG_ORDER_AVX2, METRIC>(tmap, a, g, b, smem, st)
                                  : launch_t<QB, false, SG_ORDER_SCALAR, METRIC>(tmap, a, g, b, smem, st);
}

// NormalizedDotProduct sweeps serve 1, 2 or 4 queries; the other metrics (not on RetrievalEngine's path,
// retrieval.rs:188-193) are instantiated for 1 and 4 only.
cudaError_t launch_scan_exact(const CUtensorMap &tmap, const ScanArgs &a, int qb, bool f16, int order, int metric,
                              uint32_t grid, uint32_t W, size_t smem, cudaStream_t st)
{
    dim3 g(grid), b(W * 32);
    if (metric == SG_METRIC_NEG_DOT) {
        if (qb == 1) return launch_qm<1, SG_METRIC_NEG_DOT>(tmap, a, f16, order, g, b, smem, st);
        if (qb == 2) return launch_qm<2, SG_METRIC_NEG_DOT>(tmap, a, f16, order, g, b, smem, st);
        if (qb == 4) return launch_qm<4, SG_METRIC_NEG_DOT>(tmap, a, f16, order, g, b, smem, st);
        return cudaErrorInvalidValue;
    }
#define SG_DISPATCH_METRIC(M)                                                                  \
    if (metric == M) {                                                                         \
        if (qb == 1) return launch_qm<1, M>(tmap, a, f16, order, g, b, smem, st);              \
        if (qb == 4) return launch_qm<4, M>(tmap, a, f16, order, g, b, smem, st);              \
        return cudaErrorInvalidValue;                                                          \
    }
    SG_DISPATCH_METRIC(SG_METRIC_EUCLIDEAN)
    SG_DISPATCH_METRIC(SG_METRIC_COSINE)
    SG_DISPATCH_METRIC(SG_METRIC_COSINE_SIM)
#undef SG_DISPATCH_METRIC
    return cudaErrorInvalidValue;
}

}  // namespace sg
