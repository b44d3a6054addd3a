// scan_tc.cu -- the tensor-core corpus sweep ("path B") for batched queries.
//
// For b >= 5 queries the row loop of VamanaIndex::brute_force_search (reference
// src/vector_db/vamana.rs:1167-1188) is a [b x dim] . [dim x n_rows] contraction: at b = 256
// it needs ~2 PFLOP/s to keep up with HBM, which only tcgen05 can deliver.  The tensor cores
// see fp16(query) x fp16(row) with f32 accumulation, i.e. APPROXIMATE scores; exactness is
// restored by over-selection + exact re-scoring:
//   K1 (this file, scan_tc_kernel): one sweep; per query keep the KK = k + slack best
//       approximate scores (a GPU-wide shared threshold keeps the candidate traffic tiny);
//   K2 (finalize_tc_kernel): per query take the global top-KK approximate candidates,
//       re-score them in EXACTLY the reference's summation order (distance_inline.rs:157-173 /
//       :67-111), sort by (dist, id), and certify with a rigorous error bound E that no row
//       outside the candidate set can reach the top k:  kth_exact_sim > weakest_kept_approx + E.
//       Queries that cannot be certified are flagged and re-run on the exact path A.
// So the result is bit-identical to path A / the CPU oracle, or the query is flagged.
//
// B200 mapping (one CTA per SM, 256 threads, persistent over 128-row corpus tiles):
//   warp 0   TMA producer: the fp16 query block (A operand, 128 x dim, resident for the whole
//            sweep) and a ring of K-blocks of corpus tiles (B operand, 128 rows x 64 elems,
//            SWIZZLE_128B, K-major -- the same boxes path A streams);
//   warp 1   MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128 (queries) x N=128 (rows),
//            K=16 per instruction, accumulators in TMEM (4 stages x 128 columns);
//   warp 2   TMEM allocator;
//   warps 4-7 epilogue: thread t owns query t = TMEM lane t; tcgen05.ld 32 columns at a time,
//            compares against its threshold, inserts the (rare) hits into its sorted list in
//            shared memory.
// HBM-bound: algorithmic bytes per sweep = n_rows * dim * 2; the MMA work per row is
// 12 cycles against ~33 cycles of HBM time per row per SM.
#include "sg_internal.h"
#include "sg_tc.cuh"

namespace sg {
namespace {

constexpr uint32_t TC_M = 128;          // queries per sweep (TMEM lanes)
constexpr uint32_t TC_N = 128;          // corpus rows per tile (TMEM columns per accumulator)
constexpr uint32_t TC_KB = 64;          // fp16 elements per K-block (one 128B swizzle row)
constexpr uint32_t TC_STAGE_BYTES = TC_N * 128;   // 16 KB
constexpr uint32_t TC_ACC_STAGES = 4;   // 4 x 128 = 512 TMEM columns
constexpr uint32_t TC_THREADS = 256;

SG_DEVICE_INLINE uint32_t ld_relaxed_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}


// ---- the epilogue shared by the 1-CTA and the CTA-pair kernels -------------------------------------
// thread t (0..127 of the 4 epilogue warps) <-> query q_base + t <-> TMEM lane t.  Each query's
// candidate list (32 sorted keys) lives in shared memory as one row of 32 entries.  The common case
// per 32-column chunk is "no score reaches any threshold": one max-tree per lane and one vote.  A hit
// is re-read with tcgen05.ld.x1 and inserted by the WHOLE warp (lane i holds entry i; one ballot
// finds the position, one shuffle shifts the tail) -- no dependent shared-memory chain.
// NCOLS = corpus rows (TMEM columns) per accumulator stage.  tempty_cluster_addr is the
// shared::cluster address of the accumulator-empty barriers (the leader CTA's in pair mode).
template <uint32_t NCOLS>
SG_DEVICE_INLINE void tc_epilogue(const TcArgs &a, uint64_t *lists, uint32_t tmem_base, uint64_t *tfull_bar,
                                  uint32_t tempty_cluster_addr, uint64_t my_tiles, uint64_t tile0, uint64_t tile_stride,
                                  uint32_t q_base, uint32_t nq_local, uint32_t list_idx, uint32_t n_lists,
                                  uint32_t acc_stages)
{
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t = threadIdx.x - 128;
    const uint32_t wq = (warp & 3) * 32;             // first query of this warp
    uint64_t *wlists = lists + (size_t)wq * 32;      // [32 queries][32 entries]
    for (uint32_t i = lane; i < 32 * 32; i += 32) wlists[i] = KEY_SENTINEL;
    __syncwarp();
    const bool active = t < nq_local;
    const uint32_t KKE = a.kk;                       // candidates kept per query and list (<= 32)
    const float NEG_INF = -__uint_as_float(0x7F800000u);
    float tau_f = active ? NEG_INF : -NEG_INF;       // similarity threshold: -inf / never
    float mx_all = NEG_INF;                          // pre-pass: best live score seen by this lane
    uint32_t tau_pub = 0xFFFFFFFFu;                  // last threshold key this lane shared
    uint32_t *gtau = a.gtau + (size_t)q_base + t;
    const uint32_t lane_base = (wq) << 16;

    for (uint64_t it = 0; it < my_tiles; ++it) {
        const uint64_t tile = tile0 + it * tile_stride;
        const uint32_t acc = (uint32_t)(it % acc_stages);
        const uint32_t acc_ph = (uint32_t)((it / acc_stages) & 1);
        if (active && a.mode == 0) {
            const uint32_t g = ld_relaxed_u32(gtau);
            if (g != 0xFFFFFFFFu) tau_f = fmaxf(tau_f, -key32_to_dist(g));
        }
        mbar_wait(smem_u32(tfull_bar + acc), acc_ph);
        tc_fence_after();
#pragma unroll 1
        for (uint32_t c = 0; c < NCOLS / 32; ++c) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + lane_base + acc * NCOLS + c * 32;
            tmem_ld32(taddr, r);
            tmem_ld_wait();
            const uint64_t row0 = tile * NCOLS + c * 32;     // 32-aligned: one bitmap word
            uint32_t dead = 0;
            if (a.deleted && row0 < a.n_rows) dead = __ldg(a.deleted + (row0 >> 5));
            if (row0 + 32 > a.n_rows) dead |= row0 >= a.n_rows ? 0xFFFFFFFFu : ~((1u << (uint32_t)(a.n_rows - row0)) - 1u);
            if (dead) {                                      // rare: neutralise dead columns
#pragma unroll
                for (uint32_t j = 0; j < 32; ++j)
                    if ((dead >> j) & 1u) r[j] = __float_as_uint(NEG_INF);
            }
            float m0 = fmaxf(fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1])), fmaxf(__uint_as_float(r[2]), __uint_as_float(r[3])));
            float m1 = fmaxf(fmaxf(__uint_as_float(r[4]), __uint_as_float(r[5])), fmaxf(__uint_as_float(r[6]), __uint_as_float(r[7])));
            float m2 = fmaxf(fmaxf(__uint_as_float(r[8]), __uint_as_float(r[9])), fmaxf(__uint_as_float(r[10]), __uint_as_float(r[11])));
            float m3 = fmaxf(fmaxf(__uint_as_float(r[12]), __uint_as_float(r[13])), fmaxf(__uint_as_float(r[14]), __uint_as_float(r[15])));
            float m4 = fmaxf(fmaxf(__uint_as_float(r[16]), __uint_as_float(r[17])), fmaxf(__uint_as_float(r[18]), __uint_as_float(r[19])));
            float m5 = fmaxf(fmaxf(__uint_as_float(r[20]), __uint_as_float(r[21])), fmaxf(__uint_as_float(r[22]), __uint_as_float(r[23])));
            float m6 = fmaxf(fmaxf(__uint_as_float(r[24]), __uint_as_float(r[25])), fmaxf(__uint_as_float(r[26]), __uint_as_float(r[27])));
            float m7 = fmaxf(fmaxf(__uint_as_float(r[28]), __uint_as_float(r[29])), fmaxf(__uint_as_float(r[30]), __uint_as_float(r[31])));
            const float mx = fmaxf(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)), fmaxf(fmaxf(m4, m5), fmaxf(m6, m7)));
            if (a.mode == 1) {                               // pre-pass: only the running maximum
                mx_all = fmaxf(mx_all, mx);
                continue;
            }
            if (__any_sync(0xFFFFFFFFu, mx >= tau_f)) {
                // slow path, cost proportional to the number of hits
                uint32_t mask = 0;
#pragma unroll
                for (uint32_t j = 0; j < 32; ++j) mask |= (__uint_as_float(r[j]) >= tau_f) ? (1u << j) : 0u;
                mask &= ~dead;                               // (-inf >= -inf) would let dead columns through
                uint32_t lanes = __ballot_sync(0xFFFFFFFFu, mask != 0);
                while (lanes) {
                    const uint32_t L = __ffs(lanes) - 1;
                    lanes &= lanes - 1;
                    uint32_t mL = __shfl_sync(0xFFFFFFFFu, mask, L);
                    uint64_t *ql = wlists + (size_t)L * 32;
                    while (mL) {
                        const uint32_t j = __ffs(mL) - 1;
                        mL &= mL - 1;
                        uint32_t v;
                        tmem_ld1(taddr + j, v);              // column j for every lane; lane L's is the hit
                        tmem_ld_wait();
                        const float sc = __uint_as_float(__shfl_sync(0xFFFFFFFFu, v, L));
                        const float tauL = __shfl_sync(0xFFFFFFFFu, tau_f, L);
                        if (!(sc >= tauL)) continue;         // the threshold moved since the mask was taken
                        const uint64_t key = make_key(order_key32(-sc), (uint32_t)(row0 + j));
                        const uint64_t e = ql[lane];
                        const uint32_t p = __popc(__ballot_sync(0xFFFFFFFFu, e < key));
                        if (p < KKE) {                       // lanes >= KKE always hold KEY_SENTINEL
                            const uint32_t e_lo = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)e, 1);
                            const uint32_t e_hi = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)(e >> 32), 1);
                            uint64_t ne = lane < p ? e : (lane == p ? key : (((uint64_t)e_hi << 32) | e_lo));
                            if (lane >= KKE) ne = KEY_SENTINEL;
                            ql[lane] = ne;
                            const uint32_t last = __shfl_sync(0xFFFFFFFFu, (uint32_t)(ne >> 32), KKE - 1);
                            if (lane == L && last != 0xFFFFFFFFu) tau_f = -key32_to_dist(last);
                        }
                        __syncwarp();
                    }
                }
            }
        }
        tc_fence_before();
        mbar_arrive_cluster(tempty_cluster_addr + 8 * acc);  // all epilogue threads release the accumulator
        if (a.mode == 0 && active && tau_f > NEG_INF) {      // share the threshold GPU-wide when it moved
            const uint32_t k32 = order_key32(-tau_f);
            if (k32 < tau_pub) {
                tau_pub = k32;
                atomicMin(gtau, k32);
            }
        }
    }

    __syncwarp();
    if (a.mode == 1) {
        // pre-pass result: this CTA's (pair's) best live score per query -- a distinct row per list
        if (active) a.maxima[((size_t)q_base + t) * n_lists + list_idx] = mx_all;
    } else {
        // publish the sorted list of every query of the warp (KEY_SENTINEL padded)
        for (uint32_t qi = 0; qi < 32; ++qi) {
            const uint32_t qt = wq + qi;
            if (qt >= nq_local) break;
            a.lists[((size_t)(q_base + qt) * n_lists + list_idx) * 32 + lane] = wlists[(size_t)qi * 32 + lane];
        }
    }
}

// ---- K1: the sweep -----------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_q,
               const TcArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nkb = a.dim / TC_KB;          // K-blocks per row
    const uint32_t S = a.stages;

    // carve-up: A block | B ring | per-query lists | barriers | tmem ptr
    uint8_t *a_base = smem;                                         // nkb x (128 x 128 B)
    uint8_t *b_base = a_base + (size_t)nkb * TC_M * 128;            // S x 16 KB
    uint64_t *lists = reinterpret_cast<uint64_t *>(b_base + (size_t)S * TC_STAGE_BYTES);  // [128 queries][32]
    uint64_t *bars = lists + (size_t)32 * TC_M;
    uint64_t *full_bar = bars;                       // [S]
    uint64_t *empty_bar = bars + S;                  // [S]
    uint64_t *tfull_bar = bars + 2 * S;              // [TC_ACC_STAGES]
    uint64_t *tempty_bar = tfull_bar + TC_ACC_STAGES;  // [TC_ACC_STAGES]
    uint64_t *a_bar = tempty_bar + TC_ACC_STAGES;    // [1]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_bar + 1);

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < S; ++i) {
            mbar_init(smem_u32(full_bar + i), 1);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        for (uint32_t i = 0; i < TC_ACC_STAGES; ++i) {
            mbar_init(smem_u32(tfull_bar + i), 1);
            mbar_init(smem_u32(tempty_bar + i), 128);
        }
        mbar_init(smem_u32(a_bar), 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmap_rows);
        tma_prefetch_desc(&tmap_q);
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint64_t n_tiles = (a.n_rows + TC_N - 1) / TC_N;
    uint64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (a.tile_limit && my_tiles > a.tile_limit) my_tiles = a.tile_limit;   // pre-pass: a prefix of this CTA's tiles

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_arrive_expect_tx(smem_u32(a_bar), nkb * TC_M * 128);
            for (uint32_t kb = 0; kb < nkb; ++kb)
                tma_load_2d(smem_u32(a_base + (size_t)kb * TC_M * 128), &tmap_q, smem_u32(a_bar), (int32_t)(kb * TC_KB),
                            (int32_t)(a.q_tile * TC_M));
            uint32_t st = 0, ph = 0;
            for (uint64_t it = 0; it < my_tiles; ++it) {
                const uint64_t tile = blockIdx.x + it * gridDim.x;
                for (uint32_t kb = 0; kb < nkb; ++kb) {
                    mbar_wait(smem_u32(empty_bar + st), ph ^ 1);
                    mbar_arrive_expect_tx(smem_u32(full_bar + st), TC_STAGE_BYTES);
                    tma_load_2d(smem_u32(b_base + (size_t)st * TC_STAGE_BYTES), &tmap_rows, smem_u32(full_bar + st),
                                (int32_t)(kb * TC_KB), (int32_t)(tile * TC_N));
                    if (++st == S) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(TC_M, TC_N);
            mbar_wait(smem_u32(a_bar), 0);
            tc_fence_after();
            uint32_t st = 0, ph = 0;
            for (uint64_t it = 0; it < my_tiles; ++it) {
                const uint32_t acc = (uint32_t)(it % TC_ACC_STAGES);
                const uint32_t acc_ph = (uint32_t)((it / TC_ACC_STAGES) & 1);
                mbar_wait(smem_u32(tempty_bar + acc), acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * TC_N;
                for (uint32_t kb = 0; kb < nkb; ++kb) {
                    mbar_wait(smem_u32(full_bar + st), ph);
                    tc_fence_after();
                    const uint64_t da = make_sw128_desc(smem_u32(a_base + (size_t)kb * TC_M * 128));
                    const uint64_t db = make_sw128_desc(smem_u32(b_base + (size_t)st * TC_STAGE_BYTES));
#pragma unroll
                    for (uint32_t k = 0; k < TC_KB / 16; ++k)   // advance 32 B (= 16 fp16) inside the swizzle row
                        umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(smem_u32(empty_bar + st));     // frees the B slot when these MMAs retire
                    if (++st == S) { st = 0; ph ^= 1; }
                }
                umma_commit(smem_u32(tfull_bar + acc));        // accumulator complete -> epilogue
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread t <-> query t <-> TMEM lane t =====
        tc_epilogue<TC_N>(a, lists, tmem_base, tfull_bar, smem_u32(tempty_bar), my_tiles, /*tile0=*/blockIdx.x,
                          /*tile_stride=*/gridDim.x, /*q_base=*/a.q_tile * TC_M, /*nq_local=*/a.nq,
                          /*list_idx=*/blockIdx.x, /*n_lists=*/gridDim.x, TC_ACC_STAGES);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}


// ---- K1 (CTA-pair form): 256 queries per sweep ----------------------------------------------------
// Two CTAs of a cluster (one TPC) act as one MMA unit: tcgen05.mma.cta_group::2 with M = 256 (each
// CTA holds its own 128 queries as the A operand and receives their 128 accumulator rows in its own
// TMEM) and N = 256 corpus rows per tile, of which each CTA loads HALF (128 rows) -- the tensor
// core reads both halves, so every corpus byte is fetched once per pair and the shared-memory read
// traffic per FLOP is half that of the 1-CTA form.  One sweep therefore serves 256 queries.
// Barrier choreography (the leader is cluster rank 0):
//   full[s]   leader's only, 2 arrivals: leader's arrive.expect_tx(2 x 16 KB) + the peer's remote
//             arrive; both CTAs' cta_group::2 TMA loads credit their bytes to it;
//   empty[s]  per CTA; the leader's tcgen05.commit multicasts the arrival to both;
//   tfull[a]  per CTA, same multicast commit;  tempty[a] leader's only, 256 epilogue arrivals.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
scan_tc2_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_q,
                const TcArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t N2 = 256;                 // corpus rows per pair tile (TMEM columns per stage)
    constexpr uint32_t ACC2 = 2;                 // 2 x 256 = 512 TMEM columns
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t nkb = a.dim / TC_KB;
    const uint32_t S = a.stages;

    uint8_t *a_base = smem;                                         // this CTA's 128 queries
    uint8_t *b_base = a_base + (size_t)nkb * TC_M * 128;            // S x 16 KB: this CTA's half of each tile
    uint64_t *lists = reinterpret_cast<uint64_t *>(b_base + (size_t)S * TC_STAGE_BYTES);
    uint64_t *bars = lists + (size_t)32 * TC_M;
    uint64_t *full_bar = bars;
    uint64_t *empty_bar = bars + S;
    uint64_t *tfull_bar = bars + 2 * S;
    uint64_t *tempty_bar = tfull_bar + ACC2;
    uint64_t *a_bar = tempty_bar + ACC2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_bar + 1);

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < S; ++i) {
            mbar_init(smem_u32(full_bar + i), 2);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        for (uint32_t i = 0; i < ACC2; ++i) {
            mbar_init(smem_u32(tfull_bar + i), 1);
            mbar_init(smem_u32(tempty_bar + i), 256);
        }
        mbar_init(smem_u32(a_bar), 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmap_rows);
        tma_prefetch_desc(&tmap_q);
    }
    cluster_sync_all();                                  // both CTAs resident, barriers initialised
    if (warp == 2) tmem_alloc2(smem_u32(tmem_slot), 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const uint64_t n_tiles = (a.n_rows + N2 - 1) / N2;
    uint64_t my_tiles = cluster_id < n_tiles ? (n_tiles - cluster_id + n_clusters - 1) / n_clusters : 0;
    if (a.tile_limit && my_tiles > a.tile_limit) my_tiles = a.tile_limit;

    // resident A operand: this CTA's 128 queries.  Everyone waits, then the pair synchronises so the
    // leader's MMA thread may read the peer's A through the pair datapath.
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(smem_u32(a_bar), nkb * TC_M * 128);
        for (uint32_t kb = 0; kb < nkb; ++kb)
            tma_load_2d(smem_u32(a_base + (size_t)kb * TC_M * 128), &tmap_q, smem_u32(a_bar), (int32_t)(kb * TC_KB),
                        (int32_t)(a.q_tile * 256 + rank * TC_M));
    }
    mbar_wait(smem_u32(a_bar), 0);
    cluster_sync_all();

    if (warp == 0) {
        // ===== TMA producer (both CTAs): my half of every tile =====
        if (lane == 0) {
            uint32_t st = 0, ph = 0;
            for (uint64_t it = 0; it < my_tiles; ++it) {
                const uint64_t tile = cluster_id + it * n_clusters;
                for (uint32_t kb = 0; kb < nkb; ++kb) {
                    mbar_wait(smem_u32(empty_bar + st), ph ^ 1);
                    const uint32_t lfull = mapa_u32(smem_u32(full_bar + st), 0);
                    if (leader) mbar_arrive_expect_tx(smem_u32(full_bar + st), 2 * TC_STAGE_BYTES);
                    else mbar_arrive_cluster(lfull);
                    tma_load_2d_2cta(smem_u32(b_base + (size_t)st * TC_STAGE_BYTES), &tmap_rows, lfull,
                                     (int32_t)(kb * TC_KB), (int32_t)(tile * N2 + rank * TC_M));
                    if (++st == S) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the leader's single thread drives the pair =====
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(256, N2);
            uint32_t st = 0, ph = 0;
            for (uint64_t it = 0; it < my_tiles; ++it) {
                const uint32_t acc = (uint32_t)(it % ACC2);
                const uint32_t acc_ph = (uint32_t)((it / ACC2) & 1);
                mbar_wait(smem_u32(tempty_bar + acc), acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * N2;
                for (uint32_t kb = 0; kb < nkb; ++kb) {
                    mbar_wait(smem_u32(full_bar + st), ph);
                    tc_fence_after();
                    const uint64_t da = make_sw128_desc(smem_u32(a_base + (size_t)kb * TC_M * 128));
                    const uint64_t db = make_sw128_desc(smem_u32(b_base + (size_t)st * TC_STAGE_BYTES));
#pragma unroll
                    for (uint32_t k = 0; k < TC_KB / 16; ++k)
                        umma_f16_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit_2cta(smem_u32(empty_bar + st), 3);     // frees the slot in BOTH CTAs
                    if (++st == S) { st = 0; ph ^= 1; }
                }
                umma_commit_2cta(smem_u32(tfull_bar + acc), 3);        // accumulators ready in BOTH CTAs
            }
        }
    } else if (warp >= 4) {
        tc_epilogue<N2>(a, lists, tmem_base, tfull_bar, mapa_u32(smem_u32(tempty_bar), 0), my_tiles, /*tile0=*/cluster_id,
                        /*tile_stride=*/n_clusters, /*q_base=*/a.q_tile * 256 + rank * TC_M,
                        /*nq_local=*/a.nq > rank * TC_M ? (a.nq - rank * TC_M < TC_M ? a.nq - rank * TC_M : TC_M) : 0,
                        /*list_idx=*/cluster_id, /*n_lists=*/n_clusters, ACC2);
    }

    tc_fence_before();
    cluster_sync_all();          // nobody may still signal or read the peer's barriers / TMEM
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, 512);
    }
}

// ---- query preparation: fp16 rounding + the rigorous error bound -------------------------------
// qh[q] = fp16(q) (zero rows for q >= b up to the 128-padded count); err[q] = |q - fp16(q)|_2,
// qnorm[q] = |q|_2 (both rounded up a little by computing in f64).
__global__ void prep_queries_kernel(const float *__restrict__ q, uint32_t b, uint32_t b_pad, uint32_t dim,
                                    __half *__restrict__ qh, float *__restrict__ err, float *__restrict__ qnorm,
                                    uint32_t *__restrict__ gtau, uint32_t *__restrict__ flags)
{
    const uint32_t row = blockIdx.x;
    __shared__ double s_e[32], s_n[32];
    double e = 0.0, n2 = 0.0;
    for (uint32_t j = threadIdx.x; j < dim; j += blockDim.x) {
        float v = row < b ? q[(size_t)row * dim + j] : 0.0f;
        __half h = __float2half_rn(v);
        qh[(size_t)row * dim + j] = h;
        double d = (double)v - (double)__half2float(h);
        e += d * d;
        n2 += (double)v * (double)v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        e += __shfl_down_sync(0xFFFFFFFFu, e, o);
        n2 += __shfl_down_sync(0xFFFFFFFFu, n2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_e[threadIdx.x >> 5] = e;
        s_n[threadIdx.x >> 5] = n2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double te = 0, tn = 0;
        for (uint32_t w = 0; w < (blockDim.x + 31) / 32; ++w) {
            te += s_e[w];
            tn += s_n[w];
        }
        if (row < b) {
            err[row] = (float)(sqrt(te) * 1.000001);
            qnorm[row] = (float)(sqrt(tn) * 1.000001);
        }
        gtau[row] = 0xFFFFFFFFu;
        if (row < b) flags[row] = 0;
        (void)b_pad;
    }
}

// ---- threshold seeding from the pre-pass -------------------------------------------------------
// maxima[q][cta] = best live approximate score CTA `cta` saw for query q over its first tiles.  The
// kk-th largest of them is a valid lower bound on the global kk-th best (kk distinct rows reach it).
__global__ void __launch_bounds__(256)
tau_init_kernel(const float *__restrict__ maxima, uint32_t n_cta, uint32_t kk, uint32_t *__restrict__ gtau)
{
    __shared__ uint64_t keys[256];
    const uint32_t q = blockIdx.x;
    const uint32_t M = next_pow2(n_cta < 2 ? 2 : n_cta);
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
        uint64_t key = KEY_SENTINEL;
        if (i < n_cta) {
            const float m = maxima[(size_t)q * n_cta + i];
            if (m > -__uint_as_float(0x7F800000u)) key = make_key(order_key32(-m), i);
        }
        keys[i] = key;
    }
    __syncthreads();
    block_bitonic_sort(keys, M);
    if (threadIdx.x == 0) {
        const uint64_t kth = n_cta >= kk ? keys[kk - 1] : KEY_SENTINEL;
        gtau[q] = kth == KEY_SENTINEL ? 0xFFFFFFFFu : (uint32_t)(kth >> 32);
    }
}

// ---- K2: exact re-score + certification ---------------------------------------------------------
// exact -dot in the reference's order for one (query,row): one thread, sequential chain
template <int ORDER>
SG_DEVICE_INLINE float exact_dot(const float *__restrict__ q, const __half *__restrict__ row, uint32_t dim)
{
    if constexpr (ORDER == SG_ORDER_SCALAR) {
        float sum = 0.0f;
        for (uint32_t i = 0; i < dim; i += 4) {
            float p0 = __fmul_rn(q[i], __half2float(row[i])), p1 = __fmul_rn(q[i + 1], __half2float(row[i + 1]));
            float p2 = __fmul_rn(q[i + 2], __half2float(row[i + 2])), p3 = __fmul_rn(q[i + 3], __half2float(row[i + 3]));
            sum = __fadd_rn(sum, __fadd_rn(__fadd_rn(__fadd_rn(p0, p1), p2), p3));
        }
        return sum;
    } else {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t i = 0; i < dim; i += 8)
#pragma unroll
            for (int t = 0; t < 8; ++t) acc[t] = __fmaf_rn(q[i + t], __half2float(row[i + t]), acc[t]);
        float r = __fadd_rn(acc[0], acc[1]);
#pragma unroll
        for (int t = 2; t < 8; ++t) r = __fadd_rn(r, acc[t]);
        return r;
    }
}

// one CTA (256 threads) per query
__global__ void __launch_bounds__(256)
finalize_tc_kernel(const FinalizeArgs f)
{
    __shared__ uint64_t keys[TC_CAND_CAP];      // heads, then surviving approx keys, then exact keys
    __shared__ float q_s[2048];
    __shared__ uint32_t s_flag, s_n;
    const uint32_t q = blockIdx.x;
    const uint32_t L = f.n_lists;
    const uint64_t *ql = f.lists + (size_t)q * L * 32;   // every list is 32 slots, the first f.kk are used
    for (uint32_t i = threadIdx.x; i < f.dim; i += blockDim.x) q_s[i] = f.queries[(size_t)q * f.dim + i];
    // (1) threshold from the list heads: at least kk candidates (kk distinct heads) have key <= T0,
    //     so every member of the global top-kk has key <= T0.
    const uint32_t MH = next_pow2(L < 2 ? 2 : L);
    for (uint32_t i = threadIdx.x; i < MH; i += blockDim.x) keys[i] = i < L ? ql[(size_t)i * 32] : KEY_SENTINEL;
    if (threadIdx.x == 0) {
        s_flag = 0;
        s_n = 0;
    }
    __syncthreads();
    block_bitonic_sort(keys, MH);
    const uint64_t T0 = L >= f.kk ? keys[f.kk - 1] : KEY_SENTINEL;
    __syncthreads();
    // (2) gather every entry with key <= T0 -- all threads, coalesced over the [L][32] list block
    for (uint32_t e = threadIdx.x; e < L * 32; e += blockDim.x) {
        if ((e & 31) >= f.kk) continue;
        const uint64_t key = ql[e];
        if (key != KEY_SENTINEL && key <= T0) {
            const uint32_t slot = atomicAdd(&s_n, 1u);
            if (slot < TC_CAND_CAP) keys[slot] = key;
        }
    }
    __syncthreads();
    uint32_t n = s_n;
    const bool overflow = n > TC_CAND_CAP;
    if (overflow) n = TC_CAND_CAP;
    const uint32_t M = next_pow2(n < 2 ? 2 : n);
    for (uint32_t i = n + threadIdx.x; i < M; i += blockDim.x) keys[i] = KEY_SENTINEL;
    if (threadIdx.x == 0 && overflow) s_flag = 1u;
    __syncthreads();
    block_bitonic_sort(keys, M);                 // ascending approx key: best approximate first
    const uint32_t kk = n < f.kk ? n : f.kk;     // the global top-KK by approximate score
    // weakest approximate similarity that was kept == bound on everything that was dropped
    const bool dropped_any = (f.n_live > kk);
    const float weakest_approx_sim = kk ? -key32_to_dist((uint32_t)(keys[kk - 1] >> 32)) : 0.0f;
    __syncthreads();
    // exact re-score: stage the kk candidate rows in shared memory (coalesced), then one thread per
    // candidate walks the reference's sequential chain.  Row pitch dim/2 + 1 words: conflict-free.
    extern __shared__ __align__(16) uint8_t rows_raw[];
    uint32_t *rows_s = reinterpret_cast<uint32_t *>(rows_raw);
    const uint32_t pitch = f.dim / 2 + 1;
    for (uint32_t c = 0; c < kk; ++c) {
        const uint32_t row = (uint32_t)(keys[c] & 0xFFFFFFFFu);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(reinterpret_cast<const __half *>(f.rows) + (size_t)row * f.dim);
        for (uint32_t w = threadIdx.x; w < f.dim / 2; w += blockDim.x) rows_s[(size_t)c * pitch + w] = src[w];
    }
    __syncthreads();
    uint64_t exact_key = KEY_SENTINEL;
    if (threadIdx.x < kk) {
        const uint32_t row = (uint32_t)(keys[threadIdx.x] & 0xFFFFFFFFu);
        const __half *rp = reinterpret_cast<const __half *>(rows_s + (size_t)threadIdx.x * pitch);
        float dot = f.order == SG_ORDER_AVX2 ? exact_dot<SG_ORDER_AVX2>(q_s, rp, f.dim)
                                             : exact_dot<SG_ORDER_SCALAR>(q_s, rp, f.dim);
        exact_key = make_key(order_key32(neg_dot_to_dist(dot)), row);
    }
    __syncthreads();
    const uint32_t M2 = next_pow2(kk < 2 ? 2 : kk);
    if (threadIdx.x < M2) keys[threadIdx.x] = exact_key;
    __syncthreads();
    block_bitonic_sort(keys, M2);
    const uint32_t out_n = kk < f.k ? kk : f.k;
    // certification: every dropped row has approx sim <= weakest_approx_sim, hence exact sim <=
    // weakest_approx_sim + E; it cannot enter (or tie into) the top k if kth_exact_sim is larger.
    if (threadIdx.x == 0 && dropped_any) {
        bool ok = false;
        if (out_n == f.k) {
            const float kth_exact_sim = -key32_to_dist((uint32_t)(keys[f.k - 1] >> 32));
            const float R = *f.max_row_norm;
            const float E = f.err[q] * R + f.c_acc * f.qnorm[q] * R;
            ok = kth_exact_sim > weakest_approx_sim + E;     // NaN compares false -> flagged
        }
        if (!ok) s_flag = 1u;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < f.k; i += blockDim.x) {
        const bool valid = i < out_n;
        const uint64_t key = valid ? keys[i] : KEY_SENTINEL;
        f.ids_out[(size_t)q * f.k + i] = valid ? (uint32_t)(key & 0xFFFFFFFFu) + f.id_base : 0xFFFFFFFFu;
        f.dist_out[(size_t)q * f.k + i] = valid ? key32_to_dist((uint32_t)(key >> 32)) : __uint_as_float(0x7F800000u);
    }
    if (threadIdx.x == 0) {
        f.counts_out[q] = out_n;
        f.flags[q] = s_flag;
        if (s_flag) {
            const uint32_t slot = atomicAdd(f.n_flagged, 1u);   // arrival order is irrelevant: slots are independent
            f.flag_list[slot] = q;
            atomicAdd(f.flag_total, 1ull);
        }
    }
}

}  // namespace

size_t tc_smem_bytes(uint32_t dim, uint32_t stages, uint32_t kk)
{
    size_t s = (size_t)(dim / TC_KB) * TC_M * 128;          // A
    s += (size_t)stages * TC_STAGE_BYTES;                    // B ring
    (void)kk;
    s += (size_t)32 * TC_M * 8;                              // lists: 32 slots per query
    s += (size_t)(2 * stages + 2 * TC_ACC_STAGES + 1) * 8 + 16;
    return s;
}

cudaError_t launch_prep_queries(const float *d_q, uint32_t b, uint32_t b_pad, uint32_t dim, __half *d_qh, float *d_err,
                                float *d_qnorm, uint32_t *d_gtau, uint32_t *d_flags, cudaStream_t st)
{
    prep_queries_kernel<<<b_pad, 128, 0, st>>>(d_q, b, b_pad, dim, d_qh, d_err, d_qnorm, d_gtau, d_flags);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_scan_tc(const CUtensorMap &tmap_rows, const CUtensorMap &tmap_q, const TcArgs &a, uint32_t grid,
                           size_t smem, cudaStream_t st)
{
    static size_t smem_set_dev[64] = {0};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    size_t &smem_set = smem_set_dev[dev_ & 63];
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        smem_set = smem;
    }
    scan_tc_kernel<<<grid, TC_THREADS, smem, st>>>(tmap_rows, tmap_q, a);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_scan_tc2(const CUtensorMap &tmap_rows, const CUtensorMap &tmap_q, const TcArgs &a, uint32_t grid,
                            size_t smem, cudaStream_t st)
{
    static size_t smem_set_dev[64] = {0};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    size_t &smem_set = smem_set_dev[dev_ & 63];
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(scan_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        smem_set = smem;
    }
    scan_tc2_kernel<<<grid, TC_THREADS, smem, st>>>(tmap_rows, tmap_q, a);   // __cluster_dims__(2,1,1): grid must be even
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_tau_init(const float *d_maxima, uint32_t b_pad, uint32_t n_cta, uint32_t kk, uint32_t *d_gtau,
                            cudaStream_t st)
{
    if (n_cta > 256) return cudaErrorInvalidValue;
    tau_init_kernel<<<b_pad, 256, 0, st>>>(d_maxima, n_cta, kk, d_gtau);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_finalize_tc(const FinalizeArgs &f, uint32_t b, cudaStream_t st)
{
    const size_t smem = (size_t)f.kk * (f.dim / 2 + 1) * 4;
    if (smem > 32 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(finalize_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    finalize_tc_kernel<<<b, 256, smem, st>>>(f);
    count_launch();
    return cudaGetLastError();
}

}  // namespace sg
