// scan_tc.cu -- the tensor-core corpus sweep ("path B") for batched queries.
//
// For b >= 3 queries the row loop of VamanaIndex::brute_force_search (reference
// src/vector_db/vamana.rs:1167-1188) is a [b x dim] . [dim x n_rows] contraction: at b = 256
// it needs ~2 PFLOP/s to keep up with HBM, which only tcgen05 can deliver.  The tensor cores
// see fp16(query) x fp16(row) with f32 accumulation, i.e. APPROXIMATE scores; exactness is
// restored by over-selection + exact re-scoring:
//   K1 (scan_tc_kernel / scan_tc2_kernel): ONE sweep.  Every CTA (pair) first scores one sample
//       tile, the grid meets at a device-side barrier and derives a valid starting threshold per
//       query from the sample maxima; from then on every score at or above the query's threshold
//       is APPENDED to the (query, CTA) candidate region and counted in a per-query score
//       histogram shared by the whole GPU.  "kk rows have been appended at or above bucket edge e"
//       makes e a valid threshold, so the threshold follows the running kk-th best score of ALL
//       rows seen so far by ALL CTAs -- no pre-pass launch, no sorted lists, no atomics that return;
//   K2 (finalize_tc_kernel): per query take the kk = k + slack best approximate candidates,
//       re-score them in EXACTLY the reference's summation order (distance_inline.rs:157-173 /
//       :67-111), sort by (dist, id), and certify with a rigorous error bound E that no row
//       outside the candidate set can reach the top k:  kth_exact_sim > weakest_kept_approx + E.
//       Queries that cannot be certified are flagged and re-run on the exact path A.  Certified
//       results can be pushed straight into the peers' exchange buffers (multi-GPU, sg_push.cuh).
// So the result is bit-identical to path A / the CPU oracle, or the query is flagged.
//
// B200 mapping (one CTA per SM, 256 threads, persistent over corpus tiles):
//   warp 0   TMA producer: the fp16 query block (A operand, 128 x dim, resident for the whole
//            sweep) and a ring of K-blocks of corpus tiles (B operand, 128 rows x 64 elems,
//            SWIZZLE_128B, K-major -- the same boxes path A streams);
//   warp 1   MMA issuer: tcgen05.mma kind::f16, M=128 (queries) x N=128 rows (1 CTA) or
//            cta_group::2 with M=256 x N=256 (CTA pair), K=16 per instruction, accumulators in TMEM;
//   warp 2   TMEM allocator;
//   warps 4-7 epilogue: thread t owns query t = TMEM lane t; tcgen05.ld 32 columns at a time,
//            one max-tree + one vote per chunk; hits (rare) are appended by the owning lane.
// Algorithmic bytes per sweep = n_rows * dim * 2; 1-CTA form: HBM-bound (12 MMA cycles per row
// per SM against ~33 cycles of HBM time); pair form at 256 queries: tensor-bound.
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>

#include <atomic>

#include "sg_internal.h"
#include "sg_push.cuh"
#include "sg_tc.cuh"

namespace sg {
namespace {

constexpr uint32_t TC_M = 128;          // queries per CTA (TMEM lanes)
constexpr uint32_t TC_N = 128;          // corpus rows per CTA tile
constexpr uint32_t TC_KB = 64;          // fp16 elements per K-block (one 128B swizzle row)
constexpr uint32_t TC_STAGE_BYTES = TC_N * 128;   // 16 KB
constexpr uint32_t TC_ACC_STAGES = 4;   // 4 x 128 = 512 TMEM columns (1-CTA form)
constexpr uint32_t TC_THREADS = 384;     // warps 0-3: TMA producer, MMA issuer, TMEM allocator, spare; warps 4-11: epilogue
constexpr uint32_t TC_EPI_THREADS = 256;
constexpr float TC_NEG_INF = -__builtin_huge_valf();

SG_DEVICE_INLINE uint32_t ld_relaxed_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
SG_DEVICE_INLINE uint32_t ld_acquire_gpu_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
SG_DEVICE_INLINE unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
SG_DEVICE_INLINE void epi_bar_sync()      // the 256 epilogue threads only (named barrier 1)
{
    asm volatile("bar.sync 1, 256;" ::: "memory");
}

// tile handled in iteration `it` of a CTA (pair): iteration 0 is the sample pass over the list's first
// tile (maxima only); iterations 1..my_tiles are the sweep proper, which scores that tile again (it is
// still in L2) together with all the others.
SG_DEVICE_INLINE uint64_t tile_of_iter(uint64_t tile0, uint64_t stride, uint64_t it)
{
    return tile0 + (it ? it - 1 : 0) * stride;
}

// The A operand: this CTA's 128 query rows, f32 in global memory -> fp16 in shared memory, written directly in the
// K-major SWIZZLE_128B layout the MMA descriptors expect (K-block kb = 128 rows x 128 B; the 16-byte chunk c of row r
// sits at chunk c ^ (r & 7)).  Rows past the live count are zero.  All threads of the CTA take part; the caller
// fences the generic-proxy stores for the async proxy (tcgen05.mma) and synchronises.
SG_DEVICE_INLINE void stage_queries(const TcArgs &a, uint8_t *a_base, uint32_t q_first, uint32_t nq_local)
{
    const uint32_t cpr = a.dim / 8;                      // 16-byte chunks per row
    const uint32_t a_addr = smem_u32(a_base);
    for (uint32_t idx = threadIdx.x; idx < TC_M * cpr; idx += blockDim.x) {
        const uint32_t r = idx / cpr, c8 = idx - r * cpr, kb = c8 >> 3, c = c8 & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < nq_local) {
            const float4 *src = reinterpret_cast<const float4 *>(a.queries + (size_t)(q_first + r) * a.dim + c8 * 8);
            const float4 f0 = __ldg(src), f1 = __ldg(src + 1);
            const __half2 h0 = __floats2half2_rn(f0.x, f0.y), h1 = __floats2half2_rn(f0.z, f0.w);
            const __half2 h2 = __floats2half2_rn(f1.x, f1.y), h3 = __floats2half2_rn(f1.z, f1.w);
            v.x = *reinterpret_cast<const uint32_t *>(&h0);
            v.y = *reinterpret_cast<const uint32_t *>(&h1);
            v.z = *reinterpret_cast<const uint32_t *>(&h2);
            v.w = *reinterpret_cast<const uint32_t *>(&h3);
        }
        const uint32_t dst = a_addr + kb * (TC_M * 128) + r * 128 + ((c ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
}

// ---- the epilogue shared by the 1-CTA and the CTA-pair kernels -------------------------------------
// EIGHT epilogue warps: two per TMEM lane quarter.  Thread (grp, t) <-> query q_base + t <-> TMEM lane t; group
// grp = 0/1 scans the even/odd 32-column chunks of every accumulator, so a tile's scan time is halved and two
// warps per scheduler hide each other's TMEM-load latency -- the scan must stay well ahead of the MMA pipe (a
// 256-row pair tile is ~3000 tensor cycles).  Each (query, list, grp) owns its own candidate region, so nothing is
// shared between the two threads of a query except the read-only thresholds.
// NCOLS = corpus rows (TMEM columns) per accumulator stage.  tempty_cluster_addr is the shared::cluster address of
// the accumulator-empty barriers (the leader CTA's in pair mode).  n_lists counts REGIONS (2 per CTA / CTA pair).
//
// Threshold validity.  tau_f never exceeds the kk-th best approximate score over the whole corpus:
//   * seed: the sample contributes n_lists * J maxima of DISTINCT 32-row chunks; when at least kk of
//     them are finite, their minimum has kk distinct rows at or above it;
//   * histogram: every appended row increments exactly one bucket, bucket(s) = min(31, floor((s - tau0) / w)).
//     If the buckets >= idx hold kk rows then kk rows have s >= tau0 + (idx - 1) * w: the edge ONE bucket
//     below absorbs every rounding in bucket() and in the edge (w > 1e-5 * scale >> 2^-23 * scale);
//   * gtau: the best threshold any CTA derived, shared with atomicMin on its order key.
// A row that is not appended therefore has a score strictly below the final kk-th best.
template <uint32_t NCOLS, bool DUMP>
SG_DEVICE_INLINE void tc_epilogue(const TcArgs &a, uint32_t *s_ctl, uint32_t tmem_base, uint64_t *tfull_bar,
                                  uint32_t tempty_cluster_addr, uint64_t my_tiles, uint64_t tile0, uint64_t tile_stride,
                                  uint32_t q_base, uint32_t nq_local, uint32_t unit_idx, uint32_t n_lists,
                                  uint32_t acc_stages)
{
    constexpr uint32_t NCH = NCOLS / 32;             // chunks per accumulator; this thread scans grp, grp + 2, ...
    const uint32_t e = threadIdx.x - 128;            // 0..255 over the eight epilogue warps
    const uint32_t grp = e >> 7, t = e & 127;
    const uint32_t wq = ((threadIdx.x >> 5) & 3) * 32;   // first TMEM lane of this warp's quarter
    const bool active = t < nq_local;
    const uint32_t q = q_base + t;
    const uint32_t list_idx = 2 * unit_idx + grp;    // this thread's region among the query's n_lists regions
    const uint32_t lane_base = wq << 16;
    float tau_f = active ? TC_NEG_INF : -TC_NEG_INF; // similarity threshold: -inf (everything) / +inf (never)
    float tau0 = TC_NEG_INF, w = 0.0f, inv_w = 0.0f;
    bool use_hist = false;
    uint32_t my_cnt = 0;                             // entries this thread appended to ITS region
    uint32_t tau_pub = 0xFFFFFFFFu;                  // last threshold key this lane shared
    uint64_t *my_cand = a.cand + ((size_t)q * n_lists + list_idx) * a.cap_l;
    uint32_t *my_hist = a.hist + (size_t)q * TC_HIST;
    uint32_t *gtau = a.gtau + q;

    auto dead_mask = [&](uint64_t row0) -> uint32_t {
        uint32_t dead = 0;
        if (a.deleted && row0 < a.n_rows) dead = __ldg(a.deleted + (row0 >> 5));
        if (row0 + 32 > a.n_rows) dead |= row0 >= a.n_rows ? 0xFFFFFFFFu : ~((1u << (uint32_t)(a.n_rows - row0)) - 1u);
        return dead;
    };
    // maximum of a 32-column chunk; g8[i] = maximum of columns 8i..8i+7 (by-products of the same tree)
    auto chunk_max = [&](const uint32_t (&r)[32], float (&g8)[4]) -> float {
        float m0 = fmaxf(fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1])), fmaxf(__uint_as_float(r[2]), __uint_as_float(r[3])));
        float m1 = fmaxf(fmaxf(__uint_as_float(r[4]), __uint_as_float(r[5])), fmaxf(__uint_as_float(r[6]), __uint_as_float(r[7])));
        float m2 = fmaxf(fmaxf(__uint_as_float(r[8]), __uint_as_float(r[9])), fmaxf(__uint_as_float(r[10]), __uint_as_float(r[11])));
        float m3 = fmaxf(fmaxf(__uint_as_float(r[12]), __uint_as_float(r[13])), fmaxf(__uint_as_float(r[14]), __uint_as_float(r[15])));
        float m4 = fmaxf(fmaxf(__uint_as_float(r[16]), __uint_as_float(r[17])), fmaxf(__uint_as_float(r[18]), __uint_as_float(r[19])));
        float m5 = fmaxf(fmaxf(__uint_as_float(r[20]), __uint_as_float(r[21])), fmaxf(__uint_as_float(r[22]), __uint_as_float(r[23])));
        float m6 = fmaxf(fmaxf(__uint_as_float(r[24]), __uint_as_float(r[25])), fmaxf(__uint_as_float(r[26]), __uint_as_float(r[27])));
        float m7 = fmaxf(fmaxf(__uint_as_float(r[28]), __uint_as_float(r[29])), fmaxf(__uint_as_float(r[30]), __uint_as_float(r[31])));
        g8[0] = fmaxf(m0, m1); g8[1] = fmaxf(m2, m3); g8[2] = fmaxf(m4, m5); g8[3] = fmaxf(m6, m7);
        return fmaxf(fmaxf(g8[0], g8[1]), fmaxf(g8[2], g8[3]));
    };

    const bool tr = a.trace != nullptr && e == 0;
    unsigned long long *trp = a.trace + (size_t)blockIdx.x * 16;
    if (tr) trp[0] = gtimer();                       // epilogue entered (prologue: barriers, TMEM, query block)
    uint64_t it = 1;
    if constexpr (!DUMP) {
        // ===== sample pass (iteration 0): the J best chunk maxima of this thread's chunks of the unit's first tile =====
        float m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) m[i] = TC_NEG_INF;
        if (my_tiles) {
            mbar_wait(smem_u32(tfull_bar), 0);
            tc_fence_after();
            if (tr) trp[1] = gtimer();               // sample tile's accumulator ready
#pragma unroll 1
            for (uint32_t c = grp; c < NCH; c += 2) {
                uint32_t r[32];
                tmem_ld32(tmem_base + lane_base + c * 32, r);
                tmem_ld_wait();
                const uint32_t dead = dead_mask(tile0 * NCOLS + c * 32);
                if (dead) {
#pragma unroll
                    for (uint32_t j = 0; j < 32; ++j)
                        if ((dead >> j) & 1u) r[j] = __float_as_uint(TC_NEG_INF);
                }
                float g8s[4];
                float mx = chunk_max(r, g8s);
                if (!(mx == mx)) mx = TC_NEG_INF;        // an all-NaN chunk must not enter (it would duplicate entries)
#pragma unroll
                for (int i = 0; i < 4; ++i) {            // insert into the descending list
                    const float hi = fmaxf(m[i], mx);
                    mx = fminf(m[i], mx);
                    m[i] = hi;
                }
            }
            tc_fence_before();
            mbar_arrive_cluster(tempty_cluster_addr);        // accumulator stage 0 is free again
        }
        // the unit's sample = the 4 best chunk maxima over BOTH groups' chunks: group 1 hands its list to group 0
        float4 *xchg = reinterpret_cast<float4 *>(s_ctl + 4);            // [128] float4, shared by the two groups
        if (grp == 1) xchg[t] = make_float4(m[0], m[1], m[2], m[3]);
        epi_bar_sync();
        if (grp == 0) {
            const float4 o = xchg[t];
            const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float mx = ov[jj];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float hi = fmaxf(m[i], mx);
                    mx = fminf(m[i], mx);
                    m[i] = hi;
                }
            }
            if (active) {
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j)
                    if (j < a.J) a.maxima[(size_t)q * a.nv_pad + unit_idx * a.J + j] = m[j];
            }
        }
        // CTA 0 resets this query tile's per-search state (there is no separate "prepare" launch): nothing below is
        // touched by any CTA before the meeting point, and the fence + barrier publish the stores
        if (blockIdx.x == 0) {
            const uint32_t q0 = a.q_tile * a.qtile, qn = a.qtile;
            for (uint32_t i = e; i < qn * TC_HIST; i += TC_EPI_THREADS) a.hist[(size_t)q0 * TC_HIST + i] = 0;
            for (uint32_t i = e; i < qn; i += TC_EPI_THREADS) {
                a.gtau[q0 + i] = 0xFFFFFFFFu;
                for (uint32_t j = a.nv; j < a.nv_pad; ++j) a.maxima[(size_t)(q0 + i) * a.nv_pad + j] = TC_NEG_INF;
            }
            if (e == 0 && a.q_tile == 0) *a.nflag = 0;
        }
        // device-wide meeting point: every CTA of the grid publishes its sample, then reads all of them.  The arrival
        // counter is never reset: this launch waits for the value the host computed for it (wrap-safe compare).
        if (tr) trp[2] = gtimer();                   // sample scanned
        epi_bar_sync();                              // the CTA's stores are ordered before thread 0's fence + arrival
        if (e == 0) {
            __threadfence();
            atomicAdd(a.arrive, 1u);
            const long long t0 = clock64();
            uint32_t ok = 1;
            while ((int32_t)(ld_acquire_gpu_u32(a.arrive) - a.arrive_target) < 0) {
                if (clock64() - t0 > 2000000000ll) {         // ~1 s: the grid is not co-resident (should not happen:
                    ok = 0;                                  // cooperative launch); degrade to "no threshold" -> the
                    break;                                   // queries overflow, get flagged and re-run on path A
                }
            }
            __threadfence();
            s_ctl[0] = ok;
        }
        epi_bar_sync();                              // (the sample values are read with ld.global.cg: L2, never a stale L1 line)
        if (tr) trp[3] = gtimer();                   // every CTA has arrived
        // ===== seed threshold: the best PROVABLE lower bound on the kk-th best score this sample gives =====
        // Each sample value is the maximum of a DISTINCT 32-row chunk, so "c values >= x" means c distinct rows score >= x.
        // Seven levels mn + j (mean - mn) / 2 are counted, the highest with kk values wins, and seven sub-levels refine
        // inside the winning interval (the first tiles of a sweep run under this seed alone, and every chunk that reaches
        // it takes the slow path).  The levels hang off the MEAN of the sample, which -- unlike its maximum -- is not dragged
        // away by a near-duplicate of the query that happens to sit in the sample.
        const uint32_t nv4 = a.nv_pad / 4;
        const float4 *mp4 = reinterpret_cast<const float4 *>(a.maxima + (size_t)q * a.nv_pad);
        const bool seed_on = active && s_ctl[0];
        float mn = -TC_NEG_INF, sum = 0.0f, mean = 0.0f, mn0 = 0.0f;
        uint32_t nfin = 0;
        float lv[7];
        uint32_t cn[7];
        auto stats = [&](const float4 &v4) {
            const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (vv[c] > TC_NEG_INF) {
                    ++nfin;
                    mn = fminf(mn, vv[c]);
                    sum += vv[c];
                }
            }
        };
        auto set_levels = [&](float base, float step) {
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                lv[j] = base + (float)(j + 1) * step;
                cn[j] = 0;
            }
        };
        auto count = [&](const float4 &v4) {
#pragma unroll
            for (int j = 0; j < 7; ++j)                      // -inf / NaN compare false
                cn[j] += (v4.x >= lv[j] ? 1u : 0u) + (v4.y >= lv[j] ? 1u : 0u) + (v4.z >= lv[j] ? 1u : 0u) +
                         (v4.w >= lv[j] ? 1u : 0u);
        };
        auto best_level = [&](float base) {                  // levels ascend: the last one that qualifies wins
            float lo = base;
#pragma unroll
            for (int j = 0; j < 7; ++j)
                if (cn[j] >= a.kk) lo = lv[j];
            return lo;
        };
        constexpr uint32_t NC = 10;                          // float4 per thread held in registers
        if (nv4 <= 2 * NC) {
            // The common shape (k <= 24: one value per unit, <= 80 values per query).  BOTH warp groups work on it: thread
            // (query t, group g) holds half of the query's values in registers -- one round trip to L2 instead of two
            // passes of dependent loads by one group (measured 7 us of every sweep) -- and the halves are combined through
            // the exchange slots (.xy = group 0's words, .zw = group 1's).
            float4 cv[NC];
#pragma unroll
            for (uint32_t i = 0; i < NC; ++i) {
                const uint32_t idx = grp * NC + i;
                cv[i] = (seed_on && idx < nv4) ? __ldcg(mp4 + idx) : make_float4(TC_NEG_INF, TC_NEG_INF, TC_NEG_INF, TC_NEG_INF);
            }
#pragma unroll
            for (uint32_t i = 0; i < NC; ++i) stats(cv[i]);
            float2 *slot = reinterpret_cast<float2 *>(xchg + t);                 // [0] group 0, [1] group 1
            auto swap2 = [&](float x, float y) -> float2 {   // publish two words, return the other group's
                slot[grp] = make_float2(x, y);
                epi_bar_sync();
                const float2 o = slot[grp ^ 1];
                epi_bar_sync();
                return o;
            };
            {
                const float2 o = swap2(mn, __uint_as_float(nfin));
                const float2 o2 = swap2(sum, 0.0f);
                mn = fminf(mn, o.x);
                nfin += __float_as_uint(o.y);
                sum = grp == 0 ? sum + o2.x : o2.x + sum;    // the same operand order in both groups: identical seeds
            }
            const bool enough = seed_on && nfin >= a.kk;
            auto count_all = [&]() {
#pragma unroll
                for (uint32_t i = 0; i < NC; ++i) count(cv[i]);
                // 7 counts <= 80 each: one byte apiece, two words
                const uint32_t w0 = cn[0] | (cn[1] << 8) | (cn[2] << 16) | (cn[3] << 24), w1 = cn[4] | (cn[5] << 8) | (cn[6] << 16);
                const float2 o = swap2(__uint_as_float(w0), __uint_as_float(w1));
                const uint32_t o0 = __float_as_uint(o.x), o1 = __float_as_uint(o.y);
                cn[0] += o0 & 0xFFu; cn[1] += (o0 >> 8) & 0xFFu; cn[2] += (o0 >> 16) & 0xFFu; cn[3] += o0 >> 24;
                cn[4] += o1 & 0xFFu; cn[5] += (o1 >> 8) & 0xFFu; cn[6] += (o1 >> 16) & 0xFFu;
            };
            mean = nfin ? sum / (float)nfin : 0.0f;
            set_levels(mn, 0.5f * (mean - mn));
            count_all();
            const float lo = best_level(mn);
            mn0 = mn;
            set_levels(lo, 0.0625f * (mean - mn));
            count_all();
            if (enough) {
                tau0 = best_level(lo);
                tau_f = tau0;
            }
        } else {
            // many values per query (large k): group 0 alone, two passes of 16-byte loads
            if (grp == 0 && seed_on) {
#pragma unroll 8
                for (uint32_t i = 0; i < nv4; ++i) stats(__ldcg(mp4 + i));
                if (nfin >= a.kk) {
                    mean = sum / (float)nfin;
                    mn0 = mn;
                    set_levels(mn, 0.5f * (mean - mn));
#pragma unroll 8
                    for (uint32_t i = 0; i < nv4; ++i) count(__ldcg(mp4 + i));
                    tau0 = best_level(mn);
                    tau_f = tau0;
                }
            }
        }
        if (seed_on && nfin >= a.kk && (grp == 0 || nv4 <= 2 * NC)) {
            // histogram over [tau0, tau0 + 3 (mean - mn)]: the final threshold lies well beyond the sample's typical maximum
            // (the sample is ~1/500 of a 10M-row corpus).  Scores beyond the range land in the last bucket: the threshold
            // then stops one bucket short of them, which only costs a few extra candidates.
            const float range = 3.0f * (mean - mn0);
            w = range * (1.0f / TC_HIST);
            if (w > 1.0e-5f * fmaxf(fabsf(tau0), fabsf(tau0 + range)) && w < -TC_NEG_INF) {
                inv_w = 1.0f / w;
                use_hist = true;
            }
        }
        if (nv4 > 2 * NC) {
            // group 0 computed the query's seed; group 1 takes it over (identical thresholds and bucket mapping)
            if (grp == 0) xchg[t] = make_float4(tau0, w, inv_w, use_hist ? 1.0f : 0.0f);
            epi_bar_sync();
            if (grp == 1 && active) {
                const float4 sd = xchg[t];
                tau0 = sd.x; w = sd.y; inv_w = sd.z; use_hist = sd.w != 0.0f;
                tau_f = tau0;
            }
        }
    }

    // ===== the sweep proper =====
    //  * the threshold inputs (shared key, every 4th tile the histogram) are requested ONE TILE AHEAD and consumed
    //    at the top of the next iteration: their L2 latency never sits between "accumulator ready" and the scan;
    //  * dead / out-of-range columns are NOT neutralised on the fast path: they can at worst trigger the slow
    //    path, which masks them out -- the fast path is one 3-input max tree and one vote per 32 columns.
    uint32_t accn = (uint32_t)((DUMP ? it - 1 : it) % acc_stages), acc_ph = (uint32_t)(((DUMP ? it - 1 : it) / acc_stages) & 1);
    uint32_t g_pend = 0xFFFFFFFFu;
    uint4 h_pend[TC_HIST / 4];
    bool h_valid = false;
    if (tr) trp[4] = gtimer();                       // seeds computed
    for (; it <= my_tiles; ++it) {
        const uint64_t tile = tile_of_iter(tile0, tile_stride, it);
        if constexpr (!DUMP) {
            if (h_valid) {
                // highest bucket idx whose tail holds kk rows
                uint32_t cum = 0;
                int idx = -1;
#pragma unroll
                for (int i = TC_HIST / 4 - 1; i >= 0; --i) {
                    cum += h_pend[i].w; if (idx < 0 && cum >= a.kk) idx = 4 * i + 3;
                    cum += h_pend[i].z; if (idx < 0 && cum >= a.kk) idx = 4 * i + 2;
                    cum += h_pend[i].y; if (idx < 0 && cum >= a.kk) idx = 4 * i + 1;
                    cum += h_pend[i].x; if (idx < 0 && cum >= a.kk) idx = 4 * i;
                }
                if (idx >= 2) tau_f = fmaxf(tau_f, tau0 + (float)(idx - 1) * w);
            }
            if (g_pend != 0xFFFFFFFFu) tau_f = fmaxf(tau_f, -key32_to_dist(g_pend));
        }
        mbar_wait(smem_u32(tfull_bar + accn), acc_ph);
        tc_fence_after();
        if constexpr (!DUMP) {
            if (active) g_pend = ld_relaxed_u32(gtau);       // for the NEXT tile
            h_valid = active && use_hist && (((uint32_t)it + 1 + list_idx) & 3u) == 0;   // (refreshing EVERY early tile was measured slower)
            if (h_valid) {
                const uint4 *hp = reinterpret_cast<const uint4 *>(my_hist);
#pragma unroll
                for (uint32_t i = 0; i < TC_HIST / 4; ++i) h_pend[i] = __ldcg(hp + i);
            }
        }
        const uint32_t tbase = tmem_base + lane_base + accn * NCOLS;
#pragma unroll 1
        for (uint32_t c = grp; c < NCH; c += 2) {
            uint32_t r[32];
            tmem_ld32(tbase + c * 32, r);
            tmem_ld_wait();
            const uint64_t row0 = tile * NCOLS + c * 32;     // 32-aligned: one bitmap word
            if constexpr (DUMP) {
                if (active) {
#pragma unroll
                    for (uint32_t j = 0; j < 32; ++j) {
                        const uint64_t row = row0 + j;
                        if (row >= a.dump_row0 && row < a.dump_row0 + a.dump_rows && row < a.n_rows)
                            a.dump[(size_t)(row - a.dump_row0) * a.nq + q] = __uint_as_float(r[j]);   // [row][query]: coalesced
                    }
                }
                continue;
            }
            float g8[4];
            const float mx = chunk_max(r, g8);
            if (__any_sync(0xFFFFFFFFu, mx >= tau_f)) {
                // slow path: the owning lane appends its hits.  The max tree already holds the maximum of every 8-column
                // group: only groups that reach the threshold are looked at column by column (under a weak early threshold
                // most chunks come here, so this path's fixed cost is what the first tiles of a sweep pay)
                const uint32_t dead = dead_mask(row0);
#pragma unroll
                for (uint32_t gi = 0; gi < 4; ++gi) {
                    if (g8[gi] >= tau_f) {
#pragma unroll
                        for (uint32_t jj = 0; jj < 8; ++jj) {
                            const uint32_t j = 8 * gi + jj;
                            const float sc = __uint_as_float(r[j]);
                            if (sc >= tau_f && !((dead >> j) & 1u)) {
                                if (my_cnt < a.cap_l) my_cand[my_cnt] = make_key(order_key32(-sc), (uint32_t)(row0 + j));
                                ++my_cnt;
                                if (use_hist) {
                                    const float fb = (sc - tau0) * inv_w;
                                    const uint32_t bi = fb >= (float)(TC_HIST - 1) ? TC_HIST - 1 : (uint32_t)fb;
                                    atomicAdd(my_hist + bi, 1u);     // result unused: a fire-and-forget RED
                                }
                            }
                        }
                    }
                }
            }
        }
        tc_fence_before();
        mbar_arrive_cluster(tempty_cluster_addr + 8 * accn);  // all epilogue threads release the accumulator
        if (tr && it <= 32 && (it & (it - 1)) == 0) trp[5 + (31 - __clz(it))] = gtimer();   // tiles 1,2,4,8,16,32 scanned
        if (++accn == acc_stages) {
            accn = 0;
            acc_ph ^= 1;
        }
        if constexpr (!DUMP) {
            if (active && tau_f > TC_NEG_INF) {              // share the threshold GPU-wide when it moved
                const uint32_t k32 = order_key32(-tau_f);
                if (k32 < tau_pub) {
                    tau_pub = k32;
                    atomicMin(gtau, k32);
                }
            }
        }
    }
    if constexpr (!DUMP) {
        if (active) a.cnts[(size_t)q * n_lists + list_idx] = my_cnt;
    }
    if (tr) trp[11] = gtimer();                      // sweep done
}

// ---- K1: the sweep, one CTA per SM (128 queries per sweep) ------------------------------------------
template <bool DUMP>
__global__ void __launch_bounds__(TC_THREADS, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tmap_rows, const TcArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nkb = a.dim / TC_KB;          // K-blocks per row
    const uint32_t S = a.stages;

    // carve-up: A block | B ring | barriers | control word | tmem ptr
    uint8_t *a_base = smem;                                         // nkb x (128 x 128 B)
    uint8_t *b_base = a_base + (size_t)nkb * TC_M * 128;            // S x 16 KB
    uint64_t *bars = reinterpret_cast<uint64_t *>(b_base + (size_t)S * a.kbs * TC_STAGE_BYTES);
    uint64_t *full_bar = bars;                       // [S]
    uint64_t *empty_bar = bars + S;                  // [S]
    uint64_t *tfull_bar = bars + 2 * S;              // [TC_ACC_STAGES]
    uint64_t *tempty_bar = tfull_bar + TC_ACC_STAGES;  // [TC_ACC_STAGES]
    uint64_t *a_bar = tempty_bar + TC_ACC_STAGES;    // [1]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_bar + 1);
    uint32_t *s_ctl = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~static_cast<uintptr_t>(15));   // [4] control + [128] float4

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < S; ++i) {
            mbar_init(smem_u32(full_bar + i), 1);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        for (uint32_t i = 0; i < TC_ACC_STAGES; ++i) {
            mbar_init(smem_u32(tfull_bar + i), 1);
            mbar_init(smem_u32(tempty_bar + i), TC_EPI_THREADS);
        }
        mbar_init(smem_u32(a_bar), 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmap_rows);
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
    pdl_wait();                                              // everything above overlapped the previous step's tail
    stage_queries(a, a_base, a.q_tile * TC_M, a.nq);         // the resident A operand (fp16 queries)
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint64_t n_tiles = (a.n_rows + TC_N - 1) / TC_N;
    const uint64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint64_t it0 = (DUMP || my_tiles == 0) ? 1 : 0;   // iteration 0 = the sample pass

    // A ring stage holds KBS consecutive K-blocks of one tile (KBS = 2 when the row has an even number of them): one
    // barrier round trip, one election and one commit per stage instead of per K-block.  The MMA issue loop is a
    // single thread's dependent instruction chain; per K-block it cost more than the 4 MMAs it issues take to
    // execute (measured: 630-750 cycles against 512 / 256), i.e. the issue loop, not HBM or the tensor pipe, set the
    // pace.  Halving its trips per tile puts it back under the MMA time.
    const uint32_t KBS = a.kbs, n_grp = nkb / KBS, stage_bytes = KBS * TC_STAGE_BYTES;
    if (warp == 0) {
        // ===== TMA producer (warp-uniform loop, one elected lane issues) =====
        uint32_t st = 0, ph = 0;
        for (uint64_t it = it0; it <= my_tiles; ++it) {
            const uint64_t tile = tile_of_iter(blockIdx.x, gridDim.x, it);
            for (uint32_t g = 0; g < n_grp; ++g) {
                mbar_wait(smem_u32(empty_bar + st), ph ^ 1);
                if (elect_one_sync()) {
                    const uint32_t fb = smem_u32(full_bar + st), dst = smem_u32(b_base) + st * stage_bytes;
                    mbar_arrive_expect_tx(fb, stage_bytes);
                    for (uint32_t j = 0; j < KBS; ++j)
                        tma_load_2d(dst + j * TC_STAGE_BYTES, &tmap_rows, fb, (int32_t)((g * KBS + j) * TC_KB),
                                    (int32_t)(tile * TC_N));
                }
                __syncwarp();
                if (++st == S) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the (uniform) loop, one elected lane issues =====
        constexpr uint32_t idesc = make_idesc_f16(TC_M, TC_N);
        uint32_t st = 0, ph = 0;
        uint32_t acc = 0, acc_ph = 0;
        // descriptors differ only in their low word (address >> 4): keep that as a running integer
        const uint64_t desc_hi = make_sw128_desc(0);
        const uint32_t a_lo = smem_u32(a_base) >> 4, b_lo = smem_u32(b_base) >> 4;
        for (uint64_t it = it0; it <= my_tiles; ++it) {
            mbar_wait(smem_u32(tempty_bar + acc), acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * TC_N;
            for (uint32_t g = 0; g < n_grp; ++g) {
                mbar_wait(smem_u32(full_bar + st), ph);
                tc_fence_after();
                if (elect_one_sync()) {
                    for (uint32_t j = 0; j < KBS; ++j) {
                        const uint64_t da = desc_hi | (uint64_t)(a_lo + (g * KBS + j) * (TC_M * 128 / 16));
                        const uint64_t db = desc_hi | (uint64_t)(b_lo + (st * stage_bytes + j * TC_STAGE_BYTES) / 16);
#pragma unroll
                        for (uint32_t k = 0; k < TC_KB / 16; ++k)   // advance 32 B (= 16 fp16) inside the swizzle row
                            umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (g | j | k) ? 1u : 0u);
                    }
                    umma_commit(smem_u32(empty_bar + st));     // frees the stage when these MMAs retire
                    if (g + 1 == n_grp) umma_commit(smem_u32(tfull_bar + acc));   // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++st == S) { st = 0; ph ^= 1; }
            }
            if (++acc == TC_ACC_STAGES) { acc = 0; acc_ph ^= 1; }
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread t <-> query t <-> TMEM lane t =====
        tc_epilogue<TC_N, DUMP>(a, s_ctl, tmem_base, tfull_bar, smem_u32(tempty_bar), my_tiles, /*tile0=*/blockIdx.x,
                                /*tile_stride=*/gridDim.x, /*q_base=*/a.q_tile * TC_M, /*nq_local=*/a.nq,
                                /*unit_idx=*/blockIdx.x, /*n_lists=*/2 * gridDim.x, TC_ACC_STAGES);
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
template <bool DUMP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
scan_tc2_kernel(const __grid_constant__ CUtensorMap tmap_rows, const TcArgs a)
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
    uint64_t *bars = reinterpret_cast<uint64_t *>(b_base + (size_t)S * a.kbs * TC_STAGE_BYTES);
    uint64_t *full_bar = bars;
    uint64_t *empty_bar = bars + S;
    uint64_t *tfull_bar = bars + 2 * S;
    uint64_t *tempty_bar = tfull_bar + TC_ACC_STAGES;               // same carve-up as the 1-CTA form (ACC2 used)
    uint64_t *a_bar = tempty_bar + TC_ACC_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_bar + 1);
    uint32_t *s_ctl = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~static_cast<uintptr_t>(15));   // [4] control + [128] float4

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < S; ++i) {
            mbar_init(smem_u32(full_bar + i), 2);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        for (uint32_t i = 0; i < ACC2; ++i) {
            mbar_init(smem_u32(tfull_bar + i), 1);
            mbar_init(smem_u32(tempty_bar + i), 2 * TC_EPI_THREADS);
        }
        mbar_init(smem_u32(a_bar), 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmap_rows);
    }
    cluster_sync_all();                                  // both CTAs resident, barriers initialised
    if (warp == 2) tmem_alloc2(smem_u32(tmem_slot), 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const uint64_t n_tiles = (a.n_rows + N2 - 1) / N2;
    const uint64_t my_tiles = cluster_id < n_tiles ? (n_tiles - cluster_id + n_clusters - 1) / n_clusters : 0;
    const uint64_t it0 = (DUMP || my_tiles == 0) ? 1 : 0;

    // resident A operand: this CTA's 128 queries, converted in place; then the pair synchronises so the leader's MMA
    // thread may read the peer's A through the pair datapath
    pdl_wait();                                              // everything above overlapped the previous step's tail
    {
        const uint32_t qf = a.q_tile * 256 + rank * TC_M;
        stage_queries(a, a_base, qf, a.nq > rank * TC_M ? (a.nq - rank * TC_M < TC_M ? a.nq - rank * TC_M : TC_M) : 0);
        fence_proxy_async();
    }
    cluster_sync_all();

    const uint32_t KBS = a.kbs, n_grp = nkb / KBS, stage_bytes = KBS * TC_STAGE_BYTES;   // see scan_tc_kernel
    if (warp == 0) {
        // ===== TMA producer (both CTAs): my half of every tile; warp-uniform loop, one elected lane issues =====
        uint32_t st = 0, ph = 0;
        for (uint64_t it = it0; it <= my_tiles; ++it) {
            const uint64_t tile = tile_of_iter(cluster_id, n_clusters, it);
            for (uint32_t g = 0; g < n_grp; ++g) {
                mbar_wait(smem_u32(empty_bar + st), ph ^ 1);
                if (elect_one_sync()) {
                    const uint32_t lfull = mapa_u32(smem_u32(full_bar + st), 0), dst = smem_u32(b_base) + st * stage_bytes;
                    if (leader) mbar_arrive_expect_tx(smem_u32(full_bar + st), 2 * stage_bytes);
                    else mbar_arrive_cluster(lfull);
                    for (uint32_t j = 0; j < KBS; ++j)
                        tma_load_2d_2cta(dst + j * TC_STAGE_BYTES, &tmap_rows, lfull, (int32_t)((g * KBS + j) * TC_KB),
                                         (int32_t)(tile * N2 + rank * TC_M));
                }
                __syncwarp();
                if (++st == S) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the leader's warp walks the (uniform) loop, one elected lane drives the pair =====
        if (leader) {
            constexpr uint32_t idesc = make_idesc_f16(256, N2);
            uint32_t st = 0, ph = 0;
            uint32_t acc = 0, acc_ph = 0;
            const uint64_t desc_hi = make_sw128_desc(0);
            const uint32_t a_lo = smem_u32(a_base) >> 4, b_lo = smem_u32(b_base) >> 4;
            for (uint64_t it = it0; it <= my_tiles; ++it) {
                mbar_wait(smem_u32(tempty_bar + acc), acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * N2;
                for (uint32_t g = 0; g < n_grp; ++g) {
                    mbar_wait(smem_u32(full_bar + st), ph);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        for (uint32_t j = 0; j < KBS; ++j) {
                            const uint64_t da = desc_hi | (uint64_t)(a_lo + (g * KBS + j) * (TC_M * 128 / 16));
                            const uint64_t db = desc_hi | (uint64_t)(b_lo + (st * stage_bytes + j * TC_STAGE_BYTES) / 16);
#pragma unroll
                            for (uint32_t k = 0; k < TC_KB / 16; ++k)
                                umma_f16_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (g | j | k) ? 1u : 0u);
                        }
                        umma_commit_2cta(smem_u32(empty_bar + st), 3);     // frees the stage in BOTH CTAs
                        if (g + 1 == n_grp) umma_commit_2cta(smem_u32(tfull_bar + acc), 3);   // accumulators ready in BOTH CTAs
                    }
                    __syncwarp();
                    if (++st == S) { st = 0; ph ^= 1; }
                }
                if (++acc == ACC2) { acc = 0; acc_ph ^= 1; }
            }
        }
    } else if (warp >= 4) {
        tc_epilogue<N2, DUMP>(a, s_ctl, tmem_base, tfull_bar, mapa_u32(smem_u32(tempty_bar), 0), my_tiles,
                              /*tile0=*/cluster_id, /*tile_stride=*/n_clusters, /*q_base=*/a.q_tile * 256 + rank * TC_M,
                              /*nq_local=*/a.nq > rank * TC_M ? (a.nq - rank * TC_M < TC_M ? a.nq - rank * TC_M : TC_M) : 0,
                              /*unit_idx=*/cluster_id, /*n_lists=*/2 * n_clusters, ACC2);
    }

    tc_fence_before();
    cluster_sync_all();          // nobody may still signal or read the peer's barriers / TMEM
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, 512);
    }
}

// ---- K2: exact re-score + certification ---------------------------------------------------------
// exact -dot in the reference's order for one (query,row): one thread, sequential chain
template <int ORDER, typename QP>
SG_DEVICE_INLINE float exact_dot(QP q, const __half *__restrict__ row, uint32_t dim)
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
finalize_tc_kernel(const __grid_constant__ FinalizeArgs f)
{
    __shared__ uint64_t keys[TC_FIN_NC];        // surviving approximate keys, sorted
    __shared__ uint64_t ekeys[TC_KK_MAX];       // exact keys of the kk best
    __shared__ uint32_t s_flag, s_n;
    extern __shared__ __align__(16) uint8_t fin_dyn[];
    float *q_s = reinterpret_cast<float *>(fin_dyn);                     // [dim]
    uint32_t *rows_s = reinterpret_cast<uint32_t *>(q_s + f.dim);        // [rc][dim/2 + 1]
    const uint32_t q = blockIdx.x;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t L = f.n_lists;
    if (threadIdx.x == 0) {
        s_flag = 0;
        s_n = 0;
    }
    pdl_wait();                                  // launched early (PDL): the sweep's candidates are visible from here on
    __syncthreads();
    // (1) gather every candidate at or above the final shared threshold (a valid bound on the kk-th best, so at
    //     least kk survive whenever a threshold was ever published).  The region counts are read first (together
    //     with the query: one global round trip) and prefix-summed, then ALL threads walk the flat entry index:
    //     two dependent round trips in total, not two per region.
    const uint32_t gt = f.gtau[q];
    uint32_t *pre = reinterpret_cast<uint32_t *>(ekeys);          // [L + 1] prefix of min(count, cap_l); ekeys is free until (2)
    for (uint32_t l = threadIdx.x; l < L; l += blockDim.x) {
        uint32_t c = f.cnts[(size_t)q * L + l];
        if (c > f.cap_l) {                       // the region overflowed: candidates were lost
            s_flag = 1u;
            c = f.cap_l;
        }
        pre[l + 1] = c;
    }
    // the query, and with it the two norms of the certificate: err = |q - fp16(q)|_2 (what the tensor cores did not
    // see) and |q|_2, both in f64 and rounded up a little
    __shared__ double s_red[2][8];
    {
        double er = 0.0, n2 = 0.0;
        for (uint32_t i = threadIdx.x; i < f.dim; i += blockDim.x) {
            const float v = f.queries[(size_t)q * f.dim + i];
            q_s[i] = v;
            const double d = (double)v - (double)__half2float(__float2half_rn(v));
            er += d * d;
            n2 += (double)v * (double)v;
        }
        for (int o = 16; o > 0; o >>= 1) {
            er += __shfl_down_sync(0xFFFFFFFFu, er, o);
            n2 += __shfl_down_sync(0xFFFFFFFFu, n2, o);
        }
        if (lane == 0) {
            s_red[0][warp] = er;
            s_red[1][warp] = n2;
        }
    }
    if (threadIdx.x == 0) pre[0] = 0;
    __syncthreads();
    if (warp == 0) {                             // inclusive scan of pre[1..L] by one warp (L <= 296)
        uint32_t carry = 0;
        for (uint32_t base = 1; base <= L; base += 32) {
            const uint32_t i = base + lane;
            uint32_t v = i <= L ? pre[i] : 0;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
                if ((int)lane >= o) v += t;
            }
            if (i <= L) pre[i] = v + carry;
            carry += __shfl_sync(0xFFFFFFFFu, v, 31);
        }
    }
    __syncthreads();
    const uint32_t total = pre[L];
    // four entries per thread and step: the four candidate loads are in flight together (one global round trip per step
    // instead of four -- with k = 120 a query gathers a few thousand entries)
    for (uint32_t i0 = threadIdx.x; i0 < total; i0 += 4 * blockDim.x) {
        uint64_t key[4];
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) {
            const uint32_t i = i0 + u * blockDim.x;
            key[u] = KEY_SENTINEL;
            if (i < total) {
                uint32_t lo = 0, hi = L;         // region l with pre[l] <= i < pre[l + 1]
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (pre[mid] <= i) lo = mid; else hi = mid;
                }
                key[u] = __ldg(f.cand + ((size_t)q * L + lo) * f.cap_l + (i - pre[lo]));
            }
        }
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) {
            if (i0 + u * blockDim.x < total && (uint32_t)(key[u] >> 32) <= gt) {
                const uint32_t slot = atomicAdd(&s_n, 1u);
                if (slot < TC_FIN_NC) keys[slot] = key[u];
            }
        }
    }
    __syncthreads();
    uint32_t n = s_n;
    if (n > TC_FIN_NC) {
        n = TC_FIN_NC;
        if (threadIdx.x == 0) s_flag = 1u;
    }
    // sort the survivors ascending by approximate key (best first).  The usual case is a few dozen keys: a rank
    // sort (every key counts the keys below it -- all distinct, ids are -- and goes to that slot of the upper half
    // of the buffer) costs ONE barrier; the bitonic network with its log^2 barriers is kept for big sets.
    uint64_t *sorted = keys;
    if (n <= 512) {
        sorted = keys + TC_FIN_NC / 2;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const uint64_t mine = keys[i];
            uint32_t r = 0;
            for (uint32_t j2 = 0; j2 < n; ++j2) r += keys[j2] < mine;
            sorted[r] = mine;
        }
        __syncthreads();
    } else {
        const uint32_t M = next_pow2(n);
        for (uint32_t i = n + threadIdx.x; i < M; i += blockDim.x) keys[i] = KEY_SENTINEL;
        __syncthreads();
        block_bitonic_sort(keys, M);
    }
    const uint32_t kk = n < f.kk ? n : f.kk;     // the global top-kk by approximate score
    // weakest approximate similarity that was kept == bound on everything that was dropped
    const float weakest_approx_sim = kk ? -key32_to_dist((uint32_t)(sorted[kk - 1] >> 32)) : 0.0f;
    // (2) exact re-score, rc candidates per round: rows staged in shared memory with 16-byte loads (all in
    //     flight at once), then one thread per candidate walks the reference's sequential chain.  Row pitch
    //     dim/2 + 1 words: conflict-free.
    const uint32_t pitch = f.dim / 2 + 1, vec_per_row = f.dim / 8;
    for (uint32_t c0 = 0; c0 < kk; c0 += f.rc) {
        const uint32_t nc = kk - c0 < f.rc ? kk - c0 : f.rc;
        // six 16-byte pieces per thread and step, all loads issued before the first store (the rows are random 768-byte
        // reads from HBM: a piece-by-piece loop paid one round trip per piece)
        const uint32_t pieces = nc * vec_per_row;
        for (uint32_t i0 = threadIdx.x; i0 < pieces; i0 += 6 * blockDim.x) {
            uint4 x[6];
#pragma unroll
            for (uint32_t u = 0; u < 6; ++u) {
                const uint32_t i = i0 + u * blockDim.x;
                if (i < pieces) {
                    const uint32_t c = i / vec_per_row, v = i - c * vec_per_row;
                    const uint32_t row = (uint32_t)(sorted[c0 + c] & 0xFFFFFFFFu);
                    x[u] = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const __half *>(f.rows) + (size_t)row * f.dim) + v);
                }
            }
#pragma unroll
            for (uint32_t u = 0; u < 6; ++u) {
                const uint32_t i = i0 + u * blockDim.x;
                if (i < pieces) {
                    const uint32_t c = i / vec_per_row, v = i - c * vec_per_row;
                    uint32_t *dst = rows_s + (size_t)c * pitch + 4 * v;
                    dst[0] = x[u].x; dst[1] = x[u].y; dst[2] = x[u].z; dst[3] = x[u].w;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < nc) {
            const uint32_t row = (uint32_t)(sorted[c0 + threadIdx.x] & 0xFFFFFFFFu);
            const __half *rp = reinterpret_cast<const __half *>(rows_s + (size_t)threadIdx.x * pitch);
            const float dot = f.order == SG_ORDER_AVX2 ? exact_dot<SG_ORDER_AVX2>(q_s, rp, f.dim)
                                                       : exact_dot<SG_ORDER_SCALAR>(q_s, rp, f.dim);
            ekeys[c0 + threadIdx.x] = make_key(order_key32(neg_dot_to_dist(dot)), row);
        }
        __syncthreads();
    }
    // exact keys -> ascending (dist, id): rank sort into the lower half of `keys` (the approximate list is done with)
    if (threadIdx.x < kk) {
        const uint64_t mine = ekeys[threadIdx.x];
        uint32_t r = 0;
        for (uint32_t j2 = 0; j2 < kk; ++j2) r += ekeys[j2] < mine;
        keys[r] = mine;
    }
    __syncthreads();
    const uint64_t *fin = keys;
    const uint32_t out_n = kk < f.k ? kk : f.k;
    // (3) certification.  Rows that were not re-scored have approx sim <= weakest_approx_sim (survivors past
    //     the kk-th by the sort; everything else was below a threshold that never exceeded the kk-th best),
    //     hence exact sim <= weakest_approx_sim + E; none can enter (or tie into) the top k if kth_exact_sim
    //     is larger.  Fewer than kk candidates with more live rows around (NaN scores, no threshold) -> flag.
    if (threadIdx.x == 0 && f.n_live > kk) {
        bool ok = false;
        if (kk == f.kk && out_n == f.k) {
            const float kth_exact_sim = -key32_to_dist((uint32_t)(fin[f.k - 1] >> 32));
            const float R = *f.max_row_norm;
            double er = 0.0, n2 = 0.0;
            for (int wv = 0; wv < 8; ++wv) {
                er += s_red[0][wv];
                n2 += s_red[1][wv];
            }
            const float err_q = (float)(sqrt(er) * 1.000001), qnorm = (float)(sqrt(n2) * 1.000001);
            const float E = err_q * R + f.c_acc * qnorm * R;
            ok = kth_exact_sim > weakest_approx_sim + E;     // NaN compares false -> flagged
        }
        if (!ok) s_flag = 1u;
    }
    __syncthreads();
    const bool flagged = s_flag != 0;
    for (uint32_t i = threadIdx.x; i < f.k; i += blockDim.x) {
        const bool valid = i < out_n;
        const uint64_t key = valid ? fin[i] : KEY_SENTINEL;
        f.ids_out[(size_t)q * f.k + i] = valid ? (uint32_t)(key & 0xFFFFFFFFu) + f.id_base : 0xFFFFFFFFu;
        f.dist_out[(size_t)q * f.k + i] = valid ? key32_to_dist((uint32_t)(key >> 32)) : __uint_as_float(0x7F800000u);
    }
    if (threadIdx.x == 0) {
        f.counts_out[q] = out_n;
        f.flags[q] = flagged;
        if (flagged) {
            const uint32_t slot = atomicAdd(f.n_flagged, 1u);   // arrival order is irrelevant: slots are independent
            f.flag_list[slot] = q;
            atomicAdd(f.flag_total, 1ull);
        }
    }
    // multi-GPU: a certified result is final -- push it to the peers now (flagged ones are pushed by the re-run's merge)
    if (!flagged) push_query(f.push, q, f.ids_out + (size_t)q * f.k, f.dist_out + (size_t)q * f.k);
}

// ---- diagnostics: exact score of every (query, row) pair ------------------------------------------------
__global__ void __launch_bounds__(256)
exact_scores_kernel(const float *__restrict__ q, uint32_t b, const __half *__restrict__ rows, uint32_t dim, uint64_t row0,
                    uint64_t n, int order, float *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const __half *rp = rows + (row0 + i) * dim;
    for (uint32_t qi = 0; qi < b; ++qi) {
        const float *qp = q + (size_t)qi * dim;
        const float dot = order == SG_ORDER_AVX2 ? exact_dot<SG_ORDER_AVX2>(qp, rp, dim) : exact_dot<SG_ORDER_SCALAR>(qp, rp, dim);
        out[(size_t)qi * n + i] = dot;
    }
}

// The sweep's sample phase ends in a device-wide meeting point, so the whole grid must be co-resident:
// launch cooperatively (the runtime then guarantees it); SG_TC_COOP=0 or a refusal falls back to a plain
// launch (the kernel's own time-out keeps that safe, at worst slow).
template <typename K>
cudaError_t launch_sweep(K kern, const CUtensorMap &tmap_rows, const TcArgs &a, uint32_t grid,
                         size_t smem, cudaStream_t st)
{
    const char *v = getenv("SG_TC_COOP");
    const int coop = (v && *v) ? atoi(v) : 1;
    // cooperative (co-residency guaranteed) + programmatic stream serialisation (the prologue -- barrier init, TMEM
    // allocation, descriptor prefetch -- overlaps the tail of the previous step; the kernel calls pdl_wait() before it
    // touches global memory).  A driver that refuses the combination gets the cooperative launch alone.
    static std::atomic<int> both_ok{1};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    if (coop) {
        if (both_ok.load(std::memory_order_relaxed)) {
            cfg.attrs = attr;
            cfg.numAttrs = 2;
            cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmap_rows, a);
            if (e == cudaSuccess) return e;
            cudaGetLastError();
            both_ok.store(0, std::memory_order_relaxed);
        }
        cfg.attrs = attr + 1;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmap_rows, a);
        if (e == cudaSuccess) return e;
        cudaGetLastError();                      // clear; fall through to the plain launch
    }
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, tmap_rows, a);
}

}  // namespace

size_t tc_smem_bytes(uint32_t dim, uint32_t stages, uint32_t kbs)
{
    size_t s = (size_t)(dim / TC_KB) * TC_M * 128;          // A
    s += (size_t)stages * kbs * TC_STAGE_BYTES;              // B ring: stages of kbs K-blocks
    s += (size_t)(2 * stages + 2 * TC_ACC_STAGES + 1) * 8 + 16;
    s += 16 + 128 * 16;                                      // control words + the epilogue groups' exchange buffer
    return s;
}

// Accumulation-error coefficient of the certificate: |tensor-core dot - reference-order dot| <= c_acc * |q| * |row|
// for the same fp16 operands.  Tensor core: products exact in f32, <= dim accumulation steps of <= 1 ulp each
// (truncation allowed): dim * 2^-23 * sum|a_i b_i|.  Reference chain: dim/4 + 3 round-to-nearest additions of at
// most sum|a_i b_i| plus one rounding per product: (dim/4 + 4) * 2^-24.  sum|a_i b_i| <= |fp16(q)| |row| <=
// (1 + 2^-11) |q| |row|; the factor 1.25 covers that and the rounding of the bound itself.
// tests/test_scan_tc_gpu.py::test_accumulation_error_bound measures the real maximum against it.
float tc_c_acc(uint32_t dim)
{
    return 1.25f * ((float)dim * 1.1920929e-7f + (float)(dim / 4 + 4) * 5.9604645e-8f);
}

cudaError_t launch_scan_tc(const CUtensorMap &tmap_rows, const TcArgs &a, uint32_t grid, size_t smem, cudaStream_t st)
{
    static DevAttr da, da_dump;
    cudaError_t e;
    if (a.dump) {
        if ((e = ensure_smem(scan_tc_kernel<true>, da_dump, smem)) != cudaSuccess) return e;
        scan_tc_kernel<true><<<grid, TC_THREADS, smem, st>>>(tmap_rows, a);
        e = cudaGetLastError();
    } else {
        if ((e = ensure_smem(scan_tc_kernel<false>, da, smem)) != cudaSuccess) return e;
        e = launch_sweep(scan_tc_kernel<false>, tmap_rows, a, grid, smem, st);
    }
    count_launch();
    return e;
}

cudaError_t launch_scan_tc2(const CUtensorMap &tmap_rows, const TcArgs &a, uint32_t grid, size_t smem, cudaStream_t st)
{
    static DevAttr da, da_dump;
    cudaError_t e;
    if (a.dump) {
        if ((e = ensure_smem(scan_tc2_kernel<true>, da_dump, smem)) != cudaSuccess) return e;
        scan_tc2_kernel<true><<<grid, TC_THREADS, smem, st>>>(tmap_rows, a);   // __cluster_dims__(2,1,1): grid must be even
        e = cudaGetLastError();
    } else {
        if ((e = ensure_smem(scan_tc2_kernel<false>, da, smem)) != cudaSuccess) return e;
        e = launch_sweep(scan_tc2_kernel<false>, tmap_rows, a, grid, smem, st);
    }
    count_launch();
    return e;
}

cudaError_t launch_finalize_tc(const FinalizeArgs &f, uint32_t b, cudaStream_t st)
{
    const size_t smem = (size_t)f.dim * 4 + (size_t)f.rc * (f.dim / 2 + 1) * 4;
    static DevAttr da;
    if (smem > 48 * 1024) {
        cudaError_t e = ensure_smem(finalize_tc_kernel, da, smem);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = launch_pdl(finalize_tc_kernel, dim3(b), dim3(256), smem, st, f);
    count_launch();
    return e;
}

cudaError_t launch_exact_scores(const float *d_q, uint32_t b, const void *rows, uint32_t dim, uint64_t row0, uint64_t n,
                                int order, float *d_out, cudaStream_t st)
{
    if (n == 0 || b == 0) return cudaSuccess;
    exact_scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_q, b, reinterpret_cast<const __half *>(rows), dim, row0,
                                                                     n, order, d_out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace sg
