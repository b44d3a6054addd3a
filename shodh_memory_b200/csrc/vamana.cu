// vamana.cu -- the Vamana graph index on the device.
//
// Reference: src/vector_db/vamana.rs -- greedy_search :576-657 (beam search from the medoid, the
// default arm of VamanaIndex::search :779-807), robust_prune :665-746, find_medoid :400-441,
// build :200-284, add_vector :853-973.
//
// ann_search_kernel: ONE WARP PER QUERY, and the result is the reference's, not "a" beam search:
// SearchCandidate is ordered by (distance total_cmp, id) (:1663-1673), a total order, so which element
// the two BinaryHeaps of greedy_search pop is uniquely determined and any priority queue reproduces
// it.  The warp keeps ONE sorted array P in shared memory that is both heaps at once:
//   * w (the bounded max-heap of the k best) is the first min(k, |P|) entries of P;
//   * candidates (the min-heap) is the set of not-yet-expanded entries of P.  An entry that falls out
//     of w can only ever end the search when it is popped (`current.distance > w.peek().distance`,
//     :597-603), so it is dropped at once -- unless its distance TIES the k-th (float ==), in which
//     case the reference would still expand it and so do we (up to 32 such entries are kept).
// Per step the warp pops the first unexpanded entry, reads its <= R neighbour ids (one per lane),
// filters them through the query's visited set (open-addressing hash table in global memory, L2
// resident, atomicCAS insert == `visited.insert`), scores the new ones -- one row per lane, straight
// from HBM in 16-byte pieces, in the reference's summation order (sg_dist.cuh), so distances are
// bit-identical to the CPU path -- and inserts them in neighbour order with the reference's
// `w.len() < k || dist < w.peek().distance` test evaluated against the list as it evolves.
// Random 768-byte row reads: latency-bound per query, so throughput comes from many queries in
// flight (4 warps per CTA, all SMs).  Algorithmic bytes = distance evaluations x row bytes.
#include <cstdlib>

#include "sg_internal.h"
#include "sg_dist.cuh"

namespace sg {

namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t ANN_SLACK = 32;   // entries kept beyond the beam: candidates that tie the k-th distance

SG_DEVICE_INLINE uint32_t hash32(uint32_t x)
{
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

// `visited.insert(id)` for up to 32 ids at once, one per lane (want = this lane has an id to insert): true when
// the id was not in the set yet.  The table belongs to ONE warp, so no atomics are needed: a lane reads a whole
// 4-slot bucket with one 16-byte L2 load, and lanes that pick the same empty slot in the same round are
// separated through a 1 KB arbiter in shared memory (each writes its lane number at arb[slot % 1024]; who reads
// its own number back takes the slot, the others look again and see it filled -- a list that repeated an id would
// be caught the same way: same id, same slot).  match_any would do this too, but costs ~1000 cycles on 32
// distinct values.  Buckets are probed linearly.  All 32 lanes must call.
SG_DEVICE_INLINE bool vis_insert_warp(uint32_t *tab, uint32_t mask, uint32_t id, bool want, uint32_t lane, uint8_t *arb)
{
    uint32_t h = hash32(id) & mask & ~3u;
    bool pending = want, isnew = false;
    for (uint32_t probes = 0; __any_sync(FULL, pending); ++probes) {
        uint32_t tgt = NONE;
        if (pending) {
            const uint4 b = __ldcg(reinterpret_cast<const uint4 *>(tab + h));
            if (b.x == id || b.y == id || b.z == id || b.w == id) {
                pending = false;
            } else {
                const int e = b.x == NONE ? 0 : (b.y == NONE ? 1 : (b.z == NONE ? 2 : (b.w == NONE ? 3 : -1)));
                if (e >= 0) {
                    tgt = h + (uint32_t)e;
                    arb[tgt & 1023u] = (uint8_t)lane;
                } else if (probes > mask) {
                    pending = false;                             // table full (cannot happen below half load)
                } else {
                    h = (h + 4) & mask;
                }
            }
        }
        __syncwarp();
        if (tgt != NONE && arb[tgt & 1023u] == (uint8_t)lane) {
            __stcg(tab + tgt, id);
            isnew = true;
            pending = false;
        }
        __syncwarp();                                            // the round's stores are visible to the next round's loads
    }
    return isnew;
}

SG_DEVICE_INLINE size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// store row -> f32 vector in shared memory (fp16 -> f32 is exact), one warp
template <bool F16>
SG_DEVICE_INLINE void load_row_f32(const void *rows, uint32_t dim, uint64_t row, float *dst, uint32_t lane)
{
    if constexpr (F16) {
        const __half *p = static_cast<const __half *>(rows) + row * dim;
        for (uint32_t d = lane; d < dim; d += 32) dst[d] = __half2float(p[d]);
    } else {
        const float *p = static_cast<const float *>(rows) + row * dim;
        for (uint32_t d = lane; d < dim; d += 32) dst[d] = p[d];
    }
}

// one whole row global -> shared by the TMA engine (cp.async.bulk, SASS UBLKCP), completion on an mbarrier
SG_DEVICE_INLINE void bulk_row_load(void *dst_smem, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
SG_DEVICE_INLINE void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// STAGE: the <= 32 rows of a step are first copied into shared memory by the TMA engine -- each lane issues ONE
// bulk copy for its row, all rows in flight at once behind one mbarrier, one HBM round trip per step -- and
// each lane then walks ITS row out of shared memory (rows padded by 16 bytes: conflict-free).  Without STAGE (rows too long for shared memory) each lane
// reads its row straight from global memory, eight 16-byte loads at a time.
template <bool F16, int ORDER, int METRIC, bool STAGE>
__global__ void __launch_bounds__(128)
ann_search_kernel(const AnnArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const uint32_t q = blockIdx.x * W + warp;
    if (q >= a.nq) return;                       // warps are independent: no block-wide barrier anywhere
    const uint32_t dim = a.dim, K = a.K, CAP = K + ANN_SLACK;
    uint8_t *base = smem + (size_t)warp * (ann_warp_smem_bytes(dim, K) + (STAGE ? ann_stage_smem_bytes(dim, F16) : 0));
    float *q_s = reinterpret_cast<float *>(base);
    const size_t list_bytes = align16((size_t)CAP * 8) + align16((size_t)CAP);
    uint64_t *P = reinterpret_cast<uint64_t *>(base + align16((size_t)dim * 4));
    uint8_t *ex = reinterpret_cast<uint8_t *>(P) + align16((size_t)CAP * 8);
    uint64_t *P2 = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(P) + list_bytes);   // the merge's output list
    uint8_t *ex2 = reinterpret_cast<uint8_t *>(P2) + align16((size_t)CAP * 8);
    uint64_t *Cs = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(P2) + list_bytes);  // [32] candidate keys
    uint8_t *arb = reinterpret_cast<uint8_t *>(Cs) + 256;                                       // [1024] slot arbiter
    uint8_t *rowbuf = base + ann_warp_smem_bytes(dim, K);          // STAGE: [32][row_bytes + 16]
    const uint32_t row_bytes = dim * (F16 ? 2u : 4u), pitch = row_bytes + 16;
    const uint32_t bar = smem_u32(rowbuf + (size_t)32 * pitch);   // STAGE: the warp's mbarrier
    uint32_t bar_phase = 0;
    if constexpr (STAGE) {
        if (lane == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
    }

    if (a.query_rows) {
        load_row_f32<F16>(a.rows, dim, a.query_rows[q], q_s, lane);
    } else {
        for (uint32_t d = lane; d < dim; d += 32) q_s[d] = a.queries[(size_t)q * dim + d];
    }
    __syncwarp();
    float qn2 = 0.0f;
    if constexpr (METRIC == SG_METRIC_COSINE) qn2 = norm2_smem<ORDER>(q_s, dim);

    uint32_t *vis = a.vis + (size_t)q * a.H;
    const uint32_t hmask = a.H - 1;
    uint32_t vis_count = 1, flag = 0;
    unsigned long long hops = 0, evals = 1;

    // entry point (:586-595)
    {
        const float d0 = row_distance<F16, ORDER, METRIC>(a.rows, dim, a.entry, q_s, qn2);
        if (lane == 0) {
            P[0] = make_key(order_key32(d0), a.entry);
            ex[0] = 0;
            __stcg(vis + (hash32(a.entry) & hmask & ~3u), a.entry);
        }
    }
    uint32_t pn = 1;
    __syncwarp();
    // profiling hook: cycles per phase (0 pop, 1 list read + visited set, 2 row fetch, 3 scoring, 4 insertion)
    const bool prof = a.phase != nullptr;
    long long ph[5] = {0, 0, 0, 0, 0}, tprev = prof ? clock64() : 0;
    auto mark = [&](int i) {
        if (prof) {
            const long long t = clock64();
            ph[i] += t - tprev;
            tprev = t;
        }
    };

    for (;;) {
        // candidates.pop(): the first entry of P that has not been expanded
        uint32_t cur = NONE;
        for (uint32_t b0 = 0; b0 < pn; b0 += 32) {
            const uint32_t i = b0 + lane;
            const uint32_t m = __ballot_sync(FULL, i < pn && !ex[i]);
            if (m) {
                cur = b0 + __ffs(m) - 1;
                break;
            }
        }
        if (cur == NONE) break;
        const uint64_t ckey = P[cur];
        const uint32_t wl = pn < K ? pn : K;
        if (key32_to_dist((uint32_t)(ckey >> 32)) > key32_to_dist((uint32_t)(P[wl - 1] >> 32))) break;   // :597-603
        __syncwarp();
        if (lane == 0) ex[cur] = 1;
        __syncwarp();
        const uint32_t cid = (uint32_t)ckey;
        if (cid >= a.graph_len) continue;                                                                // :606-609
        ++hops;
        mark(0);
        for (uint32_t c0 = 0; c0 < a.gs; c0 += 32) {
            uint32_t nb = NONE;
            if (c0 + lane < a.gs) nb = a.graph[(size_t)cid * a.gs + c0 + lane];
            const bool valid = nb != NONE && nb < a.n_rows;
            const bool isnew = vis_insert_warp(vis, hmask, nb, valid && !(flag & 1u), lane, arb);
            const uint32_t newmask = __ballot_sync(FULL, isnew);
            mark(1);
            if (!newmask) continue;
            vis_count += __popc(newmask);
            evals += __popc(newmask);
            if (vis_count > (a.H >> 1)) flag |= 1u;                       // visited table half full: stop admitting ids
            float d = 0.0f;
            if constexpr (STAGE) {
                if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)__popc(newmask) * row_bytes);
                __syncwarp();
                if (isnew)
                    bulk_row_load(rowbuf + (size_t)lane * pitch, static_cast<const uint8_t *>(a.rows) + (size_t)nb * row_bytes,
                                  row_bytes, bar);
                mbar_wait(bar, bar_phase);
                bar_phase ^= 1;
                mark(2);
                if (isnew) d = row_distance_ptr<F16, ORDER, METRIC, false>(rowbuf + (size_t)lane * pitch, dim, q_s, qn2);
                __syncwarp();                                             // rowbuf is reused by the next chunk / step
            } else {
                if (isnew) d = row_distance<F16, ORDER, METRIC>(a.rows, dim, nb, q_s, qn2);
            }
            mark(3);
            {
                // the node most likely to be expanded next is the better of the first unexpanded entry of P and
                // the best new row: start pulling its neighbour list into L2 while the insertion runs
                uint64_t best = isnew ? make_key(order_key32(d), nb) : KEY_SENTINEL;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const uint64_t other = __shfl_xor_sync(FULL, best, o);
                    best = other < best ? other : best;
                }
                if (lane == 0 && best != KEY_SENTINEL && (uint32_t)best < a.graph_len)
                    prefetch_l2(a.graph + (size_t)(uint32_t)best * a.gs);
            }
            // a lane whose distance already fails `dist < w.peek()` on a full w can never pass later
            // (the k-th distance only falls): drop it before the sequential part
            bool cand = isnew;
            if (pn >= K && cand) cand = d < key32_to_dist((uint32_t)(P[K - 1] >> 32));
            uint32_t todo = __ballot_sync(FULL, cand);
            // Batched merge.  With no float-equal distances in play the sequential rule below is order-independent:
            // the list after all of this step's candidates is exactly the first K of sorted(P + candidates) (a
            // candidate inside the union's top K is strictly closer than the K-th it is tested against; one outside
            // is refused or pushed out and trimmed).  So every entry's new position is computed in parallel --
            // old index + number of smaller candidates; candidate: lower_bound in P + rank among the candidates --
            // and scattered into the second list.  Equal distances (next to a new entry, or across the K-th
            // boundary), NaNs, or a tail of ties already present make the outcome order-dependent: then the new
            // list is discarded and the candidates go through the sequential path, the reference's rule verbatim.
            {
                const uint32_t m = __popc(todo);
                bool fast = m > 2 && pn <= K;
                if (fast && __ballot_sync(FULL, cand && d != d)) fast = false;
                if (fast) {
                    const float last = key32_to_dist((uint32_t)(P[pn - 1] >> 32));
                    if (last != last) fast = false;
                }
                if (fast) {
                    const uint32_t total = pn + m, keepN = total < K + 1 ? total : K + 1;
                    const uint64_t kx = make_key(order_key32(d), nb);
                    if (cand) Cs[__popc(todo & ((1u << lane) - 1))] = kx;
                    __syncwarp();
                    if (cand) {
                        uint32_t np = 0;
                        for (uint32_t j = 0; j < m; ++j) np += Cs[j] < kx;
                        uint32_t lo = 0, hi = pn;
                        while (lo < hi) {
                            const uint32_t mid = (lo + hi) >> 1;
                            if (P[mid] < kx) lo = mid + 1;
                            else hi = mid;
                        }
                        np += lo;
                        if (np < keepN) {
                            P2[np] = kx;
                            ex2[np] = 2;                                  // 2 = entered in this step
                        }
                    }
                    for (uint32_t i = lane; i < pn; i += 32) {
                        const uint64_t key = P[i];
                        uint32_t np = i;
                        for (uint32_t j = 0; j < m; ++j) np += Cs[j] < key;
                        if (np < keepN) {
                            P2[np] = key;
                            ex2[np] = ex[i];
                        }
                    }
                    __syncwarp();
                    bool tie = false;
                    for (uint32_t i = lane; i + 1 < keepN; i += 32) {
                        const float da = key32_to_dist((uint32_t)(P2[i] >> 32)), db = key32_to_dist((uint32_t)(P2[i + 1] >> 32));
                        if (da == db && (ex2[i] == 2 || ex2[i + 1] == 2 || i == K - 1)) tie = true;
                    }
                    if (__ballot_sync(FULL, tie)) {
                        fast = false;                                     // P is untouched: redo in order
                    } else {
                        const uint32_t newpn = total < K ? total : K;
                        for (uint32_t i = lane; i < newpn; i += 32)
                            if (ex2[i] == 2) ex2[i] = 0;
                        uint64_t *tp = P; P = P2; P2 = tp;
                        uint8_t *te = ex; ex = ex2; ex2 = te;
                        pn = newpn;
                        todo = 0;
                    }
                    __syncwarp();
                }
            }
            while (todo) {                                                // neighbour order == lane order (:612)
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const float dx = __shfl_sync(FULL, d, src);
                const uint32_t idx = __shfl_sync(FULL, nb, src);
                if (pn >= K && !(dx < key32_to_dist((uint32_t)(P[K - 1] >> 32)))) continue;              // :625
                const uint64_t kx = make_key(order_key32(dx), idx);
                uint32_t pos = 0;
                for (uint32_t b0 = 0; b0 < pn; b0 += 32) {
                    const uint32_t i = b0 + lane;
                    pos += __popc(__ballot_sync(FULL, i < pn && P[i] < kx));
                }
                if (pn > pos) {                                           // shift [pos, pn) up by one, top chunk first
                    const int nch = (int)((pn - pos + 31) >> 5);
                    for (int ch = nch - 1; ch >= 0; --ch) {
                        const uint32_t i = pos + (uint32_t)ch * 32 + lane;
                        uint64_t kv = 0;
                        uint8_t ev = 0;
                        const bool act = i < pn;
                        if (act) {
                            kv = P[i];
                            ev = ex[i];
                        }
                        __syncwarp();
                        if (act) {
                            P[i + 1] = kv;
                            ex[i + 1] = ev;
                        }
                        __syncwarp();
                    }
                }
                if (lane == 0) {
                    P[pos] = kx;
                    ex[pos] = 0;
                }
                ++pn;
                __syncwarp();
                if (pn > K) {
                    // w.pop(): what is beyond the k-th stays only while it ties the k-th distance
                    const float wd = key32_to_dist((uint32_t)(P[K - 1] >> 32));
                    const uint32_t i = K + lane;
                    const uint32_t bad = __ballot_sync(FULL, i < pn && key32_to_dist((uint32_t)(P[i] >> 32)) > wd);
                    if (bad) pn = K + __ffs(bad) - 1;
                    if (pn == CAP) {
                        flag |= 2u;                                       // > 32 ties at the boundary: one is dropped
                        pn = CAP - 1;
                    }
                }
            }
            mark(4);
        }
    }
    if (prof && lane == 0)
        for (int i = 0; i < 5; ++i) atomicAdd(a.phase + i, (unsigned long long)ph[i]);

    // w, ascending; VamanaIndex::search drops soft-deleted ids and takes k (:798-804)
    const uint32_t wl = pn < K ? pn : K;
    uint32_t outn = 0;
    for (uint32_t b0 = 0; b0 < wl && outn < a.k_out; b0 += 32) {
        const uint32_t i = b0 + lane;
        uint64_t kv = 0;
        bool keep = false;
        if (i < wl) {
            kv = P[i];
            const uint32_t id = (uint32_t)kv;
            keep = !(a.deleted && ((a.deleted[id >> 5] >> (id & 31)) & 1u));
        }
        const uint32_t m = __ballot_sync(FULL, keep);
        const uint32_t r = outn + __popc(m & ((1u << lane) - 1));
        if (keep && r < a.k_out) {
            a.ids_out[(size_t)q * a.k_out + r] = (uint32_t)kv + a.id_base;
            a.dist_out[(size_t)q * a.k_out + r] = key32_to_dist((uint32_t)(kv >> 32));
        }
        outn += __popc(m);
    }
    if (outn > a.k_out) outn = a.k_out;
    for (uint32_t i = outn + lane; i < a.k_out; i += 32) {
        a.ids_out[(size_t)q * a.k_out + i] = NONE;
        a.dist_out[(size_t)q * a.k_out + i] = __uint_as_float(0x7F800000u);
    }
    if (lane == 0) {
        a.counts_out[q] = outn;
        if (a.flags) a.flags[q] = flag;
        if (a.stats) {
            a.stats[2 * (size_t)q] = hops;
            a.stats[2 * (size_t)q + 1] = evals;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// add_vector linking (vamana.rs:905-966): the new node takes the search result as its list; every
// neighbour gets the back-edge and, beyond max_degree, keeps its max_degree closest by (distance, id).
// One CTA per inserted node, one warp per neighbour.
// ---------------------------------------------------------------------------------------------------------
template <bool F16, int ORDER, int METRIC>
__global__ void __launch_bounds__(256)
graph_link_kernel(const void *rows, uint32_t dim, uint32_t *graph, uint32_t gs, uint32_t id, const uint32_t *res_ids,
                  const uint32_t *res_count)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const uint32_t r = *res_count < gs ? *res_count : gs;
    for (uint32_t t = threadIdx.x; t < gs; t += blockDim.x) graph[(size_t)id * gs + t] = t < r ? res_ids[t] : NONE;
    float *q_s = reinterpret_cast<float *>(smem + (size_t)warp * (align16((size_t)dim * 4) + 128 * 8));
    uint64_t *keys = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(q_s) + align16((size_t)dim * 4));
    for (uint32_t t = warp; t < r; t += W) {
        const uint32_t v = res_ids[t];
        uint32_t *row = graph + (size_t)v * gs;
        uint32_t deg = 0;
        for (uint32_t c0 = 0; c0 < gs; c0 += 32) {
            const uint32_t x = c0 + lane < gs ? row[c0 + lane] : NONE;
            deg += __popc(__ballot_sync(FULL, x != NONE));
        }
        if (deg < gs) {
            if (lane == 0) row[deg] = id;
            continue;
        }
        // gs + 1 candidates: distance(v, candidate) in the reference's order, sort, keep the gs closest
        load_row_f32<F16>(rows, dim, v, q_s, lane);
        __syncwarp();
        float qn2 = 0.0f;
        if constexpr (METRIC == SG_METRIC_COSINE) qn2 = norm2_smem<ORDER>(q_s, dim);
        const uint32_t nc = gs + 1;
        for (uint32_t c0 = 0; c0 < nc; c0 += 32) {
            const uint32_t i = c0 + lane;
            if (i < nc) {
                const uint32_t cidv = i < gs ? row[i] : id;
                const float d = row_distance<F16, ORDER, METRIC>(rows, dim, cidv, q_s, qn2);
                keys[i] = make_key(order_key32(d), cidv);
            }
        }
        __syncwarp();
        for (uint32_t c0 = 0; c0 < nc; c0 += 32) {                        // rank sort: ids are distinct, so are the keys
            const uint32_t i = c0 + lane;
            if (i < nc) {
                const uint64_t mine = keys[i];
                uint32_t rank = 0;
                for (uint32_t j = 0; j < nc; ++j) rank += keys[j] < mine;
                if (rank < gs) row[rank] = (uint32_t)mine;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// find_medoid, first half (vamana.rs:407-424): centroid[j] = (sum over rows, in row order, f32) / n.
// One thread owns one column so the additions happen in exactly the reference's order; the CTA's other
// warps only stage row tiles into shared memory.  grid = dim / 32 CTAs of 256 threads.
// ---------------------------------------------------------------------------------------------------------
template <bool F16>
__global__ void __launch_bounds__(256)
centroid_kernel(const void *rows, uint32_t dim, uint64_t n, float *centroid)
{
    __shared__ float tile[256][33];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t col = blockIdx.x * 32 + lane;
    float sum = 0.0f;
    for (uint64_t r0 = 0; r0 < n; r0 += 256) {
        for (uint32_t t = warp; t < 256; t += 8) {
            const uint64_t r = r0 + t;
            float v = 0.0f;
            if (r < n) {
                if constexpr (F16) v = __half2float(static_cast<const __half *>(rows)[r * dim + col]);
                else v = static_cast<const float *>(rows)[r * dim + col];
            }
            tile[t][lane] = v;
        }
        __syncthreads();
        if (warp == 0) {
            const uint32_t m = n - r0 < 256 ? (uint32_t)(n - r0) : 256;
            for (uint32_t t = 0; t < m; ++t) sum = __fadd_rn(sum, tile[t][lane]);
        }
        __syncthreads();
    }
    if (warp == 0) centroid[col] = __fdiv_rn(sum, (float)n);
}

// ---------------------------------------------------------------------------------------------------------
// Bulk build (VamanaIndex::build :200-284, batched).  The reference inserts nodes one after the other;
// here a batch of nodes searches the graph as it stood before the batch, is pruned, and its edges and
// back-edges are applied together -- the GPU formulation of the same construction.  The resulting
// graph is not the reference's edge for edge (its own start graph is random per run, :287-316); it is
// checked by recall against the exact scan, next to the oracle's build of the same data.
// ---------------------------------------------------------------------------------------------------------
__global__ void graph_init_kernel(uint32_t *graph, uint32_t gs, uint64_t n, uint64_t seed)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t *row = graph + i * gs;
    const uint32_t degree = (uint32_t)(n - 1 < gs ? n - 1 : gs);
    uint32_t m = 0;
    uint64_t ctr = 0;
    while (m < degree) {
        uint64_t x = seed ^ (i * 0x9E3779B97F4A7C15ull + (ctr++) * 0xD1B54A32D192ED03ull);
        x ^= x >> 31;
        x *= 0xBF58476D1CE4E5B9ull;
        x ^= x >> 29;
        x *= 0x94D049BB133111EBull;
        x ^= x >> 32;
        const uint64_t idx = x % (n - 1);
        const uint32_t j = (uint32_t)(idx >= i ? idx + 1 : idx);          // never the node itself (:301-303)
        bool dup = false;
        for (uint32_t t = 0; t < m; ++t) dup |= row[t] == j;
        if (!dup) row[m++] = j;
    }
    for (uint32_t t = degree; t < gs; ++t) row[t] = NONE;
}

// distance between two store rows with the work split over a warp (build only: deterministic, not the
// reference's summation order -- the build has no bit-level contract)
template <bool F16>
SG_DEVICE_INLINE float warp_pair_distance(const void *rows, uint32_t dim, uint32_t x, uint32_t y, int metric, uint32_t lane)
{
    float dot = 0.0f, nx = 0.0f, ny = 0.0f, l2 = 0.0f;
    if constexpr (F16) {
        const uint4 *px = reinterpret_cast<const uint4 *>(static_cast<const __half *>(rows) + (size_t)x * dim);
        const uint4 *py = reinterpret_cast<const uint4 *>(static_cast<const __half *>(rows) + (size_t)y * dim);
        for (uint32_t c = lane; c < (dim >> 3); c += 32) {
            uint4 a = __ldg(px + c), b = __ldg(py + c);
            const __half2 *ha = reinterpret_cast<const __half2 *>(&a), *hb = reinterpret_cast<const __half2 *>(&b);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 fa = __half22float2(ha[t]), fb = __half22float2(hb[t]);
                dot = fmaf(fa.x, fb.x, fmaf(fa.y, fb.y, dot));
                nx = fmaf(fa.x, fa.x, fmaf(fa.y, fa.y, nx));
                ny = fmaf(fb.x, fb.x, fmaf(fb.y, fb.y, ny));
                const float dx = fa.x - fb.x, dy = fa.y - fb.y;
                l2 = fmaf(dx, dx, fmaf(dy, dy, l2));
            }
        }
    } else {
        const float4 *px = reinterpret_cast<const float4 *>(static_cast<const float *>(rows) + (size_t)x * dim);
        const float4 *py = reinterpret_cast<const float4 *>(static_cast<const float *>(rows) + (size_t)y * dim);
        for (uint32_t c = lane; c < (dim >> 2); c += 32) {
            const float4 a = __ldg(px + c), b = __ldg(py + c);
            dot = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, dot))));
            nx = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, nx))));
            ny = fmaf(b.x, b.x, fmaf(b.y, b.y, fmaf(b.z, b.z, fmaf(b.w, b.w, ny))));
            const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
            l2 = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, l2))));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dot += __shfl_xor_sync(FULL, dot, o);
        nx += __shfl_xor_sync(FULL, nx, o);
        ny += __shfl_xor_sync(FULL, ny, o);
        l2 += __shfl_xor_sync(FULL, l2, o);
    }
    if (metric == SG_METRIC_EUCLIDEAN) return l2;
    if (metric == SG_METRIC_COSINE) {
        const float den = sqrtf(nx * ny);
        return 1.0f - (den < 1e-10f ? 0.0f : dot / den);
    }
    return -dot;
}

constexpr uint32_t PRUNE_MAX_CAND = 128;

// robust_prune (:665-746) for one node by one CTA of 128 threads.  cand[0..n_cand) (shared memory) are
// candidate ids, duplicates and the node itself allowed; on return out[0..gs) (global) holds the kept
// ids in selection order, NONE-padded.  keys/kept/kept_ne/flag are shared scratch.
// by_id = the back-edge form (:251-262): the reference hands robust_prune the neighbour list with distance 0.0
// for every entry, so the candidates are visited in ID order (node->candidate distances are still the true
// ones); dnc is shared scratch for those distances.
template <bool F16>
SG_DEVICE_INLINE void block_robust_prune(const void *rows, uint32_t dim, int metric, float alpha, uint32_t node,
                                         const uint32_t *cand, uint32_t n_cand, uint32_t gs, uint32_t *out,
                                         uint64_t *keys, uint32_t *kept, float *kept_ne, uint32_t *flag,
                                         bool by_id = false, float *dnc = nullptr)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    // distance node -> candidate, one warp per candidate
    for (uint32_t i = threadIdx.x; i < PRUNE_MAX_CAND; i += blockDim.x) keys[i] = KEY_SENTINEL;
    __syncthreads();
    for (uint32_t i = warp; i < n_cand; i += W) {
        const uint32_t c = cand[i];
        if (c == NONE || c == node) continue;
        if (by_id) {
            if (lane == 0) keys[i] = make_key(0u, c);
        } else {
            const float d = warp_pair_distance<F16>(rows, dim, node, c, metric, lane);
            if (lane == 0) keys[i] = make_key(order_key32(canon_nan(d)), c);
        }
    }
    __syncthreads();
    block_bitonic_sort(keys, PRUNE_MAX_CAND);          // (distance, id) ascending; duplicates become neighbours
    if (by_id) {
        for (uint32_t i = warp; i < n_cand; i += W) {
            const uint64_t kc = keys[i];
            if (kc == KEY_SENTINEL) break;
            const float d = warp_pair_distance<F16>(rows, dim, node, (uint32_t)kc, metric, lane);
            if (lane == 0) dnc[i] = canon_nan(d);
        }
        __syncthreads();
    }
    const float off = metric == SG_METRIC_NEG_DOT ? 1.0f : 0.0f;
    uint32_t m = 0;                                    // kept so far (uniform: every thread tracks it)
    for (uint32_t i = 0; i < n_cand && m < gs; ++i) {
        const uint64_t kc = keys[i];
        if (kc == KEY_SENTINEL) break;
        if (i > 0 && kc == keys[i - 1]) continue;      // the same id twice
        const uint32_t c = (uint32_t)kc;
        const float d_nc = by_id ? dnc[i] : key32_to_dist((uint32_t)(kc >> 32));
        if (threadIdx.x == 0) *flag = 0;
        __syncthreads();
        for (uint32_t e = warp; e < m; e += W) {
            const float d_ce = warp_pair_distance<F16>(rows, dim, c, kept[e], metric, lane);
            if (lane == 0 && alpha * (d_ce + off) <= (d_nc + off) && d_ce <= kept_ne[e]) *flag = 1;
        }
        __syncthreads();
        const bool occluded = *flag != 0;
        __syncthreads();
        if (!occluded) {
            if (threadIdx.x == 0) {
                kept[m] = c;
                kept_ne[m] = d_nc;
            }
            ++m;
            __syncthreads();
        }
    }
    for (uint32_t t = threadIdx.x; t < gs; t += blockDim.x) out[t] = t < m ? kept[t] : NONE;
    __syncthreads();
}

// one CTA per batch node: candidates = its search result + its current list -> new list (not yet applied)
template <bool F16>
__global__ void __launch_bounds__(128)
build_prune_kernel(const void *rows, uint32_t dim, int metric, float alpha, const uint32_t *graph, uint32_t gs,
                   const uint32_t *node_ids, uint32_t n_nodes, const uint32_t *res_ids, const uint32_t *res_counts,
                   uint32_t L, uint32_t *new_lists)
{
    __shared__ uint64_t keys[PRUNE_MAX_CAND];
    __shared__ uint32_t cand[PRUNE_MAX_CAND];
    __shared__ uint32_t kept[64];
    __shared__ float kept_ne[64];
    __shared__ uint32_t flag;
    const uint32_t b = blockIdx.x;
    if (b >= n_nodes) return;
    const uint32_t node = node_ids[b];
    const uint32_t rc = res_counts[b] < L ? res_counts[b] : L;
    uint32_t nc = rc + gs;
    if (nc > PRUNE_MAX_CAND) nc = PRUNE_MAX_CAND;
    for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x)
        cand[i] = i < rc ? res_ids[(size_t)b * L + i] : graph[(size_t)node * gs + (i - rc)];
    __syncthreads();
    block_robust_prune<F16>(rows, dim, metric, alpha, node, cand, nc, gs, new_lists + (size_t)b * gs, keys, kept, kept_ne,
                            &flag);
}

// graph[node] = new list, for the whole batch
__global__ void build_apply_kernel(uint32_t *graph, uint32_t gs, const uint32_t *node_ids, uint32_t n_nodes,
                                   const uint32_t *new_lists)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n_nodes * gs) return;
    const uint32_t b = (uint32_t)(t / gs), j = (uint32_t)(t % gs);
    graph[(size_t)node_ids[b] * gs + j] = new_lists[t];
}

// "ensure bidirectional edges" (:244-249): for every new edge node -> p, if p does not list node yet,
// queue node in p's inbox (up to `ecap` per batch; later arrivals are dropped) and note p as touched
__global__ void build_backedge_collect_kernel(const uint32_t *graph, uint32_t gs, const uint32_t *node_ids, uint32_t n_nodes,
                                              const uint32_t *new_lists, uint32_t *inbox, uint32_t *inbox_cnt, uint32_t ecap,
                                              uint32_t *touched, uint32_t *n_touched)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n_nodes * gs) return;
    const uint32_t node = node_ids[t / gs];
    const uint32_t p = new_lists[t];
    if (p == NONE) return;
    const uint32_t *row = graph + (size_t)p * gs;
    for (uint32_t j = 0; j < gs; ++j)
        if (row[j] == node) return;
    const uint32_t slot = atomicAdd(inbox_cnt + p, 1u);
    if (slot == 0) touched[atomicAdd(n_touched, 1u)] = p;
    if (slot < ecap) inbox[(size_t)p * ecap + slot] = node;
}

// one CTA per touched node p: append the inbox if it fits, else robust_prune(p, list + inbox) (:252-262)
template <bool F16>
__global__ void __launch_bounds__(128)
build_backedge_apply_kernel(const void *rows, uint32_t dim, int metric, float alpha, uint32_t *graph, uint32_t gs,
                            const uint32_t *touched, const uint32_t *n_touched, const uint32_t *inbox, uint32_t *inbox_cnt,
                            uint32_t ecap, bool by_id)
{
    __shared__ uint64_t keys[PRUNE_MAX_CAND];
    __shared__ uint32_t cand[PRUNE_MAX_CAND];
    __shared__ uint32_t kept[64];
    __shared__ float kept_ne[64];
    __shared__ uint32_t flag;
    __shared__ float dnc[PRUNE_MAX_CAND];
    for (uint32_t b = blockIdx.x; b < *n_touched; b += gridDim.x) {
        const uint32_t p = touched[b];
        uint32_t ne = inbox_cnt[p];
        if (ne > ecap) ne = ecap;
        uint32_t *row = graph + (size_t)p * gs;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < gs + ne; i += blockDim.x) cand[i] = i < gs ? row[i] : inbox[(size_t)p * ecap + (i - gs)];
        __syncthreads();
        uint32_t deg = 0;
        for (uint32_t j = 0; j < gs; ++j) deg += cand[j] != NONE;
        if (deg + ne <= gs) {
            for (uint32_t i = threadIdx.x; i < ne; i += blockDim.x) row[deg + i] = cand[gs + i];
        } else {
            block_robust_prune<F16>(rows, dim, metric, alpha, p, cand, gs + ne, gs, row, keys, kept, kept_ne, &flag, by_id, dnc);
        }
        __syncthreads();
        if (threadIdx.x == 0) inbox_cnt[p] = 0;
    }
}

// batch node ids first .. first+n-1
__global__ void iota_kernel(uint32_t *out, uint32_t first, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = first + i;
}

// per-list degree (graph -> deg[]) for sg_store_get_graph
__global__ void graph_degree_kernel(const uint32_t *graph, uint32_t gs, uint64_t n, uint32_t *deg)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t d = 0;
    for (uint32_t j = 0; j < gs; ++j) d += graph[i * gs + j] != NONE;
    deg[i] = d;
}

template <typename K>
cudaError_t set_smem(K kern, size_t smem)
{
    if (smem > 48 * 1024) return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return cudaSuccess;
}

template <bool F16, int ORDER, int METRIC>
cudaError_t launch_ann_t(const AnnArgs &a, cudaStream_t st)
{
    // warps per CTA: whatever keeps the most query warps resident per SM (227 KB of shared memory, 1 KB reserved per CTA)
    const size_t per_warp_plain = ann_warp_smem_bytes(a.dim, a.K), per_warp_staged = per_warp_plain + ann_stage_smem_bytes(a.dim, F16);
    uint32_t W = 4;
    {
        size_t best = 0;
        for (uint32_t w = 4; w >= 1; --w) {
            const size_t cta = w * per_warp_staged + 1024;
            const size_t resident = cta <= 227 * 1024 ? w * ((227 * 1024) / cta) : 0;
            if (resident > best) {
                best = resident;
                W = w;
            }
        }
    }
    const size_t staged = (size_t)W * per_warp_staged;
    // Staging buys latency (one HBM round trip per step) and costs occupancy (27 KB per warp at 384-d fp16: 8 warps per
    // SM).  Measured on 300k x 384: staged 2.2 vs 1.3 Mq/s at 1024 queries, 2.8 vs 3.7 Mq/s at 8192 -- large batches
    // run unstaged, where ~28 resident warps per SM hide the latency instead (2.9 TB/s of row reads).
    static const bool no_stage = getenv("SG_ANN_NO_STAGE") != nullptr;
    static const uint32_t stage_max_q = getenv("SG_ANN_STAGE_MAX_Q") ? (uint32_t)atoi(getenv("SG_ANN_STAGE_MAX_Q")) : 4096u;
    if (staged <= 200 * 1024 && !no_stage && a.nq < stage_max_q) {
        auto kern = ann_search_kernel<F16, ORDER, METRIC, true>;
        cudaError_t e = set_smem(kern, staged);
        if (e != cudaSuccess) return e;
        kern<<<(a.nq + W - 1) / W, W * 32, staged, st>>>(a);
    } else {
        W = 4;
        const size_t plain = (size_t)W * per_warp_plain;
        auto kern = ann_search_kernel<F16, ORDER, METRIC, false>;
        cudaError_t e = set_smem(kern, plain);
        if (e != cudaSuccess) return e;
        kern<<<(a.nq + W - 1) / W, W * 32, plain, st>>>(a);
    }
    count_launch();
    return cudaGetLastError();
}

template <bool F16, int ORDER, int METRIC>
cudaError_t launch_link_t(const void *rows, uint32_t dim, uint32_t *graph, uint32_t gs, uint32_t id, const uint32_t *res_ids,
                          const uint32_t *res_count, cudaStream_t st)
{
    const uint32_t W = 8;
    const size_t smem = (size_t)W * (((size_t)dim * 4 + 15) / 16 * 16 + 128 * 8);
    auto kern = graph_link_kernel<F16, ORDER, METRIC>;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<1, W * 32, smem, st>>>(rows, dim, graph, gs, id, res_ids, res_count);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

#define SG_ANN_DISPATCH(FN, ...)                                                                              \
    do {                                                                                                      \
        if (f16) {                                                                                            \
            if (order == SG_ORDER_AVX2) {                                                                     \
                if (metric == SG_METRIC_EUCLIDEAN) return FN<true, SG_ORDER_AVX2, SG_METRIC_EUCLIDEAN>(__VA_ARGS__);   \
                if (metric == SG_METRIC_COSINE) return FN<true, SG_ORDER_AVX2, SG_METRIC_COSINE>(__VA_ARGS__);         \
                return FN<true, SG_ORDER_AVX2, SG_METRIC_NEG_DOT>(__VA_ARGS__);                               \
            }                                                                                                 \
            if (metric == SG_METRIC_EUCLIDEAN) return FN<true, SG_ORDER_SCALAR, SG_METRIC_EUCLIDEAN>(__VA_ARGS__);     \
            if (metric == SG_METRIC_COSINE) return FN<true, SG_ORDER_SCALAR, SG_METRIC_COSINE>(__VA_ARGS__);           \
            return FN<true, SG_ORDER_SCALAR, SG_METRIC_NEG_DOT>(__VA_ARGS__);                                 \
        }                                                                                                     \
        if (order == SG_ORDER_AVX2) {                                                                         \
            if (metric == SG_METRIC_EUCLIDEAN) return FN<false, SG_ORDER_AVX2, SG_METRIC_EUCLIDEAN>(__VA_ARGS__);      \
            if (metric == SG_METRIC_COSINE) return FN<false, SG_ORDER_AVX2, SG_METRIC_COSINE>(__VA_ARGS__);            \
            return FN<false, SG_ORDER_AVX2, SG_METRIC_NEG_DOT>(__VA_ARGS__);                                  \
        }                                                                                                     \
        if (metric == SG_METRIC_EUCLIDEAN) return FN<false, SG_ORDER_SCALAR, SG_METRIC_EUCLIDEAN>(__VA_ARGS__);        \
        if (metric == SG_METRIC_COSINE) return FN<false, SG_ORDER_SCALAR, SG_METRIC_COSINE>(__VA_ARGS__);              \
        return FN<false, SG_ORDER_SCALAR, SG_METRIC_NEG_DOT>(__VA_ARGS__);                                    \
    } while (0)

cudaError_t launch_ann_search(const AnnArgs &a, bool f16, int order, int metric, cudaStream_t st)
{
    SG_ANN_DISPATCH(launch_ann_t, a, st);
}

cudaError_t launch_graph_link(const void *rows, bool f16, int order, int metric, uint32_t dim, uint32_t *graph, uint32_t gs,
                              uint32_t id, const uint32_t *res_ids, const uint32_t *res_count, cudaStream_t st)
{
    SG_ANN_DISPATCH(launch_link_t, rows, dim, graph, gs, id, res_ids, res_count, st);
}

cudaError_t launch_centroid(const void *rows, bool f16, uint32_t dim, uint64_t n, float *d_centroid, cudaStream_t st)
{
    if (f16) centroid_kernel<true><<<dim / 32, 256, 0, st>>>(rows, dim, n, d_centroid);
    else centroid_kernel<false><<<dim / 32, 256, 0, st>>>(rows, dim, n, d_centroid);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_graph_init(uint32_t *graph, uint32_t gs, uint64_t n, uint64_t seed, cudaStream_t st)
{
    graph_init_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(graph, gs, n, seed);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_iota(uint32_t *out, uint32_t first, uint32_t n, cudaStream_t st)
{
    iota_kernel<<<(n + 255) / 256, 256, 0, st>>>(out, first, n);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_graph_degree(const uint32_t *graph, uint32_t gs, uint64_t n, uint32_t *deg, cudaStream_t st)
{
    graph_degree_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(graph, gs, n, deg);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_build_prune(const void *rows, bool f16, uint32_t dim, int metric, float alpha, const uint32_t *graph,
                               uint32_t gs, const uint32_t *node_ids, uint32_t n_nodes, const uint32_t *res_ids,
                               const uint32_t *res_counts, uint32_t L, uint32_t *new_lists, cudaStream_t st)
{
    if (f16) build_prune_kernel<true><<<n_nodes, 128, 0, st>>>(rows, dim, metric, alpha, graph, gs, node_ids, n_nodes, res_ids,
                                                              res_counts, L, new_lists);
    else build_prune_kernel<false><<<n_nodes, 128, 0, st>>>(rows, dim, metric, alpha, graph, gs, node_ids, n_nodes, res_ids,
                                                            res_counts, L, new_lists);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_build_apply(uint32_t *graph, uint32_t gs, const uint32_t *node_ids, uint32_t n_nodes,
                               const uint32_t *new_lists, uint32_t *inbox, uint32_t *inbox_cnt, uint32_t ecap,
                               uint32_t *touched, uint32_t *n_touched, cudaStream_t st)
{
    const uint64_t total = (uint64_t)n_nodes * gs;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    build_apply_kernel<<<blocks, 256, 0, st>>>(graph, gs, node_ids, n_nodes, new_lists);
    build_backedge_collect_kernel<<<blocks, 256, 0, st>>>(graph, gs, node_ids, n_nodes, new_lists, inbox, inbox_cnt, ecap,
                                                         touched, n_touched);
    count_launch(2);
    return cudaGetLastError();
}

cudaError_t launch_build_backedges(const void *rows, bool f16, uint32_t dim, int metric, float alpha, uint32_t *graph,
                                   uint32_t gs, const uint32_t *touched, const uint32_t *n_touched, const uint32_t *inbox,
                                   uint32_t *inbox_cnt, uint32_t ecap, bool by_id, uint32_t grid, cudaStream_t st)
{
    if (f16) build_backedge_apply_kernel<true><<<grid, 128, 0, st>>>(rows, dim, metric, alpha, graph, gs, touched, n_touched,
                                                                    inbox, inbox_cnt, ecap, by_id);
    else build_backedge_apply_kernel<false><<<grid, 128, 0, st>>>(rows, dim, metric, alpha, graph, gs, touched, n_touched,
                                                                  inbox, inbox_cnt, ecap, by_id);
    count_launch();
    return cudaGetLastError();
}

}  // namespace sg
