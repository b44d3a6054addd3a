// encoder.cu -- MiniLM-L6 (BERT, 6 layers x 384 hidden x 12 heads, FFN 1536) sentence encoder.
//
// Replaces the ONNX Runtime session + pooling of MiniLMEmbedder (reference
// src/embeddings/minilm.rs): session.run :939-948 / :1065-1074 (third-party C++ BERT forward),
// masked mean-pool :959-978 / :1083-1108, finalize_pooled :846-878 (NaN scrub, L2 norm).
// Tokenisation stays on the host (tokenizers crate); the C ABI takes input_ids / attention_mask.
//
// B200 mapping.  Every dense contraction is a tcgen05 GEMM (C[T,N] = A[T,K] . W[N,K]^T, fp16
// operands, f32 accumulation in TMEM) with the element-wise work fused into the epilogue that
// reads the accumulator out of TMEM:
//   QKV projection    N=1152  tile 128x128   epilogue: + bias                       -> fp16
//   attention out     N=384   tile 128x384   epilogue: + bias + residual, LayerNorm  -> fp16
//   FFN up            N=1536  tile 128x128   epilogue: + bias, erf-GELU              -> fp16
//   FFN down          N=384   tile 128x384   epilogue: + bias + residual, LayerNorm  -> fp16 (+f32 last layer)
// (a 128x384 tile keeps a whole hidden row inside one CTA's TMEM, so LayerNorm needs no second
// kernel).  Operands are staged by TMA (SWIZZLE_128B, K-major -- PyTorch's [out,in] weight layout
// IS the K-major B operand).  Attention (L <= 512, d = 32; ~4 % of the FLOPs) and the embedding
// gather + LayerNorm run on CUDA cores; pooling follows the reference's summation order exactly.
// Parity: "unpinned" against the reference (no weights / ONNX Runtime in this environment); the
// kernel is checked against a PyTorch fp32 BertModel with the same weights (tests/test_encoder_gpu.py).
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "sg_internal.h"
#include "sg_tc.cuh"

namespace sg {
namespace {

constexpr uint32_t EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_RES_LN = 2;
constexpr uint32_t G_THREADS = 256;
constexpr uint32_t P_THREADS = 384;   // persistent GEMM: 4 control warps + 8 epilogue warps
constexpr uint32_t LN_THREADS = 640;  // LayerNorm-fused 128x384 tile: 4 control warps + 16 epilogue warps

struct GemmArgs {
    uint32_t K;            // reduction length (multiple of 64)
    uint32_t N;            // output width
    uint32_t stages;
    const float *bias;     // [N]
    const __half *residual;  // [M_pad, N] (EPI_BIAS_RES_LN)
    const float *gamma, *beta;
    float eps;
    __half *out;           // [M_pad, N]
    float *out_f32;        // optional f32 copy (EPI_BIAS_RES_LN)
    uint32_t rows;         // rows of the activation matrices (M_pad); the pair kernels may be handed a tile beyond it
    unsigned long long *trace;   // debug (SG_ENC_TRACE): [CTA][8] %globaltimer stamps of the LayerNorm GEMM's phases
};

SG_DEVICE_INLINE unsigned long long enc_gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// erf-GELU, x * Phi(x), in 11 instructions.  Phi(x) = 1 / (1 + exp(-logit Phi(x))) and logit Phi is odd and
// smooth, so it is fitted by x * P(x^2) (degree-9 odd minimax fit on |x| <= 6, coefficients pre-multiplied by
// -log2 e): max |error| of the GELU value 3.0e-6 over the whole real line -- two orders below the fp16
// rounding of the output -- with one ex2.approx and one rcp.approx and no slow-path branches.  (The ncu
// source view of the first version, Abramowitz-Stegun 7.1.26 with an IEEE reciprocal, showed the
// FFN-up epilogue issue-bound at 32 instructions per element: profiles/r01_encoder_ncu.md.)
SG_DEVICE_INLINE float gelu_erf(float x)
{
    const float x2 = fminf(x * x, 36.0f);             // beyond |x| = 6 the logit keeps growing linearly: Phi is 0 or 1 there anyway
    float p = fmaf(x2, -3.2290400e-06f, 8.8239125e-05f);
    p = fmaf(x2, p, 3.6026838e-04f);
    p = fmaf(x2, p, -1.0522668e-01f);
    p = fmaf(x2, p, -2.3020454e+00f);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * p));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}

// C tile = 128 rows x BN columns, one tile per CTA.  grid = (M_pad/128, N/BN).
// The LayerNorm form runs 16 epilogue warps (640 threads): warps w, w+4, w+8, w+12 share a TMEM lane quarter
// and each owns a quarter of the row's columns (one tile per CTA, so the epilogue is a pure latency chain:
// four column groups shorten it 4x); the residual rows are prefetched into registers while the main loop
// runs; the row statistics are combined through shared memory.
template <uint32_t BN, uint32_t EPI>
__global__ void __launch_bounds__(EPI == EPI_BIAS_RES_LN ? LN_THREADS : G_THREADS, BN <= 128 ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const GemmArgs g)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t A_BYTES = 128 * 128;              // 128 rows x 64 fp16
    constexpr uint32_t B_BYTES = BN * 128;
    constexpr uint32_t STAGE = A_BYTES + B_BYTES;
    constexpr uint32_t TCOLS = BN <= 128 ? 128 : (BN <= 256 ? 256 : 512);
    constexpr uint32_t NSPLIT = BN > 256 ? 2 : 1;        // MMA N <= 256: a 384-wide tile is 2 x 192
    constexpr uint32_t MMA_N = BN / NSPLIT;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t S = g.stages;
    const uint32_t nkb = g.K / 64;
    unsigned long long *trp = (EPI == EPI_BIAS_RES_LN && g.trace && threadIdx.x == 128) ? g.trace + (size_t)blockIdx.x * 8 : nullptr;
    if (trp) trp[0] = enc_gtimer();                      // kernel entered
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)S * STAGE);
    uint64_t *full_bar = bars, *empty_bar = bars + S, *done_bar = bars + 2 * S;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done_bar + 1);
    // [3][BN]: bias, gamma, beta of this tile's columns (16-byte aligned for float4 reads)
    float *par_s = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~static_cast<uintptr_t>(15));

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < S; ++i) {
            mbar_init(smem_u32(full_bar + i), 1);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        mbar_init(smem_u32(done_bar), 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (trp) trp[1] = enc_gtimer();                      // barriers + TMEM ready
    pdl_wait();                                            // activations come from the previous kernel of the forward
    pdl_trigger();
    if (trp) trp[2] = enc_gtimer();                      // previous kernel complete
    const uint32_t m0 = blockIdx.x * 128, n0 = blockIdx.y * BN;

    // Producer and MMA issuer: the WHOLE warp walks the (uniform) loop and one elected lane issues, so the smem
    // descriptors and barrier addresses live in uniform registers.  An `if (lane == 0)` loop makes them per-thread
    // values that are re-broadcast for every tcgen05.mma: ~700 cycles of issue per K-block against the 256 cycles the
    // four MMAs take -- the GEMMs of this encoder were issue-bound, not tensor-bound (measured on the corpus sweep).
    if (warp == 0) {
        uint32_t st = 0, ph = 0;
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            mbar_wait(smem_u32(empty_bar + st), ph ^ 1);
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(smem_u32(full_bar + st), STAGE);
                const uint32_t sa = smem_u32(smem + (size_t)st * STAGE);
                tma_load_2d(sa, &tmap_a, smem_u32(full_bar + st), (int32_t)(kb * 64), (int32_t)m0);
#pragma unroll
                for (uint32_t h = 0; h < NSPLIT; ++h)
                    tma_load_2d(sa + A_BYTES + h * MMA_N * 128, &tmap_w, smem_u32(full_bar + st), (int32_t)(kb * 64),
                                (int32_t)(n0 + h * MMA_N));
            }
            __syncwarp();
            if (++st == S) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_f16(128, MMA_N);
        uint32_t st = 0, ph = 0;
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            mbar_wait(smem_u32(full_bar + st), ph);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t sa = smem_u32(smem + (size_t)st * STAGE);
                const uint64_t da = make_sw128_desc(sa);
#pragma unroll
                for (uint32_t h = 0; h < NSPLIT; ++h) {
                    const uint64_t db = make_sw128_desc(sa + A_BYTES + h * MMA_N * 128);
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k)
                        umma_f16(tmem_base + h * MMA_N, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                }
                umma_commit(smem_u32(empty_bar + st));
                if (kb + 1 == nkb) umma_commit(smem_u32(done_bar));
            }
            __syncwarp();
            if (++st == S) { st = 0; ph ^= 1; }
        }
    } else if (warp >= 4) {
        // epilogue: thread t <-> output row m0 + t <-> TMEM lane t
        constexpr uint32_t EPI_T = EPI == EPI_BIAS_RES_LN ? 512 : 128;     // epilogue threads
        const uint32_t et = threadIdx.x - 128;
        const uint32_t t = et & 127;                                       // tile row = TMEM lane
        const uint32_t qtr = et >> 7;                                      // LN form: which quarter of the columns
        const uint32_t row = m0 + t;
        const uint32_t lane_base = ((warp & 3) * 32) << 16;
        // stage the per-column parameters while the main loop runs (named barrier over the epilogue warps)
        for (uint32_t i = et; i < BN; i += EPI_T) {
            par_s[i] = g.bias[n0 + i];
            if constexpr (EPI == EPI_BIAS_RES_LN) {
                par_s[BN + i] = g.gamma[n0 + i];
                par_s[2 * BN + i] = g.beta[n0 + i];
            }
        }
        // LN form: this thread's quarter of the residual row (BN/4 halfs) is fetched NOW, under the main loop
        uint4 rv[EPI == EPI_BIAS_RES_LN ? BN / 32 : 1];
        if constexpr (EPI == EPI_BIAS_RES_LN) {
            const uint4 *res4 = reinterpret_cast<const uint4 *>(g.residual + (size_t)row * g.N) + qtr * (BN / 32);
#pragma unroll
            for (uint32_t q = 0; q < BN / 32; ++q) rv[q] = res4[q];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_T) : "memory");
        mbar_wait(smem_u32(done_bar), 0);
        tc_fence_after();
        if (trp) trp[3] = enc_gtimer();                  // accumulator complete
        if constexpr (EPI == EPI_BIAS_RES_LN) {
            // BN == N == hidden: the whole row is in this thread's TMEM lane.  Pass 1: v = acc + bias +
            // residual is written BACK into TMEM while sum / sum of squares accumulate; pass 2 reads v from
            // TMEM only.  bias / gamma / beta were staged in shared memory during the main loop.
            const float4 *bias4 = reinterpret_cast<const float4 *>(par_s);
            const float4 *gam4 = reinterpret_cast<const float4 *>(par_s + BN);
            const float4 *bet4 = reinterpret_cast<const float4 *>(par_s + 2 * BN);
            float sum = 0.0f, sq = 0.0f;
            constexpr uint32_t CH = BN / 128;                   // 32-column chunks per quarter
#pragma unroll
            for (uint32_t ci = 0; ci < CH; ++ci) {
                const uint32_t c = qtr * CH + ci;
                uint32_t r[32];
                tmem_ld32(tmem_base + lane_base + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (uint32_t q = 0; q < 8; ++q) {
                    const float4 bb = bias4[c * 8 + q];
                    const __half2 *hp = reinterpret_cast<const __half2 *>(&rv[ci * 4 + (q >> 1)]) + (q & 1) * 2;
                    const float2 r01 = __half22float2(hp[0]), r23 = __half22float2(hp[1]);
                    const float v0 = __uint_as_float(r[4 * q]) + bb.x + r01.x, v1 = __uint_as_float(r[4 * q + 1]) + bb.y + r01.y;
                    const float v2 = __uint_as_float(r[4 * q + 2]) + bb.z + r23.x, v3 = __uint_as_float(r[4 * q + 3]) + bb.w + r23.y;
                    sum += (v0 + v1) + (v2 + v3);
                    sq += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
                    r[4 * q] = __float_as_uint(v0); r[4 * q + 1] = __float_as_uint(v1);
                    r[4 * q + 2] = __float_as_uint(v2); r[4 * q + 3] = __float_as_uint(v3);
                }
                tmem_st32(tmem_base + lane_base + c * 32, r);
            }
            tmem_st_wait();
            // combine the four quarters' statistics (each thread wrote its own quarter back into TMEM)
            float *part = par_s + 3 * BN;                       // [4][128][2]
            part[(qtr * 128 + t) * 2] = sum;
            part[(qtr * 128 + t) * 2 + 1] = sq;
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (trp) trp[4] = enc_gtimer();              // pass 1 (bias + residual + statistics) done by all
            sum = (part[t * 2] + part[(128 + t) * 2]) + (part[(256 + t) * 2] + part[(384 + t) * 2]);
            sq = (part[t * 2 + 1] + part[(128 + t) * 2 + 1]) + (part[(256 + t) * 2 + 1] + part[(384 + t) * 2 + 1]);
            const float mean = sum / (float)BN;
            const float var = fmaxf(sq / (float)BN - mean * mean, 0.0f);
            const float rstd = rsqrtf(var + g.eps);
#pragma unroll 1
            for (uint32_t c = qtr * CH; c < qtr * CH + CH; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + lane_base + c * 32, r);
                tmem_ld_wait();
                __align__(16) __half o[32];
#pragma unroll
                for (uint32_t q = 0; q < 8; ++q) {
                    const float4 gg = gam4[c * 8 + q], be = bet4[c * 8 + q];
                    float y[4];
                    y[0] = (__uint_as_float(r[4 * q]) - mean) * rstd * gg.x + be.x;
                    y[1] = (__uint_as_float(r[4 * q + 1]) - mean) * rstd * gg.y + be.y;
                    y[2] = (__uint_as_float(r[4 * q + 2]) - mean) * rstd * gg.z + be.z;
                    y[3] = (__uint_as_float(r[4 * q + 3]) - mean) * rstd * gg.w + be.w;
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) o[4 * q + j] = __float2half_rn(y[j]);
                    if (g.out_f32)
                        *reinterpret_cast<float4 *>(g.out_f32 + (size_t)row * g.N + c * 32 + 4 * q) = make_float4(y[0], y[1], y[2], y[3]);
                }
                uint4 *dst = reinterpret_cast<uint4 *>(g.out + (size_t)row * g.N + c * 32);
#pragma unroll
                for (uint32_t q = 0; q < 4; ++q) dst[q] = reinterpret_cast<const uint4 *>(o)[q];
            }
        } else {
#pragma unroll 1
            for (uint32_t c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + lane_base + c * 32, r);
                tmem_ld_wait();
                __align__(16) __half o[32];
#pragma unroll
                for (uint32_t j = 0; j < 32; ++j) {
                    float v = __uint_as_float(r[j]) + par_s[c * 32 + j];
                    if constexpr (EPI == EPI_BIAS_GELU) v = gelu_erf(v);
                    o[j] = __float2half_rn(v);
                }
                uint4 *dst = reinterpret_cast<uint4 *>(g.out + (size_t)row * g.N + n0 + c * 32);
#pragma unroll
                for (uint32_t q = 0; q < 4; ++q) dst[q] = reinterpret_cast<const uint4 *>(o)[q];
            }
        }
    }
    if (trp) trp[5] = enc_gtimer();                      // this thread's pass 2 done
    tc_fence_before();
    __syncthreads();
    if (trp) trp[6] = enc_gtimer();                      // all done
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TCOLS);
    }
}

// ---- LayerNorm GEMM split over a 3-CTA cluster ---------------------------------------------------
// out = LayerNorm(A . W^T + bias + residual) with N = 384 = 3 x 128: the three CTAs of a cluster each
// compute 128 of the 384 output columns of the same 128 rows (so the weight matrix is streamed by three
// SMs in parallel -- what matters for the single-query case, where the whole GEMM is ONE row tile), and
// exchange per-row (sum, sum of squares) partials through distributed shared memory before normalising.
// 8 epilogue warps: warps w and w+4 share TMEM lane quarter w%4 and own 64 columns each.
// (Tried and dropped: a cluster of 6 -- three column slices x two K halves, the upper halves' raw accumulators handed over
// through distributed shared memory -- so that six SMs stream FFN-down's 1.2 MB of weights: 233.6 vs 226.2 us per single
// query.  The small-batch forward is bound by its ~26 dependent launches, not by this kernel's weight stream.)
__global__ void __cluster_dims__(3, 1, 1) __launch_bounds__(P_THREADS, 1)
gemm_ln_cluster_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const GemmArgs g)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t BN = 128, A_BYTES = 128 * 128, B_BYTES = BN * 128, STAGE = A_BYTES + B_BYTES;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t S = g.stages, nkb = g.K / 64;
    const uint32_t crank = cluster_ctarank();
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)S * STAGE);
    uint64_t *full_bar = bars, *empty_bar = bars + S, *done_bar = bars + 2 * S;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done_bar + 1);
    float *par_s = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~static_cast<uintptr_t>(15));
    float *part = par_s + 3 * BN;        // [2 halves][128 rows][2]  this CTA's partial statistics (read by the peers)
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < S; ++i) {
            mbar_init(smem_u32(full_bar + i), 1);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        mbar_init(smem_u32(done_bar), 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                            // activations come from the previous kernel of the forward
    pdl_trigger();
    const uint32_t m0 = (blockIdx.x / 3) * 128, n0 = crank * BN;

    if (warp == 0) {                                      // warp-uniform loops, one elected lane issues (see gemm_tc_kernel)
        uint32_t st = 0, ph = 0;
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            mbar_wait(smem_u32(empty_bar + st), ph ^ 1);
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(smem_u32(full_bar + st), STAGE);
                const uint32_t sa = smem_u32(smem + (size_t)st * STAGE);
                tma_load_2d(sa, &tmap_a, smem_u32(full_bar + st), (int32_t)(kb * 64), (int32_t)m0);
                tma_load_2d(sa + A_BYTES, &tmap_w, smem_u32(full_bar + st), (int32_t)(kb * 64), (int32_t)n0);
            }
            __syncwarp();
            if (++st == S) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_f16(128, BN);
        uint32_t st = 0, ph = 0;
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            mbar_wait(smem_u32(full_bar + st), ph);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t sa = smem_u32(smem + (size_t)st * STAGE);
                const uint64_t da = make_sw128_desc(sa), db = make_sw128_desc(sa + A_BYTES);
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k) umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                umma_commit(smem_u32(empty_bar + st));
                if (kb + 1 == nkb) umma_commit(smem_u32(done_bar));
            }
            __syncwarp();
            if (++st == S) { st = 0; ph ^= 1; }
        }
    }
    float sum = 0.0f, sq = 0.0f;
    const uint32_t et = threadIdx.x - 128, t = et & 127, half = (et >> 7) & 1;
    const uint32_t row = m0 + t;
    const uint32_t lane_base = ((warp & 3) * 32) << 16;
    const bool epi = warp >= 4;
    if (epi) {
        for (uint32_t i = et; i < BN; i += 256) {
            par_s[i] = g.bias[n0 + i];
            par_s[BN + i] = g.gamma[n0 + i];
            par_s[2 * BN + i] = g.beta[n0 + i];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mbar_wait(smem_u32(done_bar), 0);
        tc_fence_after();
        const uint4 *res4 = reinterpret_cast<const uint4 *>(g.residual + (size_t)row * g.N + n0);
        const float4 *bias4 = reinterpret_cast<const float4 *>(par_s);
#pragma unroll 1
        for (uint32_t c = half * 2; c < half * 2 + 2; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_base + c * 32, r);
            uint4 rv[4];
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q) rv[q] = res4[c * 4 + q];
            tmem_ld_wait();
#pragma unroll
            for (uint32_t q = 0; q < 8; ++q) {
                const float4 bb = bias4[c * 8 + q];
                const __half2 *hp = reinterpret_cast<const __half2 *>(&rv[q >> 1]) + (q & 1) * 2;
                const float2 r01 = __half22float2(hp[0]), r23 = __half22float2(hp[1]);
                const float v0 = __uint_as_float(r[4 * q]) + bb.x + r01.x, v1 = __uint_as_float(r[4 * q + 1]) + bb.y + r01.y;
                const float v2 = __uint_as_float(r[4 * q + 2]) + bb.z + r23.x, v3 = __uint_as_float(r[4 * q + 3]) + bb.w + r23.y;
                sum += (v0 + v1) + (v2 + v3);
                sq += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
                r[4 * q] = __float_as_uint(v0); r[4 * q + 1] = __float_as_uint(v1);
                r[4 * q + 2] = __float_as_uint(v2); r[4 * q + 3] = __float_as_uint(v3);
            }
            tmem_st32(tmem_base + lane_base + c * 32, r);
        }
        tmem_st_wait();
        part[(half * 128 + t) * 2] = sum;
        part[(half * 128 + t) * 2 + 1] = sq;
    }
    // every CTA's partials are in its shared memory: cluster barrier (release/acquire), then read the peers'
    cluster_sync_all();
    if (epi) {
        sum = 0.0f;
        sq = 0.0f;
        const uint32_t my_part = smem_u32(part);
#pragma unroll
        for (uint32_t r = 0; r < 3; ++r) {
            const uint32_t base = mapa_u32(my_part, r);
#pragma unroll
            for (uint32_t h = 0; h < 2; ++h) {
                float a, b;
                asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(a) : "r"(base + ((h * 128 + t) * 2) * 4) : "memory");
                asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(b) : "r"(base + ((h * 128 + t) * 2 + 1) * 4) : "memory");
                sum += a;
                sq += b;
            }
        }
        const float mean = sum / (float)g.N;
        const float var = fmaxf(sq / (float)g.N - mean * mean, 0.0f);
        const float rstd = rsqrtf(var + g.eps);
        const float4 *gam4 = reinterpret_cast<const float4 *>(par_s + BN);
        const float4 *bet4 = reinterpret_cast<const float4 *>(par_s + 2 * BN);
#pragma unroll 1
        for (uint32_t c = half * 2; c < half * 2 + 2; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_base + c * 32, r);
            tmem_ld_wait();
            __align__(16) __half o[32];
#pragma unroll
            for (uint32_t q = 0; q < 8; ++q) {
                const float4 gg = gam4[c * 8 + q], be = bet4[c * 8 + q];
                float y[4];
                y[0] = (__uint_as_float(r[4 * q]) - mean) * rstd * gg.x + be.x;
                y[1] = (__uint_as_float(r[4 * q + 1]) - mean) * rstd * gg.y + be.y;
                y[2] = (__uint_as_float(r[4 * q + 2]) - mean) * rstd * gg.z + be.z;
                y[3] = (__uint_as_float(r[4 * q + 3]) - mean) * rstd * gg.w + be.w;
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) o[4 * q + j] = __float2half_rn(y[j]);
                if (g.out_f32)
                    *reinterpret_cast<float4 *>(g.out_f32 + (size_t)row * g.N + n0 + c * 32 + 4 * q) = make_float4(y[0], y[1], y[2], y[3]);
            }
            uint4 *dst = reinterpret_cast<uint4 *>(g.out + (size_t)row * g.N + n0 + c * 32);
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q) dst[q] = reinterpret_cast<const uint4 *>(o)[q];
        }
    }
    tc_fence_before();
    cluster_sync_all();              // peers may still be reading this CTA's partials
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

// ---- persistent 128x128-tile GEMM: epilogue of tile i overlaps the main loop of tile i+1 -----------
// Same roles as gemm_tc_kernel, but every CTA walks tiles t = blockIdx.x, +gridDim.x, ... with TWO
// accumulator stages in TMEM (2 x 128 columns), so the bias/GELU/convert/store epilogue (which costs
// about as much as the 6-K-block main loop of a hidden-384 GEMM) runs while the tensor core already
// works on the next tile.  96 KB of shared memory per CTA: two CTAs per SM (2 x 256 TMEM columns).
template <uint32_t EPI>
__global__ void __launch_bounds__(P_THREADS, 2)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                       const __grid_constant__ CUtensorMap tmap_out, const GemmArgs g, const uint32_t m_tiles,
                       const uint32_t n_tiles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t BN = 128, A_BYTES = 128 * 128, B_BYTES = BN * 128, STAGE = A_BYTES + B_BYTES, ACC = 2;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t S = g.stages, nkb = g.K / 64;
    // output staging: the 128x128 fp16 tile as two SWIZZLE_128B boxes (64 columns each), drained by TMA stores
    uint8_t *stage_out = smem + (size_t)S * STAGE;                      // 32 KB, 1024-aligned
    uint64_t *bars = reinterpret_cast<uint64_t *>(stage_out + 2 * 128 * 128);
    uint64_t *full_bar = bars, *empty_bar = bars + S, *tfull_bar = bars + 2 * S, *tempty_bar = tfull_bar + ACC;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + ACC);
    float *bias_s = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~static_cast<uintptr_t>(15));  // [ACC][128]
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < S; ++i) {
            mbar_init(smem_u32(full_bar + i), 1);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        for (uint32_t i = 0; i < ACC; ++i) {
            mbar_init(smem_u32(tfull_bar + i), 1);
            mbar_init(smem_u32(tempty_bar + i), 256);
        }
        mbar_fence_init();
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), ACC * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                            // activations come from the previous kernel of the forward
    pdl_trigger();
    const uint32_t total = m_tiles * n_tiles;
    const uint32_t my_tiles = blockIdx.x < total ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 0) {                                      // warp-uniform loops, one elected lane issues (see gemm_tc_kernel)
        uint32_t st = 0, ph = 0;
        for (uint32_t i = 0; i < my_tiles; ++i) {
            const uint32_t tile = blockIdx.x + i * gridDim.x, m0 = (tile / n_tiles) * 128, n0 = (tile % n_tiles) * BN;
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                mbar_wait(smem_u32(empty_bar + st), ph ^ 1);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(smem_u32(full_bar + st), STAGE);
                    const uint32_t sa = smem_u32(smem + (size_t)st * STAGE);
                    tma_load_2d(sa, &tmap_a, smem_u32(full_bar + st), (int32_t)(kb * 64), (int32_t)m0);
                    tma_load_2d(sa + A_BYTES, &tmap_w, smem_u32(full_bar + st), (int32_t)(kb * 64), (int32_t)n0);
                }
                __syncwarp();
                if (++st == S) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_f16(128, BN);
        uint32_t st = 0, ph = 0;
        for (uint32_t i = 0; i < my_tiles; ++i) {
            const uint32_t acc = i & 1, acc_ph = (i >> 1) & 1;
            mbar_wait(smem_u32(tempty_bar + acc), acc_ph ^ 1);
            tc_fence_after();
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                mbar_wait(smem_u32(full_bar + st), ph);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t sa = smem_u32(smem + (size_t)st * STAGE);
                    const uint64_t da = make_sw128_desc(sa), db = make_sw128_desc(sa + A_BYTES);
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k)
                        umma_f16(tmem_base + acc * BN, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(smem_u32(empty_bar + st));
                    if (kb + 1 == nkb) umma_commit(smem_u32(tfull_bar + acc));
                }
                __syncwarp();
                if (++st == S) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // 8 epilogue warps: warps w and w+4 share TMEM lane quarter (w % 4); the first takes columns
        // 0..63 of the tile, the second 64..127 -> twice the drain throughput of a 4-warp epilogue
        const uint32_t et = threadIdx.x - 128;                 // 0..255
        const uint32_t t = et & 127;                           // tile row = TMEM lane
        const uint32_t half = et >> 7;                         // which 64-column half
        const uint32_t lane_base = ((warp & 3) * 32) << 16;
        for (uint32_t i = 0; i < my_tiles; ++i) {
            const uint32_t tile = blockIdx.x + i * gridDim.x, m0 = (tile / n_tiles) * 128, n0 = (tile % n_tiles) * BN;
            const uint32_t acc = i & 1, acc_ph = (i >> 1) & 1;
            if (half == 0) bias_s[acc * BN + t] = g.bias[n0 + t];   // safe: this slot's previous tile was fully drained
            if (et == 0) tma_store_wait_read();               // the previous tile's TMA stores have consumed the staging
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(smem_u32(tfull_bar + acc), acc_ph);
            tc_fence_after();
            const uint32_t sw = (t & 7) << 4;
#pragma unroll 1
            for (uint32_t c = half * 2; c < half * 2 + 2; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + lane_base + acc * BN + c * 32, r);
                tmem_ld_wait();
                __align__(16) __half o[32];
#pragma unroll
                for (uint32_t j = 0; j < 32; ++j) {
                    float v = __uint_as_float(r[j]) + bias_s[acc * BN + c * 32 + j];
                    if constexpr (EPI == EPI_BIAS_GELU) v = gelu_erf(v);
                    o[j] = __float2half_rn(v);
                }
                // row t of box (c / 2): 16-byte pieces (c % 2) * 4 + q, XOR-swizzled like the TMA expects
                uint8_t *rowp = stage_out + (size_t)(c >> 1) * (128 * 128) + (size_t)t * 128;
#pragma unroll
                for (uint32_t q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4 *>(rowp + ((((c & 1) * 4 + q) << 4) ^ sw)) = reinterpret_cast<const uint4 *>(o)[q];
            }
            tc_fence_before();
            mbar_arrive(smem_u32(tempty_bar + acc));          // accumulator drained: the next tile may start
            fence_proxy_async();                              // generic smem writes -> visible to the TMA engine
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (et == 0) {
                tma_store_2d(&tmap_out, smem_u32(stage_out), (int32_t)n0, (int32_t)m0);
                tma_store_2d(&tmap_out, smem_u32(stage_out + 128 * 128), (int32_t)(n0 + 64), (int32_t)m0);
                tma_store_commit();
            }
        }
        if (et == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ACC * BN);
    }
}

// ---- persistent CTA-PAIR GEMM: 256 x BN tiles with tcgen05 cta_group::2 --------------------------------
// The 128x128-tile kernel above moves 32 KB from L2 per 2.1 MFLOP K-block: at 148 SMs x 15 TFLOP/s that is ~34 TB/s
// against an L2 that delivers ~12 (measured: QKV ran at 10.2 TB/s of L2->SM traffic, i.e. it was L2-bandwidth-bound at
// 46 % of the tensor peak).  A CTA pair computes a 256-row x BN-column tile: each CTA stages ITS 128 activation rows and
// HALF of the BN weight rows (the pair's MMA reads both halves), so the same FLOPs need 1.75x (BN = 192) fewer bytes.
// One CTA per SM, 6-deep ring (a whole K = 384 tile in flight), two accumulator stages in TMEM (columns 0 and 256),
// 8 epilogue warps per CTA drain its 128 rows, output through swizzled staging + TMA stores.  N = 1152 and 1536 are
// both multiples of 192: 384 / 512 pair tiles over 74 pairs.
template <uint32_t BN, uint32_t EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LN_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_out, const GemmArgs g, const uint32_t m_pairs,
                    const uint32_t n_tiles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t A_BYTES = 128 * 128, B_BYTES = (BN / 2) * 128, STAGE = A_BYTES + B_BYTES, ACC = 2, ACC_COLS = 256;
    constexpr uint32_t BOXES = BN / 64, CH = BN / 64;         // output boxes of 64 columns; 32-column chunks per half
    static_assert(BN % 64 == 0 && BN <= 256 && STAGE % 1024 == 0, "pair tile shape");
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t S = g.stages, nkb = g.K / 64;
    uint8_t *stage_out = smem + (size_t)S * STAGE;                      // 2 groups x BOXES x 16 KB, 1024-aligned
    uint64_t *bars = reinterpret_cast<uint64_t *>(stage_out + 2 * BOXES * 128 * 128);
    uint64_t *full_bar = bars, *empty_bar = bars + S, *tfull_bar = bars + 2 * S, *tempty_bar = tfull_bar + ACC;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + ACC);
    float *bias_s = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~static_cast<uintptr_t>(15));  // [ACC][BN]
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < S; ++i) {
            mbar_init(smem_u32(full_bar + i), 2);          // leader: its own expect_tx arrive + the peer's arrive
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        for (uint32_t i = 0; i < ACC; ++i) {
            mbar_init(smem_u32(tfull_bar + i), 1);
            mbar_init(smem_u32(tempty_bar + i), 2 * 256);  // leader: the epilogue threads of BOTH CTAs
        }
        mbar_fence_init();
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_out);
    }
    cluster_sync_all();                                    // both CTAs resident, barriers initialised
    if (warp == 2) tmem_alloc2(smem_u32(tmem_slot), 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const uint32_t total = m_pairs * n_tiles;
    const uint32_t my_tiles = cluster_id < total ? (total - cluster_id + n_clusters - 1) / n_clusters : 0;
    pdl_wait();                                            // the activations come from the previous kernel
    pdl_trigger();

    if (warp == 0) {
        // ===== TMA producer (both CTAs): my 128 activation rows + my half of the weight rows; credited to the leader =====
        uint32_t st = 0, ph = 0;
        for (uint32_t i = 0; i < my_tiles; ++i) {
            const uint32_t tile = cluster_id + i * n_clusters;
            const uint32_t m0 = (tile / n_tiles) * 256 + rank * 128, n0 = (tile % n_tiles) * BN + rank * (BN / 2);
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                mbar_wait(smem_u32(empty_bar + st), ph ^ 1);
                if (elect_one_sync()) {
                    const uint32_t lfull = mapa_u32(smem_u32(full_bar + st), 0);
                    const uint32_t sa = smem_u32(smem + (size_t)st * STAGE);
                    if (leader) mbar_arrive_expect_tx(smem_u32(full_bar + st), 2 * STAGE);
                    else mbar_arrive_cluster(lfull);
                    tma_load_2d_2cta(sa, &tmap_a, lfull, (int32_t)(kb * 64), (int32_t)m0);
                    tma_load_2d_2cta(sa + A_BYTES, &tmap_w, lfull, (int32_t)(kb * 64), (int32_t)n0);
                }
                __syncwarp();
                if (++st == S) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the leader's warp walks the loop, one elected lane drives the pair =====
        if (leader) {
            constexpr uint32_t idesc = make_idesc_f16(256, BN);
            uint32_t st = 0, ph = 0;
            for (uint32_t i = 0; i < my_tiles; ++i) {
                const uint32_t acc = i & 1, acc_ph = (i >> 1) & 1;
                mbar_wait(smem_u32(tempty_bar + acc), acc_ph ^ 1);
                tc_fence_after();
                for (uint32_t kb = 0; kb < nkb; ++kb) {
                    mbar_wait(smem_u32(full_bar + st), ph);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t sa = smem_u32(smem + (size_t)st * STAGE);
                        const uint64_t da = make_sw128_desc(sa), db = make_sw128_desc(sa + A_BYTES);
#pragma unroll
                        for (uint32_t k = 0; k < 4; ++k)
                            umma_f16_2cta(tmem_base + acc * ACC_COLS, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                        umma_commit_2cta(smem_u32(empty_bar + st), 3);                       // frees the stage in BOTH CTAs
                        if (kb + 1 == nkb) umma_commit_2cta(smem_u32(tfull_bar + acc), 3);   // accumulators ready in BOTH
                    }
                    __syncwarp();
                    if (++st == S) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): TWO groups of 8 warps, group = accumulator stage = tile parity, each with its own
        // staging buffer and named barrier -- the bias/GELU/convert code has little ILP, so 4 warps per scheduler instead of
        // 2 is what keeps the issue slots busy.  In a group, warps w and w+4 share TMEM lane quarter w%4 and own BN/2 columns.
        const uint32_t grp = (threadIdx.x - 128) >> 8;         // 0/1
        const uint32_t et = (threadIdx.x - 128) & 255;         // 0..255 within the group
        const uint32_t t = et & 127;                           // tile row = TMEM lane
        const uint32_t half = et >> 7;
        const uint32_t lane_base = ((warp & 3) * 32) << 16;
        const uint32_t ltempty = mapa_u32(smem_u32(tempty_bar), 0);
        uint8_t *my_stage = stage_out + (size_t)grp * BOXES * 128 * 128;
        const uint32_t acc = grp;
        for (uint32_t i = grp; i < my_tiles; i += 2) {
            const uint32_t tile = cluster_id + i * n_clusters;
            const uint32_t m0 = (tile / n_tiles) * 256 + rank * 128, n0 = (tile % n_tiles) * BN;
            const uint32_t acc_ph = (i >> 1) & 1;
            if (et < BN) bias_s[acc * BN + et] = g.bias[n0 + et];   // safe: this group's previous tile was fully drained
            if (et == 0) tma_store_wait_read();               // the group's previous TMA stores have consumed its staging
            asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");
            mbar_wait(smem_u32(tfull_bar + acc), acc_ph);
            tc_fence_after();
            const uint32_t sw = (t & 7) << 4;
#pragma unroll 1
            for (uint32_t c = half * CH; c < half * CH + CH; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + lane_base + acc * ACC_COLS + c * 32, r);
                tmem_ld_wait();
                __align__(16) __half o[32];
                const float4 *b4 = reinterpret_cast<const float4 *>(bias_s + acc * BN + c * 32);   // broadcast reads, 16 bytes each
#pragma unroll
                for (uint32_t q4 = 0; q4 < 8; ++q4) {
                    const float4 bb = b4[q4];
                    float v0 = __uint_as_float(r[4 * q4]) + bb.x, v1 = __uint_as_float(r[4 * q4 + 1]) + bb.y;
                    float v2 = __uint_as_float(r[4 * q4 + 2]) + bb.z, v3 = __uint_as_float(r[4 * q4 + 3]) + bb.w;
                    if constexpr (EPI == EPI_BIAS_GELU) {
                        v0 = gelu_erf(v0); v1 = gelu_erf(v1); v2 = gelu_erf(v2); v3 = gelu_erf(v3);
                    }
                    *reinterpret_cast<__half2 *>(o + 4 * q4) = __floats2half2_rn(v0, v1);
                    *reinterpret_cast<__half2 *>(o + 4 * q4 + 2) = __floats2half2_rn(v2, v3);
                }
                uint8_t *rowp = my_stage + (size_t)(c >> 1) * (128 * 128) + (size_t)t * 128;
#pragma unroll
                for (uint32_t q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4 *>(rowp + ((((c & 1) * 4 + q) << 4) ^ sw)) = reinterpret_cast<const uint4 *>(o)[q];
            }
            tc_fence_before();
            mbar_arrive_cluster(ltempty + 8 * acc);           // accumulator drained (counted in the leader)
            fence_proxy_async();                              // generic smem writes -> visible to the TMA engine
            asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");
            if (et == 0) {
#pragma unroll
                for (uint32_t bx = 0; bx < BOXES; ++bx)
                    tma_store_2d(&tmap_out, smem_u32(my_stage + bx * 128 * 128), (int32_t)(n0 + bx * 64), (int32_t)m0);
                tma_store_commit();
            }
        }
        if (et == 0) tma_store_wait_all();
    }
    tc_fence_before();
    cluster_sync_all();              // nobody may still signal or read the peer's barriers / TMEM
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, 512);
    }
}

// ---- LayerNorm GEMM on a CTA PAIR: out = LayerNorm(A . W^T + bias + residual), N = 384 ---------------------
// The one-CTA form above streams the whole weight matrix into EVERY CTA (64 KB per K-block for 6.3 MFLOP): measured,
// its main loop runs at the L2->SM bandwidth, not at the tensor rate (attention-out: 6.6-9 us for 2.7 us of MMAs), and
// its second epilogue pass spends 5-6 us in row-strided 16-byte global stores.  Here two CTAs own 256 rows: each stages
// its 128 activation rows and HALF of the weight rows per K-block (40 KB), the pair's MMAs (2 x N = 192) read both
// halves, every CTA ends up with its 128 rows x 384 columns in TMEM exactly as before -- and pass 2 writes the fp16 rows
// into the (now idle) ring as swizzled 64-column boxes that six TMA stores drain.
// (Tried and dropped: a persistent variant with the COLUMNS split over the cluster -- 128 x 192 per CTA, two accumulator
// stages, statistics exchanged through distributed shared memory.  It overlaps the epilogue with the next tile's MMAs but
// loads 40 KB per 3.1 MFLOP K-block: ~105 B/clk against the ~64 B/clk an SM can take from L2 -- 709 vs 651 us per 256 queries.)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LN_THREADS, 1)
gemm_ln_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_out, const GemmArgs g)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t BN = 384, A_BYTES = 128 * 128, BH_BYTES = 96 * 128, STAGE = A_BYTES + 2 * BH_BYTES;   // 40 KB
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t S = g.stages, nkb = g.K / 64;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)S * STAGE);
    uint64_t *full_bar = bars, *empty_bar = bars + S, *done_bar = bars + 2 * S;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done_bar + 1);
    float *par_s = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~static_cast<uintptr_t>(15));  // [3][BN]
    float *part = par_s + 3 * BN;                                                                                      // [4][128][2]
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < S; ++i) {
            mbar_init(smem_u32(full_bar + i), 2);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        mbar_init(smem_u32(done_bar), 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_out);
    }
    cluster_sync_all();
    if (warp == 2) tmem_alloc2(smem_u32(tmem_slot), 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t m0 = (blockIdx.x >> 1) * 256 + rank * 128;
    pdl_wait();                                            // activations come from the previous kernel of the forward
    pdl_trigger();

    if (warp == 0) {
        uint32_t st = 0, ph = 0;
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            mbar_wait(smem_u32(empty_bar + st), ph ^ 1);
            if (elect_one_sync()) {
                const uint32_t lfull = mapa_u32(smem_u32(full_bar + st), 0);
                const uint32_t sa = smem_u32(smem + (size_t)st * STAGE);
                if (leader) mbar_arrive_expect_tx(smem_u32(full_bar + st), 2 * STAGE);
                else mbar_arrive_cluster(lfull);
                tma_load_2d_2cta(sa, &tmap_a, lfull, (int32_t)(kb * 64), (int32_t)m0);
                // my half of the weight rows of each of the two N = 192 MMAs
                tma_load_2d_2cta(sa + A_BYTES, &tmap_w, lfull, (int32_t)(kb * 64), (int32_t)(rank * 96));
                tma_load_2d_2cta(sa + A_BYTES + BH_BYTES, &tmap_w, lfull, (int32_t)(kb * 64), (int32_t)(192 + rank * 96));
            }
            __syncwarp();
            if (++st == S) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        if (leader) {
            constexpr uint32_t idesc = make_idesc_f16(256, 192);
            uint32_t st = 0, ph = 0;
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                mbar_wait(smem_u32(full_bar + st), ph);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t sa = smem_u32(smem + (size_t)st * STAGE);
                    const uint64_t da = make_sw128_desc(sa);
#pragma unroll
                    for (uint32_t h = 0; h < 2; ++h) {
                        const uint64_t db = make_sw128_desc(sa + A_BYTES + h * BH_BYTES);
#pragma unroll
                        for (uint32_t k = 0; k < 4; ++k)
                            umma_f16_2cta(tmem_base + h * 192, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    }
                    umma_commit_2cta(smem_u32(empty_bar + st), 3);
                    if (kb + 1 == nkb) umma_commit_2cta(smem_u32(done_bar), 3);
                }
                __syncwarp();
                if (++st == S) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // epilogue: thread t <-> output row m0 + t <-> TMEM lane t; warps w, w+4, w+8, w+12 share a lane quarter and own a
        // quarter of the columns each (three 32-column chunks)
        const uint32_t et = threadIdx.x - 128;
        const uint32_t t = et & 127, qtr = et >> 7;
        const uint32_t row = m0 + t;
        const uint32_t lane_base = ((warp & 3) * 32) << 16;
        constexpr uint32_t CH = 3;
        for (uint32_t i = et; i < BN; i += 512) {
            par_s[i] = g.bias[i];
            par_s[BN + i] = g.gamma[i];
            par_s[2 * BN + i] = g.beta[i];
        }
        uint4 rv[12];                                       // this thread's quarter of the residual row, fetched under the main loop
        {
            const uint4 *res4 = reinterpret_cast<const uint4 *>(g.residual + (size_t)row * BN) + qtr * 12;
#pragma unroll
            for (uint32_t q = 0; q < 12; ++q) rv[q] = row < g.rows ? res4[q] : make_uint4(0, 0, 0, 0);
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        mbar_wait(smem_u32(done_bar), 0);
        tc_fence_after();
        const float4 *bias4 = reinterpret_cast<const float4 *>(par_s);
        const float4 *gam4 = reinterpret_cast<const float4 *>(par_s + BN);
        const float4 *bet4 = reinterpret_cast<const float4 *>(par_s + 2 * BN);
        float sum = 0.0f, sq = 0.0f;
#pragma unroll
        for (uint32_t ci = 0; ci < CH; ++ci) {
            const uint32_t c = qtr * CH + ci;
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_base + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (uint32_t q = 0; q < 8; ++q) {
                const float4 bb = bias4[c * 8 + q];
                const __half2 *hp = reinterpret_cast<const __half2 *>(&rv[ci * 4 + (q >> 1)]) + (q & 1) * 2;
                const float2 r01 = __half22float2(hp[0]), r23 = __half22float2(hp[1]);
                const float v0 = __uint_as_float(r[4 * q]) + bb.x + r01.x, v1 = __uint_as_float(r[4 * q + 1]) + bb.y + r01.y;
                const float v2 = __uint_as_float(r[4 * q + 2]) + bb.z + r23.x, v3 = __uint_as_float(r[4 * q + 3]) + bb.w + r23.y;
                sum += (v0 + v1) + (v2 + v3);
                sq += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
                r[4 * q] = __float_as_uint(v0); r[4 * q + 1] = __float_as_uint(v1);
                r[4 * q + 2] = __float_as_uint(v2); r[4 * q + 3] = __float_as_uint(v3);
            }
            tmem_st32(tmem_base + lane_base + c * 32, r);
        }
        tmem_st_wait();
        part[(qtr * 128 + t) * 2] = sum;
        part[(qtr * 128 + t) * 2 + 1] = sq;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        sum = (part[t * 2] + part[(128 + t) * 2]) + (part[(256 + t) * 2] + part[(384 + t) * 2]);
        sq = (part[t * 2 + 1] + part[(128 + t) * 2 + 1]) + (part[(256 + t) * 2 + 1] + part[(384 + t) * 2 + 1]);
        const float mean = sum / (float)BN;
        const float var = fmaxf(sq / (float)BN - mean * mean, 0.0f);
        const float rstd = rsqrtf(var + g.eps);
        // pass 2: normalise from TMEM into the idle ring as six swizzled 64-column boxes (box = 128 rows x 128 bytes)
        uint8_t *stage_out = smem;
        const uint32_t sw = (t & 7) << 4;
#pragma unroll 1
        for (uint32_t c = qtr * CH; c < qtr * CH + CH; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_base + c * 32, r);
            tmem_ld_wait();
            __align__(16) __half o[32];
#pragma unroll
            for (uint32_t q = 0; q < 8; ++q) {
                const float4 gg = gam4[c * 8 + q], be = bet4[c * 8 + q];
                float y[4];
                y[0] = (__uint_as_float(r[4 * q]) - mean) * rstd * gg.x + be.x;
                y[1] = (__uint_as_float(r[4 * q + 1]) - mean) * rstd * gg.y + be.y;
                y[2] = (__uint_as_float(r[4 * q + 2]) - mean) * rstd * gg.z + be.z;
                y[3] = (__uint_as_float(r[4 * q + 3]) - mean) * rstd * gg.w + be.w;
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) o[4 * q + j] = __float2half_rn(y[j]);
                if (g.out_f32 && row < g.rows)
                    *reinterpret_cast<float4 *>(g.out_f32 + (size_t)row * BN + c * 32 + 4 * q) = make_float4(y[0], y[1], y[2], y[3]);
            }
            uint8_t *rowp = stage_out + (size_t)(c >> 1) * (128 * 128) + (size_t)t * 128;
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q)
                *reinterpret_cast<uint4 *>(rowp + ((((c & 1) * 4 + q) << 4) ^ sw)) = reinterpret_cast<const uint4 *>(o)[q];
        }
        fence_proxy_async();                                  // generic smem writes -> visible to the TMA engine
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (et == 0) {
#pragma unroll
            for (uint32_t bx = 0; bx < 6; ++bx)
                tma_store_2d(&tmap_out, smem_u32(stage_out + bx * 128 * 128), (int32_t)(bx * 64), (int32_t)m0);
            tma_store_commit();
            tma_store_wait_all();
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, 512);
    }
}

// ---- embeddings: word + position + token_type(0), LayerNorm -> fp16 ---------------------------------
// one warp per token; rows >= T (padding up to a multiple of 128) are zeroed.
__global__ void embed_ln_kernel(const int64_t *__restrict__ ids, uint32_t T, uint32_t T_pad, uint32_t L, uint32_t H,
                                uint32_t vocab, const float *__restrict__ word, const float *__restrict__ pos,
                                const float *__restrict__ type0, const float *__restrict__ gamma,
                                const float *__restrict__ beta, float eps, __half *__restrict__ out)
{
    const uint32_t tok = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    pdl_wait();
    pdl_trigger();
    if (tok >= T_pad) return;
    if (tok >= T) {
        for (uint32_t d = lane; d < H; d += 32) out[(size_t)tok * H + d] = __float2half_rn(0.0f);
        return;
    }
    int64_t id = ids[tok];
    if (id < 0 || id >= (int64_t)vocab) id = 0;
    const uint32_t p = tok % L;
    // lane owns 4 consecutive dimensions per 128-wide step: 16-byte loads of the three embedding rows (H % 128 == 0 for
    // the models this kernel serves; otherwise the scalar walk below)
    if ((H & 127u) == 0 && H <= 512) {
        const float4 *w4 = reinterpret_cast<const float4 *>(word + (size_t)id * H);
        const float4 *p4 = reinterpret_cast<const float4 *>(pos + (size_t)p * H);
        const float4 *t4 = reinterpret_cast<const float4 *>(type0);
        const float4 *g4 = reinterpret_cast<const float4 *>(gamma), *b4 = reinterpret_cast<const float4 *>(beta);
        float4 v4[4];
        const uint32_t steps = H / 128;
        float sum = 0.0f;
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) {
            if (i < steps) {
                const float4 a = __ldg(w4 + i * 32 + lane), b = __ldg(p4 + i * 32 + lane), c = __ldg(t4 + i * 32 + lane);
                v4[i] = make_float4(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z, a.w + b.w + c.w);
                sum += (v4[i].x + v4[i].y) + (v4[i].z + v4[i].w);
            }
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        const float mean = sum / (float)H;
        float var = 0.0f;
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) {
            if (i < steps) {
                const float dx = v4[i].x - mean, dy = v4[i].y - mean, dz = v4[i].z - mean, dw = v4[i].w - mean;
                var += (dx * dx + dy * dy) + (dz * dz + dw * dw);
            }
        }
        for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xFFFFFFFFu, var, o);
        const float rstd = rsqrtf(var / (float)H + eps);
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) {
            if (i < steps) {
                const float4 gg = __ldg(g4 + i * 32 + lane), bb = __ldg(b4 + i * 32 + lane);
                const __half2 lo = __floats2half2_rn((v4[i].x - mean) * rstd * gg.x + bb.x, (v4[i].y - mean) * rstd * gg.y + bb.y);
                const __half2 hi = __floats2half2_rn((v4[i].z - mean) * rstd * gg.z + bb.z, (v4[i].w - mean) * rstd * gg.w + bb.w);
                uint2 o2;
                o2.x = *reinterpret_cast<const uint32_t *>(&lo);
                o2.y = *reinterpret_cast<const uint32_t *>(&hi);
                *reinterpret_cast<uint2 *>(out + (size_t)tok * H + (i * 32 + lane) * 4) = o2;
            }
        }
        return;
    }
    float v[16];                                            // H <= 512
    float sum = 0.0f;
    uint32_t n = 0;
    for (uint32_t d = lane; d < H; d += 32, ++n) {
        v[n] = word[(size_t)id * H + d] + pos[(size_t)p * H + d] + type0[d];
        sum += v[n];
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    const float mean = sum / (float)H;
    float var = 0.0f;
    for (uint32_t i = 0; i < n; ++i) var += (v[i] - mean) * (v[i] - mean);
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xFFFFFFFFu, var, o);
    const float rstd = rsqrtf(var / (float)H + eps);
    n = 0;
    for (uint32_t d = lane; d < H; d += 32, ++n)
        out[(size_t)tok * H + d] = __float2half_rn((v[n] - mean) * rstd * gamma[d] + beta[d]);
}

// ---- attention: one CTA per (sequence, head); d = 32; one WARP per query row -----------------------
// scores: lane j handles keys j, j+32, ...; softmax statistics by warp shuffles; P.V: lane d owns output
// dimension d and the probabilities are broadcast by shuffle.  Masked keys get exp(-inf) = 0, exactly
// like the additive -inf mask of the reference graph.  K is held TRANSPOSED in shared memory
// ([32][L+1] floats) so that the per-lane key reads are conflict-free; V as [L][32] floats.
__global__ void __launch_bounds__(128)
attention_kernel(const __half *__restrict__ qkv, const int64_t *__restrict__ mask, uint32_t L, uint32_t H,
                 uint32_t heads, __half *__restrict__ ctx)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    constexpr uint32_t D = 32;
    pdl_wait();
    pdl_trigger();
    const uint32_t b = blockIdx.x / heads, h = blockIdx.x % heads;
    const uint32_t LP = L + 1;
    float *kt = reinterpret_cast<float *>(smem_raw);          // [32][L+1]
    float *vs = kt + (size_t)D * LP;                         // [L][32]
    float *ps = vs + (size_t)L * D;                          // [4 warps][L] probabilities
    uint8_t *ms = reinterpret_cast<uint8_t *>(ps + 4 * (size_t)L);
    const size_t row_stride = 3 * (size_t)H;
    const __half *base = qkv + (size_t)b * L * row_stride;
    for (uint32_t i = threadIdx.x; i < L * D; i += blockDim.x) {
        const uint32_t j = i >> 5, d = i & 31;
        kt[(size_t)d * LP + j] = __half2float(base[j * row_stride + H + h * D + d]);
        vs[i] = __half2float(base[j * row_stride + 2 * H + h * D + d]);
    }
    for (uint32_t j = threadIdx.x; j < L; j += blockDim.x) ms[j] = mask[(size_t)b * L + j] == 1 ? 1 : 0;
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float scale = 0.17677669529663687f;              // 1/sqrt(32)
    const float NEG_INF = -__uint_as_float(0x7F800000u);
    float *pw = ps + (size_t)warp * L;
    for (uint32_t i = warp; i < L; i += 4) {
        const float qd = __half2float(base[i * row_stride + h * D + lane]) * scale;   // lane d holds q[d]
        float mx = NEG_INF;
        for (uint32_t j0 = 0; j0 < L; j0 += 32) {
            const uint32_t j = j0 + lane;
            float sc = 0.0f;
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) {
                const float qv = __shfl_sync(0xFFFFFFFFu, qd, d);
                if (j < L) sc = fmaf(qv, kt[(size_t)d * LP + j], sc);
            }
            if (j < L) {
                if (!ms[j]) sc = NEG_INF;
                pw[j] = sc;
                mx = fmaxf(mx, sc);
            }
        }
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        float lsum = 0.0f;
        __syncwarp();
        for (uint32_t j = lane; j < L; j += 32) {
            const float p = mx > NEG_INF ? __expf(pw[j] - mx) : 0.0f;
            pw[j] = p;
            lsum += p;
        }
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xFFFFFFFFu, lsum, o);
        __syncwarp();
        float acc = 0.0f;                                    // output dimension `lane`
        for (uint32_t j = 0; j < L; ++j) acc = fmaf(pw[j], vs[(size_t)j * D + lane], acc);
        const float inv = lsum > 0.0f ? 1.0f / lsum : 0.0f;
        ctx[((size_t)b * L + i) * H + h * D + lane] = __float2half_rn(acc * inv);
        __syncwarp();
    }
}

// ---- attention, L <= 128: register-blocked S = QK^T, softmax, O = PV in shared memory --------------
// grid = (sequences * heads, ceil(L / 64)); 128 threads; each CTA owns a block of 64 query rows.
// Phase A: every thread computes 4 rows x 4 (strided) key columns of S per micro-tile -- operands in
// padded f32 shared memory, conflict-free; Phase B: one warp per row does the masked softmax;
// Phase C: 4 rows x 4 (strided) output dims per thread.  ~4k instructions per thread with full ILP.
__global__ void __launch_bounds__(128)
attention_small_kernel(const __half *__restrict__ qkv, const int64_t *__restrict__ mask, uint32_t L, uint32_t H,
                       uint32_t heads, __half *__restrict__ ctx)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    constexpr uint32_t D = 32, DP = 33, QB = 64;
    pdl_wait();
    pdl_trigger();
    const uint32_t b = blockIdx.x / heads, h = blockIdx.x % heads;
    const uint32_t q0 = blockIdx.y * QB;
    const uint32_t L4 = (L + 3) & ~3u, SP = L4 + 1, CS = L4 / 4;
    float *Ks = reinterpret_cast<float *>(smem_raw);        // [L4][33]
    float *Vs = Ks + (size_t)L4 * DP;                       // [L4][33]
    float *Qs = Vs + (size_t)L4 * DP;                       // [64][33]
    float *Ss = Qs + (size_t)QB * DP;                       // [64][L4+1]
    const size_t row_stride = 3 * (size_t)H;
    const __half *base = qkv + (size_t)b * L * row_stride;
    const float NEG_INF = -__uint_as_float(0x7F800000u);
    // loads: 16-byte pieces (8 halfs)
    for (uint32_t i = threadIdx.x; i < L4 * 4; i += blockDim.x) {
        const uint32_t j = i >> 2, part = i & 3;
        uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
        if (j < L) {
            kv = *reinterpret_cast<const uint4 *>(base + j * row_stride + H + h * D + part * 8);
            vv = *reinterpret_cast<const uint4 *>(base + j * row_stride + 2 * H + h * D + part * 8);
        }
        const __half2 *kh = reinterpret_cast<const __half2 *>(&kv), *vh = reinterpret_cast<const __half2 *>(&vv);
#pragma unroll
        for (uint32_t e = 0; e < 4; ++e) {
            const float2 kf = __half22float2(kh[e]), vf = __half22float2(vh[e]);
            Ks[(size_t)j * DP + part * 8 + 2 * e] = kf.x; Ks[(size_t)j * DP + part * 8 + 2 * e + 1] = kf.y;
            Vs[(size_t)j * DP + part * 8 + 2 * e] = vf.x; Vs[(size_t)j * DP + part * 8 + 2 * e + 1] = vf.y;
        }
    }
    const float scale = 0.17677669529663687f;              // 1/sqrt(32)
    for (uint32_t i = threadIdx.x; i < QB * 4; i += blockDim.x) {
        const uint32_t r = i >> 2, part = i & 3;
        uint4 qv = make_uint4(0, 0, 0, 0);
        if (q0 + r < L) qv = *reinterpret_cast<const uint4 *>(base + (q0 + r) * row_stride + h * D + part * 8);
        const __half2 *qh = reinterpret_cast<const __half2 *>(&qv);
#pragma unroll
        for (uint32_t e = 0; e < 4; ++e) {
            const float2 qf = __half22float2(qh[e]);
            Qs[(size_t)r * DP + part * 8 + 2 * e] = qf.x * scale; Qs[(size_t)r * DP + part * 8 + 2 * e + 1] = qf.y * scale;
        }
    }
    __syncthreads();
    // Phase A: S block [64][L4]
    const uint32_t n_tiles = (QB / 4) * CS;
    for (uint32_t tIdx = threadIdx.x; tIdx < n_tiles; tIdx += blockDim.x) {
        const uint32_t ti = tIdx / CS, cj = tIdx - ti * CS;
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
#pragma unroll 8
        for (uint32_t d = 0; d < D; ++d) {
            float qv[4], kv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) qv[r] = Qs[(size_t)(4 * ti + r) * DP + d];
#pragma unroll
            for (int c = 0; c < 4; ++c) kv[c] = Ks[(size_t)(cj + c * CS) * DP + d];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(qv[r], kv[c], acc[r][c]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t j = cj + c * CS;
            const bool live = j < L && mask[(size_t)b * L + j] == 1;
#pragma unroll
            for (int r = 0; r < 4; ++r) Ss[(size_t)(4 * ti + r) * SP + j] = live ? acc[r][c] : NEG_INF;
        }
    }
    __syncthreads();
    // Phase B: softmax, one warp per row
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t r = warp; r < QB; r += 4) {
        float *srow = Ss + (size_t)r * SP;
        float mx = NEG_INF;
        for (uint32_t j = lane; j < L4; j += 32) mx = fmaxf(mx, srow[j]);
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        float sum = 0.0f;
        for (uint32_t j = lane; j < L4; j += 32) {
            const float p = mx > NEG_INF ? __expf(srow[j] - mx) : 0.0f;
            srow[j] = p;
            sum += p;
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
        for (uint32_t j = lane; j < L4; j += 32) srow[j] *= inv;
    }
    __syncthreads();
    // Phase C: O block [64][32] = P . V ; thread -> 4 rows x dims {dj, dj+8, dj+16, dj+24}
    {
        const uint32_t ti = threadIdx.x >> 3, dj = threadIdx.x & 7;
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
#pragma unroll 4
        for (uint32_t j = 0; j < L4; ++j) {
            float pv[4], vv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) pv[r] = Ss[(size_t)(4 * ti + r) * SP + j];
#pragma unroll
            for (int c = 0; c < 4; ++c) vv[c] = Vs[(size_t)j * DP + dj + 8 * c];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(pv[r], vv[c], acc[r][c]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t i = q0 + 4 * ti + r;
            if (i < L) {
#pragma unroll
                for (int c = 0; c < 4; ++c) ctx[((size_t)b * L + i) * H + h * D + dj + 8 * c] = __float2half_rn(acc[r][c]);
            }
        }
    }
}

// ---- attention on the legacy tensor path (mma.sync m16n8k16), L <= 128 ---------------------------
// grid = (sequences * heads, ceil(L / 64)); 4 warps, each owns 16 query rows.  S = Q K^T (2 k-steps over
// d = 32) stays in registers, the masked softmax is done on the accumulator fragments (quad shuffles),
// and the f32 probabilities are re-packed in registers as the A operand of P V (the C fragment of two
// adjacent 8-key tiles IS the A fragment of one 16-key step).  K sits row-major in shared memory
// (pitch 40 halfs), V transposed ([32][L16 + 8]); both pitches make the fragment loads conflict-free.
// Attention is ~4 % of the encoder's FLOPs; this kernel exists to get it out of the way of the tcgen05 GEMMs.
SG_DEVICE_INLINE void mma_m16n8k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
SG_DEVICE_INLINE uint32_t pack_half2(float lo, float hi)
{
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}

SG_DEVICE_INLINE float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
SG_DEVICE_INLINE void ldmatrix_x4(uint32_t (&r)[4], uint32_t smem_addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr));
}
SG_DEVICE_INLINE void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t smem_addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr));
}

SG_DEVICE_INLINE void cp_async16(uint32_t dst_smem, const void *src, bool valid)
{
    const uint32_t n = valid ? 16u : 0u;               // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(n) : "memory");
}
SG_DEVICE_INLINE void cp_async8(uint32_t dst_smem, const void *src, bool valid)
{
    const uint32_t n = valid ? 8u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src), "r"(n) : "memory");
}
SG_DEVICE_INLINE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
SG_DEVICE_INLINE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One warp's 16 query rows of one (sequence, head): S = Q K^T on mma.sync, masked softmax on the fragments, O = P V.
// Qw = the warp's 16 rows of Q, Ks / Vs = all keys / values, row-major [row][32] halfs at a pitch of 40 halfs; mk = the
// sequence's raw attention mask (0 beyond L); row_w0 = sequence position of the warp's first row; ctx_bh = output base of
// this (sequence, head): row r lands at ctx_bh + r * H.
template <uint32_t KS>
SG_DEVICE_INLINE void attend_rows(const __half *Qw, const __half *Ks, const __half *Vs, const int64_t *mk, uint32_t lane,
                                  uint32_t row_w0, uint32_t L, __half *ctx_bh, uint32_t H)
{
    constexpr uint32_t NT = KS * 2, P = 40;
    const uint32_t g = lane >> 2, tq = lane & 3;       // fragment row group / quad lane
    const float NEG_INF = -__uint_as_float(0x7F800000u);
    const uint32_t r0 = row_w0 + g;                    // this thread's rows: r0 and r0 + 8
    {
    // Q fragments for both k-steps (d 0..15, 16..31): matrices (rows 0-7 | 8-15) x (k 0-7 | 8-15)
    uint32_t qa[2][4];
    {
        const uint32_t qaddr = smem_u32(Qw + (size_t)((lane & 7) + ((lane >> 3) & 1) * 8) * P + (lane >> 4) * 8);
        ldmatrix_x4(qa[0], qaddr);
        ldmatrix_x4(qa[1], qaddr + 32);
    }
    // S = Q K^T: per 8-key tile one ldmatrix.x4 = the B fragments of both k-steps
    float sacc[NT][4];
    const uint32_t kaddr = smem_u32(Ks + (size_t)(lane & 7) * P + (lane >> 3) * 8);
#pragma unroll
    for (uint32_t nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) sacc[nt][e] = 0.0f;
        uint32_t kb[4];
        ldmatrix_x4(kb, kaddr + nt * 8 * P * 2);
        mma_m16n8k16(sacc[nt], qa[0], kb[0], kb[1]);
        mma_m16n8k16(sacc[nt], qa[1], kb[2], kb[3]);
    }
    // masked softmax on the fragments: rows r0 (elements 0,1) and r0+8 (elements 2,3).  Scores are kept in the log2
    // domain (1/sqrt(d) and log2(e) folded into one factor) so that a probability is ONE ex2.approx of a difference;
    // a fully masked row gets max = 0, and 2^(-inf) = 0 does the rest.
    const float scale = 0.17677669529663687f * 1.4426950408889634f;
    float mx0 = NEG_INF, mx1 = NEG_INF;
#pragma unroll
    for (uint32_t nt = 0; nt < NT; ++nt) {
        const longlong2 mm = *reinterpret_cast<const longlong2 *>(mk + nt * 8 + tq * 2);
        const float bx = mm.x == 1 ? 0.0f : NEG_INF, by = mm.y == 1 ? 0.0f : NEG_INF;
        sacc[nt][0] = fmaf(sacc[nt][0], scale, bx); sacc[nt][1] = fmaf(sacc[nt][1], scale, by);
        sacc[nt][2] = fmaf(sacc[nt][2], scale, bx); sacc[nt][3] = fmaf(sacc[nt][3], scale, by);
        mx0 = fmaxf(mx0, fmaxf(sacc[nt][0], sacc[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(sacc[nt][2], sacc[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, 2));
    if (!(mx0 > NEG_INF)) mx0 = 0.0f;
    if (!(mx1 > NEG_INF)) mx1 = 0.0f;
    float sum0 = 0.0f, sum1 = 0.0f;
#pragma unroll
    for (uint32_t nt = 0; nt < NT; ++nt) {
        sacc[nt][0] = ex2_approx(sacc[nt][0] - mx0);
        sacc[nt][1] = ex2_approx(sacc[nt][1] - mx0);
        sacc[nt][2] = ex2_approx(sacc[nt][2] - mx1);
        sacc[nt][3] = ex2_approx(sacc[nt][3] - mx1);
        sum0 += sacc[nt][0] + sacc[nt][1];
        sum1 += sacc[nt][2] + sacc[nt][3];
    }
    sum0 += __shfl_xor_sync(0xFFFFFFFFu, sum0, 1); sum0 += __shfl_xor_sync(0xFFFFFFFFu, sum0, 2);
    sum1 += __shfl_xor_sync(0xFFFFFFFFu, sum1, 1); sum1 += __shfl_xor_sync(0xFFFFFFFFu, sum1, 2);
    // O = P V : per 16-key step, the P A-fragment is built from two adjacent S tiles; V^T fragments by
    // ldmatrix.trans: matrices (keys 0-7 | 8-15) x (d tile dt | dt + 1)
    float oacc[4][4];
#pragma unroll
    for (uint32_t dt = 0; dt < 4; ++dt)
#pragma unroll
        for (int e = 0; e < 4; ++e) oacc[dt][e] = 0.0f;
    const uint32_t vaddr = smem_u32(Vs + (size_t)((lane & 7) + ((lane >> 3) & 1) * 8) * P + (lane >> 4) * 8);
#pragma unroll
    for (uint32_t ks = 0; ks < KS; ++ks) {
        uint32_t pa[4];
        pa[0] = pack_half2(sacc[2 * ks][0], sacc[2 * ks][1]);
        pa[1] = pack_half2(sacc[2 * ks][2], sacc[2 * ks][3]);
        pa[2] = pack_half2(sacc[2 * ks + 1][0], sacc[2 * ks + 1][1]);
        pa[3] = pack_half2(sacc[2 * ks + 1][2], sacc[2 * ks + 1][3]);
#pragma unroll
        for (uint32_t dp = 0; dp < 2; ++dp) {
            uint32_t vb[4];
            ldmatrix_x4_trans(vb, vaddr + (ks * 16 * P + dp * 16) * 2);
            mma_m16n8k16(oacc[2 * dp], pa, vb[0], vb[1]);
            mma_m16n8k16(oacc[2 * dp + 1], pa, vb[2], vb[3]);
        }
    }
    const float inv0 = sum0 > 0.0f ? 1.0f / sum0 : 0.0f, inv1 = sum1 > 0.0f ? 1.0f / sum1 : 0.0f;
#pragma unroll
    for (uint32_t dt = 0; dt < 4; ++dt) {
        const uint32_t d = dt * 8 + tq * 2;
        if (r0 < L)
            *reinterpret_cast<uint32_t *>(ctx_bh + (size_t)r0 * H + d) = pack_half2(oacc[dt][0] * inv0, oacc[dt][1] * inv0);
        if (r0 + 8 < L)
            *reinterpret_cast<uint32_t *>(ctx_bh + (size_t)(r0 + 8) * H + d) = pack_half2(oacc[dt][2] * inv1, oacc[dt][3] * inv1);
    }
    }
}

// Q, K and V of one (sequence, head) sit ROW-MAJOR in shared memory ([row][32] halfs at a pitch of 40 halfs = 80 bytes:
// eight consecutive rows start in eight different 16-byte bank groups, so every ldmatrix phase is conflict-free).  The
// fragments come from ldmatrix: Q as the A operand, K rows as the (col-major) B operand of Q K^T, and V through
// ldmatrix.trans as the B operand of P V -- the first version transposed V with eight scalar 2-byte stores per thread.
// PERSISTENT: a CTA walks work items (sequence, head, 64-query block) and copies the NEXT item's Q/K/V/mask with
// cp.async into the other half of a double buffer while the tensor cores work on the current one -- a one-item CTA
// spends most of its life waiting for its own loads (profile: the top stall was the first shared-memory store).
template <uint32_t KS>   // KS = L16 / 16: 16-key steps (1..8)
__global__ void __launch_bounds__(128, KS <= 4 ? 7 : 3)
attention_mma_kernel(const __half *__restrict__ qkv, const int64_t *__restrict__ mask, uint32_t L, uint32_t H,
                     uint32_t heads, uint32_t n_items, uint32_t q_blocks, __half *__restrict__ ctx)
{
    constexpr uint32_t D = 32, L16 = KS * 16, NT = KS * 2, P = 40;
    constexpr uint32_t KIT = L16 / 32 + (L16 % 32 ? 1 : 0);
    struct Buf {
        __align__(16) __half Ks[L16 * P];
        __align__(16) __half Vs[L16 * P];
        __align__(16) __half Qs[64 * P];
        __align__(16) int64_t mk[L16];                  // raw attention mask of the sequence (0 beyond L)
    };
    extern __shared__ __align__(16) uint8_t att_smem_raw[];
    Buf *bufs = reinterpret_cast<Buf *>(att_smem_raw);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t g = lane >> 2, tq = lane & 3;       // fragment row group / quad lane
    const size_t row_stride = 3 * (size_t)H;
    const float NEG_INF = -__uint_as_float(0x7F800000u);
    const uint32_t part = threadIdx.x & 3, j0 = threadIdx.x >> 2;

    auto prefetch = [&](uint32_t item, Buf &bf) {
        const uint32_t qb = item % q_blocks, bh = item / q_blocks, b = bh / heads, h = bh % heads;
        const __half *base = qkv + (size_t)b * L * row_stride + h * D;
#pragma unroll
        for (uint32_t it = 0; it < KIT; ++it) {
            const uint32_t j = j0 + it * 32;
            if (j < L16) {
                const uint32_t jc = j < L ? j : 0;
                cp_async16(smem_u32(bf.Ks + (size_t)j * P + part * 8), base + jc * row_stride + H + part * 8, j < L);
                cp_async16(smem_u32(bf.Vs + (size_t)j * P + part * 8), base + jc * row_stride + 2 * H + part * 8, j < L);
            }
        }
#pragma unroll
        for (uint32_t it = 0; it < 2; ++it) {
            const uint32_t j = qb * 64 + j0 + it * 32, jc = j < L ? j : 0;
            cp_async16(smem_u32(bf.Qs + (size_t)(j0 + it * 32) * P + part * 8), base + (size_t)jc * row_stride + part * 8, j < L);
        }
        if (threadIdx.x < L16) {
            const uint32_t j = threadIdx.x, jc = j < L ? j : 0;
            cp_async8(smem_u32(bf.mk + j), mask + (size_t)b * L + jc, j < L);
        }
        cp_async_commit();
    };

    pdl_wait();                                        // qkv comes from the QKV GEMM
    pdl_trigger();
    uint32_t item = blockIdx.x, cur = 0;
    if (item < n_items) prefetch(item, bufs[0]);
    for (; item < n_items; item += gridDim.x, cur ^= 1) {
        const uint32_t next = item + gridDim.x;
        if (next < n_items) {
            prefetch(next, bufs[cur ^ 1]);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const Buf &bf = bufs[cur];
        const uint32_t qb = item % q_blocks, bh = item / q_blocks, b = bh / heads, h = bh % heads;
        const uint32_t q_row0 = qb * 64;
        const uint32_t r0 = q_row0 + warp * 16 + g;    // this thread's rows: r0 and r0 + 8
        if (q_row0 + warp * 16 < L)
            attend_rows<KS>(bf.Qs + (size_t)warp * 16 * P, bf.Ks, bf.Vs, bf.mk, lane, q_row0 + warp * 16, L,
                            ctx + (size_t)b * L * H + h * D, H);
        __syncthreads();                               // the next prefetch overwrites this buffer
    }
}

template <uint32_t KS>
cudaError_t launch_attention_mma(const __half *qkv, const int64_t *mask, uint32_t b, uint32_t L, uint32_t H, uint32_t heads,
                                 __half *ctx, cudaStream_t st)
{
    const uint32_t q_blocks = (L + 63) / 64, n_items = b * heads * q_blocks;
    constexpr uint32_t smem_per_cta = 2 * ((2 * KS * 16 + 64) * 80 + KS * 16 * 8);
    static DevAttr da;
    cudaError_t err = ensure_smem(attention_mma_kernel<KS>, da, smem_per_cta);
    if (err != cudaSuccess) return err;
    static std::atomic<int> per_sm_cached{0};                      // resident CTAs per SM (registers / shared memory), asked once
    if (per_sm_cached.load(std::memory_order_relaxed) == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, attention_mma_kernel<KS>, 128, smem_per_cta) != cudaSuccess || n < 1) n = 1;
        per_sm_cached.store(n, std::memory_order_relaxed);
    }
    const uint32_t per_sm = (uint32_t)per_sm_cached.load(std::memory_order_relaxed);
    const uint32_t grid = n_items < 148 * per_sm ? n_items : 148 * per_sm;
    return launch_pdl(attention_mma_kernel<KS>, dim3(grid), dim3(128), smem_per_cta, st, qkv, mask, L, H, heads, n_items, q_blocks, ctx);
}

// ---- small batches: QKV projection + attention of one (sequence, head) in ONE CTA -------------------------------
// The single-query forward is a chain of ~26 dependent launches at ~9 us each; this kernel removes one per layer.  A CTA
// computes its head's q, k, v columns for the whole sequence on the tensor core -- [128 rows x 384] . [384 x 96]^T, the
// 96 weight rows being three 32-row boxes of Wqkv (q, k, v of head h), all six K-blocks in flight at once (no ring to
// recycle; the weight boxes are requested BEFORE the programmatic-dependency wait, they do not depend on the previous
// kernel) -- drains the accumulator (+bias, -> fp16, the rounding the separate QKV GEMM applied) straight into the
// row-major shared-memory tiles the attention fragments are read from, and runs attend_rows on them.  L <= 128.
template <uint32_t KS>
__global__ void __launch_bounds__(256, 1)
qkv_attention_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                     const float *__restrict__ bqkv, const int64_t *__restrict__ mask, uint32_t L, uint32_t H, uint32_t heads,
                     __half *__restrict__ ctx)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t D = 32, P = 40, L16 = KS * 16, A_BYTES = 128 * 128, W_BOX = 32 * 128, STAGE = A_BYTES + 3 * W_BOX;   // 28 KB
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nkb = H / 64;                                   // 6: every K-block has its own stage
    const uint32_t b = blockIdx.x / heads, h = blockIdx.x % heads;
    __half *Qs = reinterpret_cast<__half *>(smem + (size_t)nkb * STAGE);        // [128][P]
    __half *Ks = Qs + 128 * P, *Vs = Ks + 128 * P;
    int64_t *mk = reinterpret_cast<int64_t *>(Vs + 128 * P);                    // [128]
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(mk + 128), *wfull_bar = full_bar + nkb, *done_bar = wfull_bar + nkb;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done_bar + 1);
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < nkb; ++i) {
            mbar_init(smem_u32(full_bar + i), 1);
            mbar_init(smem_u32(wfull_bar + i), 1);
        }
        mbar_init(smem_u32(done_bar), 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) {
        // weights first (constant), activations after the previous kernel has completed
        if (elect_one_sync()) {
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                const uint32_t sw = smem_u32(smem + (size_t)kb * STAGE) + A_BYTES;
                mbar_arrive_expect_tx(smem_u32(wfull_bar + kb), 3 * W_BOX);
#pragma unroll
                for (uint32_t m = 0; m < 3; ++m)
                    tma_load_2d(sw + m * W_BOX, &tmap_w, smem_u32(wfull_bar + kb), (int32_t)(kb * 64), (int32_t)(m * H + h * D));
            }
        }
        __syncwarp();
        pdl_wait();
        pdl_trigger();
        if (elect_one_sync()) {
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                mbar_arrive_expect_tx(smem_u32(full_bar + kb), A_BYTES);
                tma_load_2d(smem_u32(smem + (size_t)kb * STAGE), &tmap_a, smem_u32(full_bar + kb), (int32_t)(kb * 64), (int32_t)(b * L));
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_f16(128, 96);
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            mbar_wait(smem_u32(wfull_bar + kb), 0);
            mbar_wait(smem_u32(full_bar + kb), 0);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t sa = smem_u32(smem + (size_t)kb * STAGE);
                const uint64_t da = make_sw128_desc(sa), db = make_sw128_desc(sa + A_BYTES);
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k) umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                if (kb + 1 == nkb) umma_commit(smem_u32(done_bar));
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // thread t <-> sequence position t <-> TMEM lane t: q | k | v of the head = accumulator columns 0-31 | 32-63 | 64-95
        const uint32_t t = threadIdx.x - 128;
        const uint32_t lane_base = ((warp & 3) * 32) << 16;
        pdl_wait();                                                // (mask and bias are inputs / weights; harmless)
        mk[t] = (t < L) ? mask[(size_t)b * L + t] : 0;
        mbar_wait(smem_u32(done_bar), 0);
        tc_fence_after();
#pragma unroll 1
        for (uint32_t m = 0; m < 3; ++m) {
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_base + m * 32, r);
            tmem_ld_wait();
            const float4 *bias4 = reinterpret_cast<const float4 *>(bqkv + m * H + h * D);
            __align__(16) __half o[32];
#pragma unroll
            for (uint32_t q4 = 0; q4 < 8; ++q4) {
                const float4 bb = __ldg(bias4 + q4);
                o[4 * q4] = __float2half_rn(__uint_as_float(r[4 * q4]) + bb.x);
                o[4 * q4 + 1] = __float2half_rn(__uint_as_float(r[4 * q4 + 1]) + bb.y);
                o[4 * q4 + 2] = __float2half_rn(__uint_as_float(r[4 * q4 + 2]) + bb.z);
                o[4 * q4 + 3] = __float2half_rn(__uint_as_float(r[4 * q4 + 3]) + bb.w);
            }
            // rows at or beyond L belong to the next sequence / padding: they must read as zeros (keys are masked anyway)
            __half *dst = (m == 0 ? Qs : (m == 1 ? Ks : Vs)) + (size_t)t * P;
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q)
                reinterpret_cast<uint4 *>(dst)[q] = t < L ? reinterpret_cast<const uint4 *>(o)[q] : make_uint4(0, 0, 0, 0);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const uint32_t w = warp - 4;
        for (uint32_t q_row0 = 0; q_row0 < L; q_row0 += 64)
            if (q_row0 + w * 16 < L)
                attend_rows<KS>(Qs + (size_t)(q_row0 + w * 16) * P, Ks, Vs, mk, lane, q_row0 + w * 16, L,
                                ctx + (size_t)b * L * H + h * D, H);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

template <uint32_t KS>
cudaError_t launch_qkv_attention(const CUtensorMap &ta, const CUtensorMap &tw, const float *bqkv, const int64_t *mask, uint32_t b,
                                 uint32_t L, uint32_t H, uint32_t heads, __half *ctx, cudaStream_t st)
{
    const size_t smem = (size_t)(H / 64) * (128 * 128 + 3 * 32 * 128) + 3 * 128 * 40 * 2 + 128 * 8 + (2 * (H / 64) + 1) * 8 + 16;
    static DevAttr da;
    cudaError_t e = ensure_smem(qkv_attention_kernel<KS>, da, smem);
    if (e != cudaSuccess) return e;
    e = launch_pdl(qkv_attention_kernel<KS>, dim3(b * heads), dim3(256), smem, st, ta, tw, bqkv, mask, L, H, heads, ctx);
    count_launch();
    return e;
}

// ---- masked mean-pool + finalize (reference minilm.rs:959-978, :846-878) -----------------------------
// pooled[d] += h[s][d] for s ascending over mask==1; /= mask_sum; NaN/Inf -> 0; norm = sqrt(sum x^2)
// left-to-right; divide by norm if norm > f32::EPSILON.  One CTA per sequence, one thread per dim.
__global__ void pool_norm_kernel(const float *__restrict__ hidden, const int64_t *__restrict__ mask, uint32_t L,
                                 uint32_t H, float *__restrict__ out)
{
    extern __shared__ float pooled[];                  // [H] pooled, then [L] bytes: token is live
    __shared__ float s_norm;
    uint8_t *live = reinterpret_cast<uint8_t *>(pooled + H);
    const uint32_t b = blockIdx.x;
    for (uint32_t s = threadIdx.x; s < L; s += blockDim.x) live[s] = mask[(size_t)b * L + s] == 1;
    pdl_wait();                                        // the hidden states come from the last LayerNorm GEMM
    pdl_trigger();
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < H; d += blockDim.x) {
        // rows are added in token order, masked rows skipped -- the loads of 16 rows are issued together, the adds stay a chain
        const float *hp = hidden + (size_t)b * L * H + d;
        float acc = 0.0f;
        uint32_t cnt = 0, s = 0;
        for (; s + 16 <= L; s += 16) {
            float v[16];
#pragma unroll
            for (uint32_t j = 0; j < 16; ++j) v[j] = hp[(size_t)(s + j) * H];
#pragma unroll
            for (uint32_t j = 0; j < 16; ++j)
                if (live[s + j]) {
                    acc = __fadd_rn(acc, v[j]);
                    ++cnt;
                }
        }
        for (; s < L; ++s)
            if (live[s]) {
                acc = __fadd_rn(acc, hp[(size_t)s * H]);
                ++cnt;
            }
        // mask_sum is a sum of 1.0f: exact in f32 for any L <= 2^24
        if (cnt > 0) acc = __fdiv_rn(acc, (float)cnt);
        if (isnan(acc) || isinf(acc)) acc = 0.0f;
        pooled[d] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // left-to-right sum of squares (the reference's order): the loads and products run ahead in blocks of 8, only the
        // adds form the chain
        float ss = 0.0f;
        uint32_t d = 0;
        for (; d + 8 <= H; d += 8) {
            float p[8];
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) p[j] = __fmul_rn(pooled[d + j], pooled[d + j]);
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) ss = __fadd_rn(ss, p[j]);
        }
        for (; d < H; ++d) ss = __fadd_rn(ss, __fmul_rn(pooled[d], pooled[d]));
        s_norm = __fsqrt_rn(ss);
    }
    __syncthreads();
    const float norm = s_norm;
    for (uint32_t d = threadIdx.x; d < H; d += blockDim.x) {
        float v = pooled[d];
        if (norm > 1.1920929e-7f && !isnan(norm)) v = __fdiv_rn(v, norm);
        out[(size_t)b * H + d] = v;
    }
}

__global__ void f32_to_f16_rows_kernel(const float *__restrict__ src, __half *__restrict__ dst, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = __float2half_rn(src[i]);
}

int encode_tmap(CUtensorMap *out, const void *ptr, uint32_t k, uint64_t rows, uint32_t box_rows)
{
    return encode_corpus_tmap(out, ptr, true, k, rows, box_rows);
}

template <uint32_t BN, uint32_t EPI>
cudaError_t launch_gemm(const CUtensorMap &ta, const CUtensorMap &tw, const GemmArgs &g, uint32_t m_tiles, cudaStream_t st)
{
    constexpr uint32_t STAGE = 128 * 128 + BN * 128;
    const size_t smem = (size_t)g.stages * STAGE + (2 * g.stages + 1) * 8 + 48 + 3 * BN * 4 + 4 * 128 * 2 * 4;
    auto kern = gemm_tc_kernel<BN, EPI>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = launch_pdl(kern, dim3(m_tiles, g.N / BN), dim3(EPI == EPI_BIAS_RES_LN ? LN_THREADS : G_THREADS), smem, st, ta, tw, g);
    count_launch();
    return e;
}

template <uint32_t EPI>
cudaError_t launch_gemm_persist(const CUtensorMap &ta, const CUtensorMap &tw, const CUtensorMap &tout, const GemmArgs &g,
                                uint32_t m_tiles, uint32_t sm_count, cudaStream_t st)
{
    const uint32_t n_tiles = g.N / 128, total = m_tiles * n_tiles;
    const size_t smem = (size_t)g.stages * (128 * 128 + 128 * 128) + 2 * 128 * 128 + (2 * g.stages + 4) * 8 + 48 + 2 * 128 * 4;
    auto kern = gemm_tc_persist_kernel<EPI>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint32_t grid = total < 2 * sm_count ? total : 2 * sm_count;
    e = launch_pdl(kern, dim3(grid), dim3(P_THREADS), smem, st, ta, tw, tout, g, m_tiles, n_tiles);
    count_launch();
    return e;
}

template <uint32_t BN, uint32_t EPI>
cudaError_t launch_gemm_pair(const CUtensorMap &ta, const CUtensorMap &tw, const CUtensorMap &tout, const GemmArgs &g,
                             uint32_t m_tiles, uint32_t sm_count, cudaStream_t st)
{
    const uint32_t n_tiles = g.N / BN, m_pairs = (m_tiles + 1) / 2, total = m_pairs * n_tiles;
    const size_t smem = (size_t)g.stages * (128 * 128 + (BN / 2) * 128) + 2 * (BN / 64) * 128 * 128 + (2 * g.stages + 4) * 8 + 48 +
                        2 * BN * 4;
    auto kern = gemm_tc_pair_kernel<BN, EPI>;
    static DevAttr da;
    cudaError_t e = ensure_smem(kern, da, smem);
    if (e != cudaSuccess) return e;
    const uint32_t pairs = total < sm_count / 2 ? total : sm_count / 2;
    e = launch_pdl(kern, dim3(2 * pairs), dim3(LN_THREADS), smem, st, ta, tw, tout, g, m_pairs, n_tiles);
    count_launch();
    return e;
}

cudaError_t launch_gemm_ln_pair(const CUtensorMap &ta, const CUtensorMap &tw, const CUtensorMap &tout, const GemmArgs &g,
                                uint32_t m_tiles, cudaStream_t st)
{
    const size_t ring = (size_t)g.stages * (128 * 128 + 2 * 96 * 128);
    const size_t smem = (ring > 6 * 128 * 128 ? ring : 6 * 128 * 128) + (2 * g.stages + 1) * 8 + 48 + 3 * 384 * 4 + 4 * 128 * 2 * 4;
    static DevAttr da;
    cudaError_t e = ensure_smem(gemm_ln_pair_kernel, da, smem);
    if (e != cudaSuccess) return e;
    e = launch_pdl(gemm_ln_pair_kernel, dim3((m_tiles + 1) / 2 * 2), dim3(LN_THREADS), smem, st, ta, tw, tout, g);
    count_launch();
    return e;
}

cudaError_t launch_gemm_ln_cluster(const CUtensorMap &ta, const CUtensorMap &tw, const GemmArgs &g, uint32_t m_tiles,
                                   cudaStream_t st)
{
    const size_t smem = (size_t)g.stages * (128 * 128 + 128 * 128) + (2 * g.stages + 1) * 8 + 48 + 3 * 128 * 4 + 2 * 128 * 2 * 4;
    cudaError_t e = cudaFuncSetAttribute(gemm_ln_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = launch_pdl(gemm_ln_cluster_kernel, dim3(m_tiles * 3), dim3(P_THREADS), smem, st, ta, tw, g);    // __cluster_dims__(3,1,1)
    count_launch();
    return e;
}

struct Layer {
    __half *wqkv, *wo, *w1, *w2;          // [1152,H] [H,H] [F,H] [H,F]
    float *bqkv, *bo, *b1, *b2, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};

}  // namespace
}  // namespace sg

using namespace sg;

struct sg_encoder {
    int device = 0;
    uint32_t layers = 0, hidden = 0, heads = 0, ffn = 0, vocab = 0, max_pos = 0;
    float eps = 1e-12f;
    float *word = nullptr, *pos = nullptr, *type0 = nullptr, *emb_g = nullptr, *emb_b = nullptr;
    std::vector<Layer> L;
    std::vector<void *> allocs;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    // activations (grown on demand)
    size_t cap_tokens = 0;
    __half *h0 = nullptr, *h1 = nullptr, *qkv = nullptr, *ctx = nullptr, *mid = nullptr;
    float *hf32 = nullptr;
    int64_t *d_ids = nullptr, *d_mask = nullptr;
    float *d_emb = nullptr;
    size_t cap_b = 0;
    void *h_pinned = nullptr;
    size_t pinned_cap = 0;
    // CUDA graphs of whole forwards for small batches (the single-query recall path is ~30 tiny launches:
    // launch gaps, not kernels, set its latency).  Keyed by shape AND buffer addresses; dropped when buffers regrow.
    struct Graph {
        uint32_t b = 0, L = 0;
        const void *ids = nullptr, *mask = nullptr;
        void *out = nullptr;
        cudaGraphExec_t exec = nullptr;
        uint32_t launches = 0;
        bool failed = false;
        uint64_t last_use = 0;
    };
    std::vector<Graph> graphs;
    uint64_t graph_clock = 0;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;   // fence the caller's legacy default stream around a graph launch
};

namespace {

struct Blob {
    std::vector<uint8_t> data;
    struct T { std::string name; std::vector<uint32_t> shape; const float *ptr; size_t n; };
    std::vector<T> tensors;
    uint32_t hdr[6] = {0, 0, 0, 0, 0, 0};
    const T *find(const std::string &n) const
    {
        for (const auto &t : tensors)
            if (t.name == n) return &t;
        return nullptr;
    }
};

// "SGW1" | u32 version | u32 layers hidden heads ffn vocab max_pos | u32 n_tensors |
// per tensor: u32 name_len, name, u32 ndim, u32 dims[ndim], f32 data (little endian)
int load_blob(const char *path, Blob &b)
{
    FILE *f = fopen(path, "rb");
    if (!f) {
        set_error("cannot open weights file %s", path);
        return SG_ERR_INVALID;
    }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    b.data.resize((size_t)sz);
    size_t got = fread(b.data.data(), 1, (size_t)sz, f);
    fclose(f);
    if (got != (size_t)sz || sz < 40 || memcmp(b.data.data(), "SGW1", 4) != 0) {
        set_error("%s is not an SGW1 weights file", path);
        return SG_ERR_INVALID;
    }
    const uint8_t *p = b.data.data() + 4, *end = b.data.data() + sz;
    auto rd = [&](uint32_t &v) {
        if (p + 4 > end) return false;
        memcpy(&v, p, 4);
        p += 4;
        return true;
    };
    uint32_t version = 0, n = 0;
    if (!rd(version) || version != 1) {
        set_error("unsupported SGW1 version");
        return SG_ERR_INVALID;
    }
    for (int i = 0; i < 6; ++i)
        if (!rd(b.hdr[i])) return SG_ERR_INVALID;
    if (!rd(n)) return SG_ERR_INVALID;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t nl = 0, nd = 0;
        if (!rd(nl) || p + nl > end) return SG_ERR_INVALID;
        Blob::T t;
        t.name.assign(reinterpret_cast<const char *>(p), nl);
        p += nl;
        if (!rd(nd) || nd > 4) return SG_ERR_INVALID;
        size_t cnt = 1;
        for (uint32_t d = 0; d < nd; ++d) {
            uint32_t x = 0;
            if (!rd(x)) return SG_ERR_INVALID;
            t.shape.push_back(x);
            cnt *= x;
        }
        // data is 4-byte aligned relative to the file start by construction of the writer
        if (p + cnt * 4 > end) {
            set_error("truncated tensor %s", t.name.c_str());
            return SG_ERR_INVALID;
        }
        t.ptr = reinterpret_cast<const float *>(p);
        t.n = cnt;
        p += cnt * 4;
        b.tensors.push_back(t);
    }
    return SG_OK;
}

int upload_f32(sg_encoder *e, const Blob &b, const std::string &name, size_t expect, float **out)
{
    const Blob::T *t = b.find(name);
    if (!t || t->n != expect) {
        set_error("weights: tensor %s missing or of wrong size (want %zu elements)", name.c_str(), expect);
        return SG_ERR_INVALID;
    }
    std::vector<float> tmp(t->n);
    memcpy(tmp.data(), t->ptr, t->n * 4);   // the blob pointer may be unaligned
    SG_CUDA_TRY(cudaMalloc(out, t->n * 4));
    e->allocs.push_back(*out);
    SG_CUDA_TRY(cudaMemcpy(*out, tmp.data(), t->n * 4, cudaMemcpyHostToDevice));
    return SG_OK;
}

// concatenates the named [rows_i, cols] f32 tensors along rows and stores them as fp16
int upload_f16_cat(sg_encoder *e, const Blob &b, const std::vector<std::string> &names, size_t rows_each, size_t cols,
                   __half **out)
{
    std::vector<float> tmp;
    for (const auto &nm : names) {
        const Blob::T *t = b.find(nm);
        if (!t || t->n != rows_each * cols) {
            set_error("weights: tensor %s missing or of wrong size", nm.c_str());
            return SG_ERR_INVALID;
        }
        const size_t off = tmp.size();
        tmp.resize(off + t->n);
        memcpy(tmp.data() + off, t->ptr, t->n * 4);
    }
    float *d32 = nullptr;
    SG_CUDA_TRY(cudaMalloc(&d32, tmp.size() * 4));
    cudaError_t er = cudaMemcpy(d32, tmp.data(), tmp.size() * 4, cudaMemcpyHostToDevice);
    if (er == cudaSuccess) er = cudaMalloc(out, tmp.size() * 2);
    if (er == cudaSuccess) {
        e->allocs.push_back(*out);
        f32_to_f16_rows_kernel<<<256, 256>>>(d32, *out, tmp.size());
        count_launch();
        er = cudaDeviceSynchronize();
    }
    cudaFree(d32);
    if (er != cudaSuccess) {
        set_error("weight upload failed: %s", cudaGetErrorString(er));
        return SG_ERR_CUDA;
    }
    return SG_OK;
}

int upload_bias_cat(sg_encoder *e, const Blob &b, const std::vector<std::string> &names, size_t each, float **out)
{
    std::vector<float> tmp;
    for (const auto &nm : names) {
        const Blob::T *t = b.find(nm);
        if (!t || t->n != each) {
            set_error("weights: tensor %s missing or of wrong size", nm.c_str());
            return SG_ERR_INVALID;
        }
        const size_t off = tmp.size();
        tmp.resize(off + t->n);
        memcpy(tmp.data() + off, t->ptr, t->n * 4);
    }
    SG_CUDA_TRY(cudaMalloc(out, tmp.size() * 4));
    e->allocs.push_back(*out);
    SG_CUDA_TRY(cudaMemcpy(*out, tmp.data(), tmp.size() * 4, cudaMemcpyHostToDevice));
    return SG_OK;
}

template <typename T>
int regrow(T *&p, size_t elems)
{
    if (p) cudaFree(p);
    p = nullptr;
    SG_CUDA_TRY(cudaMalloc(&p, elems * sizeof(T)));
    return SG_OK;
}

int encode_device_locked(sg_encoder *e, const int64_t *d_ids, const int64_t *d_mask, uint32_t b, uint32_t L, float *d_out,
                         cudaStream_t st)
{
    const uint32_t H = e->hidden, F = e->ffn;
    const uint32_t T = b * L, T_pad = (T + 127) / 128 * 128, m_tiles = T_pad / 128;
    int rc;
    if (T_pad > e->cap_tokens) {
        const size_t c = (size_t)T_pad;
        if ((rc = regrow(e->h0, c * H)) != SG_OK) return rc;
        if ((rc = regrow(e->h1, c * H)) != SG_OK) return rc;
        if ((rc = regrow(e->qkv, c * 3 * H)) != SG_OK) return rc;
        if ((rc = regrow(e->ctx, c * H)) != SG_OK) return rc;
        if ((rc = regrow(e->mid, c * F)) != SG_OK) return rc;
        if ((rc = regrow(e->hf32, c * H)) != SG_OK) return rc;
        e->cap_tokens = c;
        // ctx rows beyond T are never written by attention: they must be finite (zero now; later calls leave finite values
        // of earlier, larger batches there -- padding rows are computed and discarded)
        SG_CUDA_TRY(cudaMemsetAsync(e->ctx, 0, c * H * 2, st));
    }
    SG_CUDA_TRY(launch_pdl(embed_ln_kernel, dim3((T_pad * 32 + 255) / 256), dim3(256), 0, st, d_ids, T, T_pad, L, H, e->vocab,
                           (const float *)e->word, (const float *)e->pos, (const float *)e->type0, (const float *)e->emb_g,
                           (const float *)e->emb_b, e->eps, e->h0));
    count_launch();
    CUtensorMap t_h0, t_h1, t_ctx, t_mid, t_qkv_out;
    if ((rc = encode_tmap(&t_h0, e->h0, H, T_pad, 128)) != SG_OK) return rc;
    if ((rc = encode_tmap(&t_h1, e->h1, H, T_pad, 128)) != SG_OK) return rc;
    if ((rc = encode_tmap(&t_ctx, e->ctx, H, T_pad, 128)) != SG_OK) return rc;
    if ((rc = encode_tmap(&t_mid, e->mid, F, T_pad, 128)) != SG_OK) return rc;          // load map of FFN-down == store map of FFN-up
    if ((rc = encode_tmap(&t_qkv_out, e->qkv, 3 * H, T_pad, 128)) != SG_OK) return rc;
    const size_t att_smem = ((size_t)32 * (L + 1) + (size_t)L * 32 + 4 * (size_t)L) * 4 + L;
    if (att_smem > 48 * 1024)
        SG_CUDA_TRY(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem));
    // 3-CTA cluster LayerNorm GEMM (weights streamed by 3 SMs) while 3 CTAs per row tile still fit one wave;
    // beyond that the single-CTA 128x384 tile re-reads the activations less (measured: 369 vs 473 us at 1 query,
    // 1114 vs 935 us at 256 queries)
    const bool ln_cluster = m_tiles * 3 <= 148 && !getenv("SG_ENC_NO_LN_CLUSTER");
    const bool persist = m_tiles * 9 >= 2 * 148 && !getenv("SG_ENC_NO_PERSIST");   // enough tiles to fill 2 CTAs per SM
    // timing experiments only (results are garbage): bit 0 QKV, 1 attention, 2 attention-out+LN, 3 FFN-up, 4 FFN-down+LN
    // CTA-pair 256x192 tiles once the batch fills the GPU (L2 traffic per FLOP 1.75x lower than 128x128 tiles)
    const bool pair = persist && (3 * H) % 192 == 0 && F % 192 == 0 && !getenv("SG_ENC_NO_PAIR");
    unsigned long long *d_trace = nullptr;             // SG_ENC_TRACE=1: phase stamps of layer 0's two LayerNorm GEMMs (debug)
    if (getenv("SG_ENC_TRACE") && !ln_cluster) {
        SG_CUDA_TRY(cudaMalloc(&d_trace, (size_t)2 * m_tiles * 8 * 8));
        SG_CUDA_TRY(cudaMemsetAsync(d_trace, 0, (size_t)2 * m_tiles * 8 * 8, st));
    }
    const bool ln_pair = !ln_cluster && H == 384 && !getenv("SG_ENC_NO_LN_PAIR");
    const char *skip_env = getenv("SG_ENC_SKIP");
    // one (sequence, head) per CTA while that is at most one wave (head dim 32, 12 heads: up to 12 sequences)
    const bool fused_qkv_attn = L <= 128 && H == e->heads * 32 && H % 64 == 0 && H <= 384 && b * e->heads <= 148 &&
                                !getenv("SG_ENC_NO_FUSED_QKV") && !getenv("SG_ENC_ATTN_SIMT");
    const int skip = skip_env ? atoi(skip_env) : 0;
    for (uint32_t l = 0; l < e->layers; ++l) {
        const Layer &ly = e->L[l];
        CUtensorMap t_wqkv, t_wo, t_w1, t_w2;
        if ((rc = encode_tmap(&t_wqkv, ly.wqkv, H, 3 * H, 128)) != SG_OK) return rc;
        if ((rc = encode_tmap(&t_wo, ly.wo, H, H, ln_cluster ? 128 : H / 2)) != SG_OK) return rc;
        if ((rc = encode_tmap(&t_w1, ly.w1, H, F, 128)) != SG_OK) return rc;
        if ((rc = encode_tmap(&t_w2, ly.w2, F, H, ln_cluster ? 128 : H / 2)) != SG_OK) return rc;
        GemmArgs g{};
        // QKV = h0 . Wqkv^T + b
        g.K = H; g.N = 3 * H; g.stages = 3; g.bias = ly.bqkv; g.out = e->qkv;   // 96 KB: two CTAs per SM
        if (fused_qkv_attn) {
            // small batches: QKV projection + attention per (sequence, head) in one launch
            CUtensorMap t_wqkv_head;
            if ((rc = encode_tmap(&t_wqkv_head, ly.wqkv, H, 3 * H, 32)) != SG_OK) return rc;
            switch ((L + 15) / 16) {
            case 1: SG_CUDA_TRY(launch_qkv_attention<1>(t_h0, t_wqkv_head, ly.bqkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            case 2: SG_CUDA_TRY(launch_qkv_attention<2>(t_h0, t_wqkv_head, ly.bqkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            case 3: SG_CUDA_TRY(launch_qkv_attention<3>(t_h0, t_wqkv_head, ly.bqkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            case 4: SG_CUDA_TRY(launch_qkv_attention<4>(t_h0, t_wqkv_head, ly.bqkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            case 5: SG_CUDA_TRY(launch_qkv_attention<5>(t_h0, t_wqkv_head, ly.bqkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            case 6: SG_CUDA_TRY(launch_qkv_attention<6>(t_h0, t_wqkv_head, ly.bqkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            case 7: SG_CUDA_TRY(launch_qkv_attention<7>(t_h0, t_wqkv_head, ly.bqkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            default: SG_CUDA_TRY(launch_qkv_attention<8>(t_h0, t_wqkv_head, ly.bqkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            }
        } else if (skip & 1) {
        } else if (pair) {
            CUtensorMap t_wqkv_half;
            if ((rc = encode_tmap(&t_wqkv_half, ly.wqkv, H, 3 * H, 96)) != SG_OK) return rc;
            g.stages = 4;
            SG_CUDA_TRY((launch_gemm_pair<192, EPI_BIAS>(t_h0, t_wqkv_half, t_qkv_out, g, m_tiles, 148, st)));
        } else if (persist) {
            g.stages = 2;
            SG_CUDA_TRY((launch_gemm_persist<EPI_BIAS>(t_h0, t_wqkv, t_qkv_out, g, m_tiles, 148, st)));
        }
        else SG_CUDA_TRY((launch_gemm<128, EPI_BIAS>(t_h0, t_wqkv, g, m_tiles, st)));
        if (fused_qkv_attn || (skip & 2)) {
        } else if (L <= 128 && !getenv("SG_ENC_ATTN_SIMT")) {
            switch ((L + 15) / 16) {
            case 1: SG_CUDA_TRY(launch_attention_mma<1>(e->qkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            case 2: SG_CUDA_TRY(launch_attention_mma<2>(e->qkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            case 3: SG_CUDA_TRY(launch_attention_mma<3>(e->qkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            case 4: SG_CUDA_TRY(launch_attention_mma<4>(e->qkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            case 5: SG_CUDA_TRY(launch_attention_mma<5>(e->qkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            case 6: SG_CUDA_TRY(launch_attention_mma<6>(e->qkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            case 7: SG_CUDA_TRY(launch_attention_mma<7>(e->qkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            default: SG_CUDA_TRY(launch_attention_mma<8>(e->qkv, d_mask, b, L, H, e->heads, e->ctx, st)); break;
            }
        } else if (L <= 128) {
            const uint32_t L4 = (L + 3) & ~3u;
            const size_t sm = ((size_t)2 * L4 * 33 + 64 * 33 + (size_t)64 * (L4 + 1)) * 4;
            if (sm > 48 * 1024)
                SG_CUDA_TRY(cudaFuncSetAttribute(attention_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            SG_CUDA_TRY(launch_pdl(attention_small_kernel, dim3(b * e->heads, (L + 63) / 64), dim3(128), sm, st, (const __half *)e->qkv,
                                   d_mask, L, H, e->heads, e->ctx));
        } else {
            SG_CUDA_TRY(launch_pdl(attention_kernel, dim3(b * e->heads), dim3(128), att_smem, st, (const __half *)e->qkv, d_mask, L, H,
                                   e->heads, e->ctx));
        }
        count_launch();
        // h1 = LN(ctx . Wo^T + b + h0)
        g = GemmArgs{};
        g.K = H; g.N = H; g.stages = 3; g.bias = ly.bo; g.residual = e->h0; g.gamma = ly.ln1_g; g.beta = ly.ln1_b;
        g.eps = e->eps; g.out = e->h1;
        if (skip & 4) {
        } else if (ln_cluster) {
            g.stages = 6;
            SG_CUDA_TRY(launch_gemm_ln_cluster(t_ctx, t_wo, g, m_tiles, st));
        } else if (ln_pair) {
            CUtensorMap t_wo_half;
            if ((rc = encode_tmap(&t_wo_half, ly.wo, H, H, 96)) != SG_OK) return rc;
            g.stages = 4; g.rows = T_pad;
            SG_CUDA_TRY(launch_gemm_ln_pair(t_ctx, t_wo_half, t_h1, g, m_tiles, st));
        } else {
            g.trace = l == 0 ? d_trace : nullptr;
            SG_CUDA_TRY((launch_gemm<384, EPI_BIAS_RES_LN>(t_ctx, t_wo, g, m_tiles, st)));
        }
        // mid = gelu(h1 . W1^T + b)
        g = GemmArgs{};
        g.K = H; g.N = F; g.stages = 3; g.bias = ly.b1; g.out = e->mid;
        if (skip & 8) {
        } else if (pair) {
            CUtensorMap t_w1_half;
            if ((rc = encode_tmap(&t_w1_half, ly.w1, H, F, 96)) != SG_OK) return rc;
            g.stages = 4;
            SG_CUDA_TRY((launch_gemm_pair<192, EPI_BIAS_GELU>(t_h1, t_w1_half, t_mid, g, m_tiles, 148, st)));
        } else if (persist) {
            g.stages = 2;
            SG_CUDA_TRY((launch_gemm_persist<EPI_BIAS_GELU>(t_h1, t_w1, t_mid, g, m_tiles, 148, st)));
        }
        else SG_CUDA_TRY((launch_gemm<128, EPI_BIAS_GELU>(t_h1, t_w1, g, m_tiles, st)));
        // h0 = LN(mid . W2^T + b + h1)
        g = GemmArgs{};
        g.K = F; g.N = H; g.stages = 3; g.bias = ly.b2; g.residual = e->h1; g.gamma = ly.ln2_g; g.beta = ly.ln2_b;
        g.eps = e->eps; g.out = e->h0; g.out_f32 = (l + 1 == e->layers) ? e->hf32 : nullptr;
        if (skip & 16) {
        } else if (ln_cluster) {
            g.stages = 6;
            SG_CUDA_TRY(launch_gemm_ln_cluster(t_mid, t_w2, g, m_tiles, st));
        } else if (ln_pair) {
            CUtensorMap t_w2_half;
            if ((rc = encode_tmap(&t_w2_half, ly.w2, F, H, 96)) != SG_OK) return rc;
            g.stages = 4; g.rows = T_pad;
            SG_CUDA_TRY(launch_gemm_ln_pair(t_mid, t_w2_half, t_h0, g, m_tiles, st));
        } else {
            g.trace = (l == 0 && d_trace) ? d_trace + (size_t)m_tiles * 8 : nullptr;
            SG_CUDA_TRY((launch_gemm<384, EPI_BIAS_RES_LN>(t_mid, t_w2, g, m_tiles, st)));
        }
    }
    SG_CUDA_TRY(launch_pdl(pool_norm_kernel, dim3(b), dim3(384), (size_t)H * 4 + L, st, (const float *)e->hf32, d_mask, L, H, d_out));
    count_launch();
    SG_CUDA_TRY(cudaGetLastError());
    if (d_trace) {
        std::vector<unsigned long long> h((size_t)2 * m_tiles * 8);
        SG_CUDA_TRY(cudaStreamSynchronize(st));
        SG_CUDA_TRY(cudaMemcpy(h.data(), d_trace, h.size() * 8, cudaMemcpyDeviceToHost));
        cudaFree(d_trace);
        static const char *names[7] = {"enter", "prologue done", "previous kernel complete", "accumulator complete", "pass 1 done",
                                       "pass 2 done", "all done"};
        for (int k = 0; k < 2; ++k) {
            const unsigned long long *t = h.data() + (size_t)k * m_tiles * 8;
            unsigned long long t0 = ~0ull;
            for (uint32_t c = 0; c < m_tiles; ++c)
                if (t[c * 8] && t[c * 8] < t0) t0 = t[c * 8];
            fprintf(stderr, "[sg enc trace] %s + LayerNorm, %u CTAs, us since the first CTA entered (min / max over CTAs):", k ? "FFN-down" : "attention-out", m_tiles);
            for (int i = 0; i < 7; ++i) {
                unsigned long long mn = ~0ull, mx = 0;
                for (uint32_t c = 0; c < m_tiles; ++c) {
                    const unsigned long long v = t[c * 8 + i];
                    if (v && v < mn) mn = v;
                    if (v > mx) mx = v;
                }
                fprintf(stderr, " %s %.1f/%.1f |", names[i], mn >= t0 && mn != ~0ull ? (mn - t0) / 1e3 : -1.0, mx >= t0 ? (mx - t0) / 1e3 : -1.0);
            }
            fprintf(stderr, "\n");
        }
    }
    return SG_OK;
}

// small batches: replay a captured graph of the whole forward (first call of a shape runs plainly and grows the
// buffers, the second captures, later ones replay)
int encode_device_graphed(sg_encoder *e, const int64_t *d_ids, const int64_t *d_mask, uint32_t b, uint32_t L, float *d_out,
                          cudaStream_t caller)
{
    static const bool enabled = [] {
        const char *v = getenv("SG_ENC_GRAPH");
        return !(v && *v && atoi(v) == 0);
    }();
    const uint32_t T_pad = (b * L + 127) / 128 * 128;
    if (!enabled || T_pad > 2048) return encode_device_locked(e, d_ids, d_mask, b, L, d_out, caller);
    // the legacy default stream cannot be captured: graphs for such callers live on the encoder's own stream,
    // fenced by two events so the call keeps its in-order meaning on the caller's stream
    const bool legacy = caller == nullptr || caller == cudaStreamLegacy || caller == cudaStreamPerThread;
    cudaStream_t st = legacy ? e->stream : caller;
    if (legacy && !e->ev_in) {
        SG_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming));
        SG_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming));
    }
    const size_t cap_before = e->cap_tokens;
    sg_encoder::Graph *hit = nullptr;
    for (auto &g : e->graphs)
        if (g.b == b && g.L == L && g.ids == d_ids && g.mask == d_mask && g.out == d_out) hit = &g;
    auto launch = [&](cudaGraphExec_t exec) -> int {
        if (legacy) {
            SG_CUDA_TRY(cudaEventRecord(e->ev_in, caller));
            SG_CUDA_TRY(cudaStreamWaitEvent(st, e->ev_in, 0));
        }
        SG_CUDA_TRY(cudaGraphLaunch(exec, st));
        if (legacy) {
            SG_CUDA_TRY(cudaEventRecord(e->ev_out, st));
            SG_CUDA_TRY(cudaStreamWaitEvent(caller, e->ev_out, 0));
        }
        return SG_OK;
    };
    if (hit && hit->exec) {
        hit->last_use = ++e->graph_clock;
        count_launch(hit->launches);
        return launch(hit->exec);
    }
    if (!hit || hit->failed || T_pad > e->cap_tokens) {
        const int rc = encode_device_locked(e, d_ids, d_mask, b, L, d_out, caller);
        if (e->cap_tokens != cap_before) {           // activations moved: every captured graph points at freed memory
            for (auto &g : e->graphs)
                if (g.exec) cudaGraphExecDestroy(g.exec);
            e->graphs.clear();
            hit = nullptr;
        }
        if (rc == SG_OK && !hit) {
            if (e->graphs.size() >= 8) {             // evict the least recently used
                size_t lru = 0;
                for (size_t i = 1; i < e->graphs.size(); ++i)
                    if (e->graphs[i].last_use < e->graphs[lru].last_use) lru = i;
                if (e->graphs[lru].exec) cudaGraphExecDestroy(e->graphs[lru].exec);
                e->graphs.erase(e->graphs.begin() + lru);
            }
            sg_encoder::Graph g;
            g.b = b; g.L = L; g.ids = d_ids; g.mask = d_mask; g.out = d_out;
            g.last_use = ++e->graph_clock;
            e->graphs.push_back(g);
        }
        return rc;
    }
    // second sighting of this (shape, buffers): capture
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
        hit->failed = true;                          // the caller is capturing itself: stay out of its way
        cudaGetLastError();
        return encode_device_locked(e, d_ids, d_mask, b, L, d_out, caller);
    }
    const uint64_t l0 = sg_kernel_launch_count();
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    int rc = SG_OK;
    if (ce == cudaSuccess) {
        rc = encode_device_locked(e, d_ids, d_mask, b, L, d_out, st);
        ce = cudaStreamEndCapture(st, &graph);
    }
    if (ce == cudaSuccess && rc == SG_OK && graph) ce = cudaGraphInstantiate(&hit->exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (ce != cudaSuccess || rc != SG_OK || !hit->exec) {
        cudaGetLastError();
        hit->exec = nullptr;
        hit->failed = true;
        return encode_device_locked(e, d_ids, d_mask, b, L, d_out, caller);
    }
    hit->launches = (uint32_t)(sg_kernel_launch_count() - l0);
    hit->last_use = ++e->graph_clock;
    return launch(hit->exec);
}

}  // namespace

extern "C" {

int sg_encoder_create(sg_encoder **out, int device, const char *weights_path)
{
    if (!out || !weights_path) return SG_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available; libshodh_gpu has no CPU fallback");
        return SG_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) return SG_ERR_INVALID;
    cudaDeviceProp prop;
    SG_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libshodh_gpu is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return SG_ERR_NO_DEVICE;
    }
    Blob blob;
    int rc = load_blob(weights_path, blob);
    if (rc != SG_OK) return rc;
    SG_CUDA_TRY(cudaSetDevice(device));
    sg_encoder *e = new (std::nothrow) sg_encoder();
    if (!e) return SG_ERR_OOM;
    e->device = device;
    e->layers = blob.hdr[0]; e->hidden = blob.hdr[1]; e->heads = blob.hdr[2];
    e->ffn = blob.hdr[3]; e->vocab = blob.hdr[4]; e->max_pos = blob.hdr[5];
    auto fail = [&](int code) {
        for (void *p : e->allocs) cudaFree(p);
        delete e;
        return code;
    };
    if (e->hidden != 384 || e->heads * 32 != e->hidden || e->ffn % 128 != 0 || e->layers == 0 || e->layers > 48) {
        set_error("unsupported encoder shape (hidden=%u heads=%u ffn=%u): the kernels cover hidden 384 / head dim 32", e->hidden,
                  e->heads, e->ffn);
        return fail(SG_ERR_UNSUPPORTED);
    }
    const size_t H = e->hidden, F = e->ffn;
    const std::string em = "embeddings.";
    if ((rc = upload_f32(e, blob, em + "word_embeddings.weight", (size_t)e->vocab * H, &e->word)) != SG_OK) return fail(rc);
    if ((rc = upload_f32(e, blob, em + "position_embeddings.weight", (size_t)e->max_pos * H, &e->pos)) != SG_OK) return fail(rc);
    {
        const Blob::T *t = blob.find(em + "token_type_embeddings.weight");
        if (!t || t->n < H) {
            set_error("weights: token_type_embeddings missing");
            return fail(SG_ERR_INVALID);
        }
        std::vector<float> tmp(H);
        memcpy(tmp.data(), t->ptr, H * 4);
        if (cudaMalloc(&e->type0, H * 4) != cudaSuccess) return fail(SG_ERR_OOM);
        e->allocs.push_back(e->type0);
        cudaMemcpy(e->type0, tmp.data(), H * 4, cudaMemcpyHostToDevice);
    }
    if ((rc = upload_f32(e, blob, em + "LayerNorm.weight", H, &e->emb_g)) != SG_OK) return fail(rc);
    if ((rc = upload_f32(e, blob, em + "LayerNorm.bias", H, &e->emb_b)) != SG_OK) return fail(rc);
    e->L.resize(e->layers);
    for (uint32_t l = 0; l < e->layers; ++l) {
        const std::string p = "encoder.layer." + std::to_string(l) + ".";
        Layer &ly = e->L[l];
        if ((rc = upload_f16_cat(e, blob, {p + "attention.self.query.weight", p + "attention.self.key.weight",
                                           p + "attention.self.value.weight"}, H, H, &ly.wqkv)) != SG_OK) return fail(rc);
        if ((rc = upload_bias_cat(e, blob, {p + "attention.self.query.bias", p + "attention.self.key.bias",
                                            p + "attention.self.value.bias"}, H, &ly.bqkv)) != SG_OK) return fail(rc);
        if ((rc = upload_f16_cat(e, blob, {p + "attention.output.dense.weight"}, H, H, &ly.wo)) != SG_OK) return fail(rc);
        if ((rc = upload_f32(e, blob, p + "attention.output.dense.bias", H, &ly.bo)) != SG_OK) return fail(rc);
        if ((rc = upload_f32(e, blob, p + "attention.output.LayerNorm.weight", H, &ly.ln1_g)) != SG_OK) return fail(rc);
        if ((rc = upload_f32(e, blob, p + "attention.output.LayerNorm.bias", H, &ly.ln1_b)) != SG_OK) return fail(rc);
        if ((rc = upload_f16_cat(e, blob, {p + "intermediate.dense.weight"}, F, H, &ly.w1)) != SG_OK) return fail(rc);
        if ((rc = upload_f32(e, blob, p + "intermediate.dense.bias", F, &ly.b1)) != SG_OK) return fail(rc);
        if ((rc = upload_f16_cat(e, blob, {p + "output.dense.weight"}, H, F, &ly.w2)) != SG_OK) return fail(rc);
        if ((rc = upload_f32(e, blob, p + "output.dense.bias", H, &ly.b2)) != SG_OK) return fail(rc);
        if ((rc = upload_f32(e, blob, p + "output.LayerNorm.weight", H, &ly.ln2_g)) != SG_OK) return fail(rc);
        if ((rc = upload_f32(e, blob, p + "output.LayerNorm.bias", H, &ly.ln2_b)) != SG_OK) return fail(rc);
    }
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(SG_ERR_CUDA);
    *out = e;
    return SG_OK;
}

void sg_encoder_destroy(sg_encoder *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (auto &g : e->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (e->ev_in) cudaEventDestroy(e->ev_in);
    if (e->ev_out) cudaEventDestroy(e->ev_out);
    for (void *p : e->allocs) cudaFree(p);
    cudaFree(e->h0); cudaFree(e->h1); cudaFree(e->qkv); cudaFree(e->ctx); cudaFree(e->mid); cudaFree(e->hf32);
    cudaFree(e->d_ids); cudaFree(e->d_mask); cudaFree(e->d_emb);
    if (e->h_pinned) cudaFreeHost(e->h_pinned);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

// The tail of generate_embedding_onnx on its own (minilm.rs:959-978 masked mean-pool, :846-878 finalize_pooled):
// f32 hidden states in, pooled + L2-normalised embeddings out, in the reference's summation order (pool_norm_kernel).
int sg_pool_finalize(int device, const float *hidden, const int64_t *mask, uint32_t b, uint32_t L, uint32_t H, float *out)
{
    if (b == 0) return SG_OK;
    if (!hidden || !mask || !out || L == 0 || H == 0 || H > 8192 || L > 8192) return SG_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available; libshodh_gpu has no CPU fallback");
        return SG_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) return SG_ERR_INVALID;
    SG_CUDA_TRY(cudaSetDevice(device));
    float *d_h = nullptr, *d_o = nullptr;
    int64_t *d_m = nullptr;
    const size_t hb = (size_t)b * L * H * 4, mb = (size_t)b * L * 8, ob = (size_t)b * H * 4;
    cudaError_t e = cudaMalloc(&d_h, hb);
    if (e == cudaSuccess) e = cudaMalloc(&d_m, mb);
    if (e == cudaSuccess) e = cudaMalloc(&d_o, ob);
    if (e == cudaSuccess) e = cudaMemcpy(d_h, hidden, hb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_m, mask, mb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        pool_norm_kernel<<<b, H < 384 ? (H + 31) / 32 * 32 : 384, (size_t)H * 4 + L>>>(d_h, d_m, L, H, d_o);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d_o, ob, cudaMemcpyDeviceToHost);
    cudaFree(d_h); cudaFree(d_m); cudaFree(d_o);
    if (e != cudaSuccess) {
        set_error("sg_pool_finalize: %s", cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? SG_ERR_OOM : SG_ERR_CUDA;
    }
    return SG_OK;
}

int sg_encoder_dim(const sg_encoder *e, uint32_t *dim_out)
{
    if (!e || !dim_out) return SG_ERR_INVALID;
    *dim_out = e->hidden;
    return SG_OK;
}

int sg_encoder_encode_device(sg_encoder *e, const int64_t *d_ids, const int64_t *d_mask, uint32_t b, uint32_t L,
                             float *d_emb_out, void *cuda_stream)
{
    if (!e || (b && (!d_ids || !d_mask || !d_emb_out))) return SG_ERR_INVALID;
    if (b == 0) return SG_OK;
    if (L == 0 || L > e->max_pos || L > 512) {
        set_error("sequence length %u outside 1..%u", L, e->max_pos < 512 ? e->max_pos : 512);
        return SG_ERR_INVALID;
    }
    std::lock_guard<std::mutex> g(e->mu);
    SG_CUDA_TRY(cudaSetDevice(e->device));
    return encode_device_graphed(e, d_ids, d_mask, b, L, d_emb_out, static_cast<cudaStream_t>(cuda_stream));
}

int sg_encoder_encode(sg_encoder *e, const int64_t *ids, const int64_t *mask, uint32_t b, uint32_t L, float *emb_out)
{
    if (!e || (b && (!ids || !mask || !emb_out))) return SG_ERR_INVALID;
    if (b == 0) return SG_OK;
    if (L == 0 || L > e->max_pos || L > 512) {
        set_error("sequence length %u outside 1..%u", L, e->max_pos < 512 ? e->max_pos : 512);
        return SG_ERR_INVALID;
    }
    std::lock_guard<std::mutex> g(e->mu);
    SG_CUDA_TRY(cudaSetDevice(e->device));
    const size_t n = (size_t)b * L, ob = (size_t)b * e->hidden * 4;
    if (n > e->cap_b) {
        int rc;
        if ((rc = regrow(e->d_ids, n)) != SG_OK) return rc;
        if ((rc = regrow(e->d_mask, n)) != SG_OK) return rc;
        if ((rc = regrow(e->d_emb, n * e->hidden)) != SG_OK) return rc;   // b <= n
        e->cap_b = n;
    }
    const size_t need = 2 * n * 8 + ob;
    if (need > e->pinned_cap) {
        if (e->h_pinned) cudaFreeHost(e->h_pinned);
        e->h_pinned = nullptr;
        e->pinned_cap = 0;
        SG_CUDA_TRY(cudaHostAlloc(&e->h_pinned, need, cudaHostAllocDefault));
        e->pinned_cap = need;
    }
    uint8_t *hp = static_cast<uint8_t *>(e->h_pinned);
    memcpy(hp, ids, n * 8);
    memcpy(hp + n * 8, mask, n * 8);
    SG_CUDA_TRY(cudaMemcpyAsync(e->d_ids, hp, n * 8, cudaMemcpyHostToDevice, e->stream));
    SG_CUDA_TRY(cudaMemcpyAsync(e->d_mask, hp + n * 8, n * 8, cudaMemcpyHostToDevice, e->stream));
    int rc = encode_device_graphed(e, e->d_ids, e->d_mask, b, L, e->d_emb, e->stream);
    if (rc != SG_OK) return rc;
    SG_CUDA_TRY(cudaMemcpyAsync(hp + 2 * n * 8, e->d_emb, ob, cudaMemcpyDeviceToHost, e->stream));
    SG_CUDA_TRY(cudaStreamSynchronize(e->stream));
    memcpy(emb_out, hp + 2 * n * 8, ob);
    return SG_OK;
}

}  // extern "C"
