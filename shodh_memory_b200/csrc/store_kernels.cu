// store_kernels.cu -- small data-movement kernels of the vector store: dtype
// conversion on append / extract, soft-delete bitmap update, and the deterministic
// synthetic-corpus generator used by tests and bench.py.
#include "sg_internal.h"

namespace sg {
namespace {

__global__ void f32_to_f16_kernel(const float4 *__restrict__ src, uint2 *__restrict__ dst, uint64_t n4)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        float4 v = src[i];
        __half2 lo = __floats2half2_rn(v.x, v.y);  // round-to-nearest-even, == (_Float16) cast
        __half2 hi = __floats2half2_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t *>(&lo);
        o.y = *reinterpret_cast<uint32_t *>(&hi);
        dst[i] = o;
    }
}

__global__ void f16_to_f32_kernel(const __half *__restrict__ src, float *__restrict__ dst, uint64_t n)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = __half2float(src[i]);
}

__global__ void set_bit_kernel(uint32_t *bitmap, uint32_t id, uint32_t *was_set)
{
    uint32_t old = atomicOr(bitmap + (id >> 5), 1u << (id & 31));
    *was_set = (old >> (id & 31)) & 1u;
}

// ---- synthetic corpus: mirrors oracle/shodh_oracle.c synth_normal / synth_row exactly ----
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float synth_normal(uint64_t seed_stream, uint64_t index)
{
    uint64_t h = mix64(seed_stream + index);
    int32_t s = (int32_t)(h & 0xFFFF) + (int32_t)((h >> 16) & 0xFFFF) + (int32_t)((h >> 32) & 0xFFFF) +
                (int32_t)((h >> 48) & 0xFFFF) - 131070;
    return __fmul_rn((float)s, 1.0f / 65536.0f);
}

// one thread per row (the sum of squares is a sequential left-to-right chain, like the
// oracle).  A one-off fill, not a hot path.
template <bool F16>
__global__ void __launch_bounds__(128)
synth_rows_kernel(void *rows, uint64_t seed, uint64_t first_row, uint64_t n, uint32_t dim)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint64_t row = first_row + r;
    const uint64_t ss_seed = mix64(seed ^ (1ull * 0xD1B54A32D192ED03ull));
    const uint64_t base = row * (uint64_t)dim;
    float ss = 0.0f;
    for (uint32_t j = 0; j < dim; ++j) {
        float x = synth_normal(ss_seed, base + j);
        ss = __fadd_rn(ss, __fmul_rn(x, x));
    }
    const float norm = __fsqrt_rn(ss);
    // second pass regenerates the samples and stores 8 elements (16/32 bytes) at a time
    for (uint32_t j = 0; j < dim; j += 8) {
        float x[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            x[t] = synth_normal(ss_seed, base + j + t);
            if (norm > 0.0f) x[t] = __fdiv_rn(x[t], norm);
        }
        if constexpr (F16) {
            __half2 h0 = __floats2half2_rn(x[0], x[1]), h1 = __floats2half2_rn(x[2], x[3]);
            __half2 h2 = __floats2half2_rn(x[4], x[5]), h3 = __floats2half2_rn(x[6], x[7]);
            uint4 o;
            o.x = *reinterpret_cast<uint32_t *>(&h0);
            o.y = *reinterpret_cast<uint32_t *>(&h1);
            o.z = *reinterpret_cast<uint32_t *>(&h2);
            o.w = *reinterpret_cast<uint32_t *>(&h3);
            *reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(rows) + r * dim + j) = o;
        } else {
            float4 *o = reinterpret_cast<float4 *>(reinterpret_cast<float *>(rows) + r * dim + j);
            o[0] = make_float4(x[0], x[1], x[2], x[3]);
            o[1] = make_float4(x[4], x[5], x[6], x[7]);
        }
    }
}

// one warp per row; the result only feeds an error BOUND, so it is rounded up, never down
__global__ void max_row_norm_kernel(const void *rows, bool f16, uint64_t n, uint32_t dim, float *out)
{
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (warp >= n) return;
    float ss = 0.0f;
    for (uint32_t j = lane; j < dim; j += 32) {
        float x = f16 ? __half2float(reinterpret_cast<const __half *>(rows)[warp * dim + j])
                      : reinterpret_cast<const float *>(rows)[warp * dim + j];
        ss = fmaf(x, x, ss);
    }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_down_sync(0xFFFFFFFFu, ss, o);
    if (lane == 0) {
        float nrm = sqrtf(ss) * 1.0001f;
        if (!(nrm == nrm)) nrm = __uint_as_float(0x7F800000u);   // NaN rows poison the bound -> +inf
        atomicMax(reinterpret_cast<unsigned int *>(out), __float_as_uint(nrm));  // non-negative floats order as uints
    }
}

}  // namespace

cudaError_t launch_max_row_norm(const void *d_rows, bool f16, uint64_t n, uint32_t dim, float *d_max_norm,
                                cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const uint64_t threads = n * 32;
    max_row_norm_kernel<<<(uint32_t)((threads + 255) / 256), 256, 0, st>>>(d_rows, f16, n, dim, d_max_norm);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_f32_to_f16(const float *d_src, __half *d_dst, uint64_t n_elems, cudaStream_t st)
{
    if (n_elems == 0) return cudaSuccess;
    uint64_t n4 = n_elems / 4;  // dim is a multiple of 32, so n_elems % 4 == 0
    uint32_t blocks = (uint32_t)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    f32_to_f16_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(d_src),
                                              reinterpret_cast<uint2 *>(d_dst), n4);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_f16_to_f32(const __half *d_src, float *d_dst, uint64_t n_elems, cudaStream_t st)
{
    if (n_elems == 0) return cudaSuccess;
    uint32_t blocks = (uint32_t)((n_elems + 255) / 256 < 148 * 16 ? (n_elems + 255) / 256 : 148 * 16);
    f16_to_f32_kernel<<<blocks, 256, 0, st>>>(d_src, d_dst, n_elems);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_synth_rows(void *d_rows, bool f16, uint64_t seed, uint64_t first_row, uint64_t n,
                              uint32_t dim, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    uint32_t blocks = (uint32_t)((n + 127) / 128);
    if (f16)
        synth_rows_kernel<true><<<blocks, 128, 0, st>>>(d_rows, seed, first_row, n, dim);
    else
        synth_rows_kernel<false><<<blocks, 128, 0, st>>>(d_rows, seed, first_row, n, dim);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_set_bit(uint32_t *d_bitmap, uint32_t id, uint32_t *d_was_set, cudaStream_t st)
{
    set_bit_kernel<<<1, 1, 0, st>>>(d_bitmap, id, d_was_set);
    count_launch();
    return cudaGetLastError();
}

}  // namespace sg
