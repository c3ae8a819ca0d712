// Row-wise CUDA-core kernels of the PDA block: token assembly, LayerNorm, residuals,
// ReLU, neighbourhood max-pool — each fused with the hi/lo split that feeds the error-compensated
// 3xTF32 tensor-core projections (pda_block.py).  They replace chains of PyTorch elementwise / cat / copy /
// layer_norm launches that each made a full pass over the (tokens x channels) activations:
//
//   pdab_pda_assemble_ln_split : cat[pos, feat*scale, feat, glob] -> LayerNorm -> (hi, lo)      (1 read pass, 2 writes)
// (the stand-alone residual + LayerNorm, ReLU and max-pool kernels of the first version were removed once those steps had
// moved into the epilogues of the tensor-core kernels; this one survives as the un-fused-encoder path of pda_block.py)
//
// hi keeps the top 19 bits of the fp32 value (exactly representable in TF32), lo = value - hi (exact in fp32),
// so hi + lo reproduces the value bit for bit and the un-split tensor is never stored.
// One warp per token row; a lane holds E/32 (= 8 or 16) consecutive channels as float4.  All traffic is
// 16-byte vectorised and coalesced; the kernels are HBM-bound (bytes per token in DESIGN.md).
// LayerNorm: two-pass mean / variance in fp32 over the row (biased variance, eps inside the sqrt), like
// torch.nn.LayerNorm (PB/PointFormer.py:17-18,29,34).
#include "common.cuh"

namespace {

constexpr int kRowThreads = 256;  // 8 rows per CTA

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

__device__ __forceinline__ void split_store(float4 v, float4 *hi, float4 *lo) {
    if (lo == nullptr) {  // un-split output (the tcgen05 GEMMs split their A operand themselves); it is re-read soon
        *hi = v;
        return;
    }
    float4 h;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    __stcs(hi, h);
    __stcs(lo, make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
}

// V4: float4 per lane (E = 128 * V4)
template <int V4>
__device__ __forceinline__ void layer_norm_split(float4 (&v)[V4], int lane, const float *__restrict__ gamma,
                                                 const float *__restrict__ beta, float eps, float *hi_row,
                                                 float *lo_row) {
    constexpr int E = 128 * V4;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V4; i++) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / E);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V4; i++) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / E) + eps);
#pragma unroll
    for (int i = 0; i < V4; i++) {
        const int col = (i * 32 + lane) * 4;
        const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma + col));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(beta + col));
        float4 y;
        y.x = (v[i].x - mean) * rstd * g.x + b.x;
        y.y = (v[i].y - mean) * rstd * g.y + b.y;
        y.z = (v[i].z - mean) * rstd * g.z + b.z;
        y.w = (v[i].w - mean) * rstd * g.w + b.w;
        split_store(y, reinterpret_cast<float4 *>(hi_row + col),
                    lo_row ? reinterpret_cast<float4 *>(lo_row + col) : nullptr);
    }
}

// tokens row = [pos (C) | feat * scale (C) | feat (C) | glob[group] (C)],  E = 4C
template <int V4>
__global__ void __launch_bounds__(kRowThreads)
assemble_ln_split_kernel(long long T, int ns, int C, int xpitch, const float *__restrict__ pos,
                         const float *__restrict__ X, const float *__restrict__ scale,
                         const float *__restrict__ glob, const float *__restrict__ gamma,
                         const float *__restrict__ beta, float eps, float *__restrict__ hi, float *__restrict__ lo) {
    constexpr int E = 128 * V4;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
    if (row >= T) return;
    const long long g = row / ns;
    const float sc = __ldg(scale + row);
    float4 v[V4];
#pragma unroll
    for (int i = 0; i < V4; i++) {
        const int col = (i * 32 + lane) * 4;  // 4 consecutive channels never straddle a quarter (C % 4 == 0)
        const int part = col / C, cc = col - part * C;
        if (part == 0) {
            v[i] = __ldg(reinterpret_cast<const float4 *>(pos + row * C + cc));
        } else if (part == 3) {
            v[i] = __ldg(reinterpret_cast<const float4 *>(glob + g * C + cc));
        } else {
            v[i] = __ldg(reinterpret_cast<const float4 *>(X + row * xpitch + 8 + cc));
            if (part == 1) {
                v[i].x *= sc;
                v[i].y *= sc;
                v[i].z *= sc;
                v[i].w *= sc;
            }
        }
    }
    layer_norm_split<V4>(v, lane, gamma, beta, eps, hi + row * E, lo ? lo + row * E : nullptr);
}

int row_grid(long long T, dim3 &grid) {
    const long long blocks = (T + kRowThreads / 32 - 1) / (kRowThreads / 32);
    if (blocks > 2147483647LL) return PDAB_EUNSUPPORTED;
    grid = dim3((unsigned)blocks);
    return 0;
}

}  // namespace

extern "C" int pdab_pda_assemble_ln_split(long long tokens, int nsample, int c, int xpitch, const float *pos,
                                          const float *x, const float *scale, const float *glob,
                                          const float *gamma, const float *beta, float eps, float *hi, float *lo,
                                          pdab_stream_t stream) {
    if (tokens < 0 || nsample < 1 || c < 1 || !pos || !x || !scale || !glob || !gamma || !beta || !hi)
        return PDAB_EINVAL;
    if (tokens == 0) return 0;
    if ((c & 3) || xpitch < 8 + c || (xpitch & 3) || tokens % nsample) return PDAB_EINVAL;
    dim3 grid;
    if (int rc = row_grid(tokens, grid)) return rc;
    cudaStream_t s = pdab::to_stream(stream);
    const int E = 4 * c;
    if (E == 256)
        assemble_ln_split_kernel<2><<<grid, kRowThreads, 0, s>>>(tokens, nsample, c, xpitch, pos, x, scale, glob, gamma,
                                                                  beta, eps, hi, lo);
    else if (E == 512)
        assemble_ln_split_kernel<4><<<grid, kRowThreads, 0, s>>>(tokens, nsample, c, xpitch, pos, x, scale, glob, gamma,
                                                                  beta, eps, hi, lo);
    else
        return PDAB_EUNSUPPORTED;
    PDAB_LAUNCH_CHECK();
    return 0;
}
