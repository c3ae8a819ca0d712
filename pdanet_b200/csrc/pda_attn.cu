// Neighbourhood self-attention of the PDA block's pre-norm transformer, one kernel.
//
// replaces: the attention core of nn.MultiheadAttention inside TransformerEncoderLayerPreNorm
//           (PB/PointFormer.py:30, PB/pointnet2_modules.py:929): q/k/v slicing + permute copies, q scaling,
//           bmm(q, k^T), softmax, bmm(p, v) and the permute back — ~12 launches and 6 passes over the tokens.
//
// Layout.  qkv (T, 3E) row-major as the in_proj GEMM leaves it: row t = [q (E) | k (E) | v (E)], head h owns columns
// [h*hd, (h+1)*hd) of each part.  Tokens of one neighbourhood are the NS consecutive rows g*NS .. g*NS+NS-1; attention
// never crosses neighbourhoods (sequence length NS = nsample, batch = centres).  ctx (T, E): row t = concatenated heads,
// i.e. exactly the operand of out_proj.
//
// One CTA (128 threads) per (neighbourhood, head): Q, K, V tiles (NS x HD each) staged in shared memory (row pitch
// HD+4 floats: 16-byte aligned, conflict-free for the access patterns below), scores in registers.
//   S = (Q / sqrt(hd)) K^T : thread (i, jb) owns query row i and KPT = NS/TPR keys; float4 steps along hd, the TPR threads
//                            of a row read the same q (broadcast), 8 rows per warp hit 8 different bank groups.
//   softmax               : max / sum over the TPR lanes of a row by shuffles; exp in full precision (expf) so the
//                            result tracks torch.softmax to ~1e-7.
//   O = P V               : P goes through shared memory once; thread (i, db) owns HD/TPR output columns, interleaved
//                            in float4 units so a row's TPR lanes write 16*TPR contiguous bytes per instruction.
// IEEE fp32 on the CUDA cores throughout (north_star: the distribution-aware encoding stays on CUDA cores; the
// contractions here are 16x16 / 32x32 — far below a tcgen05 tile).  FLOPs 4*NS*NS*HD per CTA; FMA:LDS ~ 3.5:1.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;

template <int NS, int HD>
struct AttnSmem {
    static constexpr int PITCH = HD + 4;
    static constexpr int PP = NS + 1;
    static constexpr int FLOATS = 3 * NS * PITCH + NS * PP;
    static constexpr int BYTES = FLOATS * 4;
};

template <int NS, int HD>
__global__ void __launch_bounds__(kThreads) group_attention_kernel(long long groups, int heads, const float *__restrict__ qkv,
                                                                   float *__restrict__ ctx) {
    using SM = AttnSmem<NS, HD>;
    constexpr int PITCH = SM::PITCH, PP = SM::PP;
    constexpr int TPR = kThreads / NS;      // threads per query row: 4 (NS=32) or 8 (NS=16)
    constexpr int KPT = NS / TPR;           // keys per thread: 8 or 2
    constexpr int DPT4 = HD / 4 / TPR;      // float4 output columns per thread
    extern __shared__ __align__(16) float sm[];
    float *sQ = sm, *sK = sQ + NS * PITCH, *sV = sK + NS * PITCH, *sP = sV + NS * PITCH;

    const long long gh = blockIdx.x;
    const long long g = gh / heads;
    const int h = (int)(gh - g * heads);
    const int E = heads * HD;
    const float *base = qkv + g * NS * 3LL * E + h * HD;
    const int tid = threadIdx.x;

    // stage Q (pre-scaled), K, V: rows of HD contiguous floats, float4 coalesced
    constexpr float scaling = HD == 64 ? 0.125f : 0.08838834764831845f;   // head_dim ** -0.5 (PyTorch scales q)
    constexpr int F4_PER_ROW = HD / 4;
    for (int i = tid; i < 3 * NS * F4_PER_ROW; i += kThreads) {
        const int part = i / (NS * F4_PER_ROW);
        const int rem = i - part * (NS * F4_PER_ROW);
        const int r = rem / F4_PER_ROW, c4 = rem - r * F4_PER_ROW;
        float4 v = __ldg(reinterpret_cast<const float4 *>(base + (long long)r * 3 * E + part * E) + c4);
        if (part == 0) {
            v.x *= scaling;
            v.y *= scaling;
            v.z *= scaling;
            v.w *= scaling;
        }
        *reinterpret_cast<float4 *>(sm + part * NS * PITCH + r * PITCH + c4 * 4) = v;
    }
    __syncthreads();

    const int i = tid / TPR;     // query row
    const int jb = tid % TPR;    // key block / output column block

    // ---- scores
    float s[KPT];
#pragma unroll
    for (int jj = 0; jj < KPT; jj++) s[jj] = 0.f;
    const float *qrow = sQ + i * PITCH;
#pragma unroll 4
    for (int d = 0; d < HD; d += 4) {
        const float4 q4 = *reinterpret_cast<const float4 *>(qrow + d);
#pragma unroll
        for (int jj = 0; jj < KPT; jj++) {
            const float4 k4 = *reinterpret_cast<const float4 *>(sK + (jb + jj * TPR) * PITCH + d);
            s[jj] = fmaf(q4.x, k4.x, s[jj]);
            s[jj] = fmaf(q4.y, k4.y, s[jj]);
            s[jj] = fmaf(q4.z, k4.z, s[jj]);
            s[jj] = fmaf(q4.w, k4.w, s[jj]);
        }
    }
    // ---- softmax over the NS keys of row i (the TPR lanes of a row are adjacent lanes of one warp)
    float mx = s[0];
#pragma unroll
    for (int jj = 1; jj < KPT; jj++) mx = fmaxf(mx, s[jj]);
#pragma unroll
    for (int off = 1; off < TPR; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < KPT; jj++) {
        s[jj] = expf(s[jj] - mx);
        sum += s[jj];
    }
#pragma unroll
    for (int off = 1; off < TPR; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int jj = 0; jj < KPT; jj++) sP[i * PP + jb + jj * TPR] = s[jj] * inv;
    __syncthreads();

    // ---- O = P V
    float4 o[DPT4];
#pragma unroll
    for (int e = 0; e < DPT4; e++) o[e] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int j = 0; j < NS; j++) {
        const float pj = sP[i * PP + j];
        const float *vrow = sV + j * PITCH;
#pragma unroll
        for (int e = 0; e < DPT4; e++) {
            const float4 v4 = *reinterpret_cast<const float4 *>(vrow + (e * TPR + jb) * 4);
            o[e].x = fmaf(pj, v4.x, o[e].x);
            o[e].y = fmaf(pj, v4.y, o[e].y);
            o[e].z = fmaf(pj, v4.z, o[e].z);
            o[e].w = fmaf(pj, v4.w, o[e].w);
        }
    }
    float *orow = ctx + (g * NS + i) * (long long)E + h * HD;
#pragma unroll
    for (int e = 0; e < DPT4; e++) *reinterpret_cast<float4 *>(orow + (e * TPR + jb) * 4) = o[e];
}

template <int NS, int HD>
int launch(long long groups, int heads, const float *qkv, float *ctx, cudaStream_t s) {
    using SM = AttnSmem<NS, HD>;
    auto kern = group_attention_kernel<NS, HD>;
    static bool configured = false;
    if (!configured) {
        PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::BYTES));
        configured = true;
    }
    const long long blocks = groups * heads;
    if (blocks > 2147483647LL) return PDAB_EUNSUPPORTED;
    kern<<<(unsigned)blocks, kThreads, SM::BYTES, s>>>(groups, heads, qkv, ctx);
    PDAB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int pdab_group_attention(long long groups, int nsample, int heads, int head_dim, const float *qkv,
                                    float *ctx, pdab_stream_t stream) {
    if (groups < 0 || heads < 1 || !qkv || !ctx) return PDAB_EINVAL;
    if (groups == 0) return 0;
    cudaStream_t s = pdab::to_stream(stream);
    if (nsample == 16 && head_dim == 64) return launch<16, 64>(groups, heads, qkv, ctx, s);
    if (nsample == 32 && head_dim == 64) return launch<32, 64>(groups, heads, qkv, ctx, s);
    if (nsample == 16 && head_dim == 128) return launch<16, 128>(groups, heads, qkv, ctx, s);
    if (nsample == 32 && head_dim == 128) return launch<32, 128>(groups, heads, qkv, ctx, s);
    return PDAB_EUNSUPPORTED;
}
