// PDA token encoder for sm_100a: everything between the point cloud and the transformer input of one PDA SA scale
// in ONE kernel —
//
//   ordered ball query -> gather -> Gaussian density + direction vectors            (PB/pointnet2_utils.py:567-614)
//   relative point position encoding [c, g, c - g, (g - c) / r] -> position MLP     (PB/pointnet2_modules.py:903-915)
//   density / max over the neighbourhood -> DensityNet (1 -> 16 -> 8 -> 1, ReLU)      (PB/pointnet2_modules.py:958-1006)
//   token = cat[pos, feat * scale, feat, glob[centre]] -> LayerNorm 1                (PB/pointnet2_modules.py:917-927,
//                                                                                     PB/PointFormer.py:29)
//
// so the (tokens, 4C) matrix the in_proj GEMM consumes is the only thing written to HBM: the grouped tensor, the
// 12-channel encoding, both position-MLP activations and the density scale never exist in memory.  (The unfused
// chain — pdab_pda_group_tokens, 2 GEMM launches, ~10 PyTorch elementwise launches, pdab_pda_assemble_ln_split —
// moved ~4.9 KB per token at C = 128; this kernel moves 2 KB + L2-resident gathers.)
//
// Mapping.  A CTA owns CTR consecutive centres of one scene.  Phase 1: thread per centre scans the cloud in index
// order (ball_scan.cuh; hit list in shared memory).  Phase 2: a warp takes 32 tokens (32 / nsample neighbourhoods)
// at a time:
//   lane = token  : geometry, density, neighbourhood max (segmented shuffles), DensityNet, 12-float encoding -> smem
//   lane = unit   : hidden layer of the position MLP for 16 tokens (W1 rows live in registers) -> smem
//   lane = CPL consecutive output channels, 8 tokens in registers: second layer from a transposed smem copy of W2
//                   (one 16-byte weight read feeds 8 x CPL FMAs; hidden activations are broadcast reads)
//   then per token the row [pos | feat * scale | feat | glob] is in the lane = 4-consecutive-channels layout, the
//   two LayerNorm reductions are shuffle sums (8 tokens interleaved for ILP) and the row is stored with fully
//   coalesced 16-byte writes.
// CUDA cores by design (north_star: "the PDA distribution-aware feature encoding stays on CUDA cores"); all fp32.
#include "ball_scan.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kDensFloats = 16 + 16 + 128 + 8 + 8 + 4;  // w1, b1, w2 (8 x 16), b2, w3, b3 (+pad)

template <int CPL>
struct Vec;
template <>
struct Vec<4> {
    using T = float4;
};
template <>
struct Vec<2> {
    using T = float2;
};

template <int CPL>
__device__ __forceinline__ void vload(float (&d)[CPL], const float *p) {
    if constexpr (CPL == 4) {
        const float4 v = *reinterpret_cast<const float4 *>(p);
        d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = v.w;
    } else {
        const float2 v = *reinterpret_cast<const float2 *>(p);
        d[0] = v.x, d[1] = v.y;
    }
}
template <int CPL>
__device__ __forceinline__ void vldg(float (&d)[CPL], const float *p) {
    if constexpr (CPL == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
        d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = v.w;
    } else {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
        d[0] = v.x, d[1] = v.y;
    }
}
template <int CPL>
__device__ __forceinline__ void vstore(float *p, const float (&d)[CPL]) {
    if constexpr (CPL == 4) *reinterpret_cast<float4 *>(p) = make_float4(d[0], d[1], d[2], d[3]);
    else *reinterpret_cast<float2 *>(p) = make_float2(d[0], d[1]);
}

struct EncParams {
    int n, m, nsample;
    float radius, r2, two_r2, dens_norm, eps;
    const float *xyz, *new_xyz, *features_t, *glob, *params;
    float *out;
};

// params (floats): W1 [H][12] | b1 [H] | W2t [H][C] | b2 [C] | dens (kDensFloats) | gamma [4C] | beta [4C],  H = C / 2
template <int C>
__global__ void __launch_bounds__(kThreads, 2) pda_encode_ln_kernel(const EncParams p) {
    constexpr int H = C / 2, CPL = C / 32, HPL = H / 32, E = 4 * C, CTR = 32, STRIDE = CTR + 1;
    constexpr int TB = CPL == 4 ? 4 : 8;  // tokens per register block of the second layer (TB * CPL accumulators)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *tile = reinterpret_cast<float4 *>(smem_raw);                   // kWarps * kSplitTile
    float *sW2t = reinterpret_cast<float *>(tile + kWarps * pdab::kSplitTile);  // H * C
    float *sb2 = sW2t + H * C;                                             // C
    float *sgamma = sb2 + C;                                               // E
    float *sbeta = sgamma + E;                                             // E
    float *sdens = sbeta + E;                                              // kDensFloats
    float *sb1 = sdens + kDensFloats;                                      // H
    float *sW1t = sb1 + H;                                                 // 12 * H (transposed; used when HPL == 2)
    float *sctr = sW1t + 12 * H;                                           // 3 * CTR
    int *sidx = reinterpret_cast<int *>(sctr + 3 * CTR);                   // nsample * STRIDE
    // phase-1 scratch (per-slice hit lists) and phase-2 scratch (encodings, hidden activations) share one region
    float *srppe_all = reinterpret_cast<float *>(sidx + p.nsample * STRIDE);   // kWarps * 32 * 12
    float *sh_all = srppe_all + kWarps * 32 * 12;                          // kWarps * 16 * H
    int *slist = reinterpret_cast<int *>(srppe_all);                       // kWarps * nsample * 33
    int *scnt = slist + kWarps * p.nsample * 33;                           // kWarps * 32

    const int scene = blockIdx.y;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int j0 = blockIdx.x * CTR;
    const int ns = p.nsample;
    const float *xyz = p.xyz + (size_t)scene * p.n * 3;
    const float *features_t = p.features_t + (size_t)scene * p.n * C;

    // ---- parameters -> shared memory / registers
    const float *gW1 = p.params, *gb1 = gW1 + H * 12, *gW2t = gb1 + H, *gb2 = gW2t + H * C, *gdens = gb2 + C,
                *ggamma = gdens + kDensFloats, *gbeta = ggamma + E;
    for (int i = t; i < H * C / 4; i += kThreads)
        reinterpret_cast<float4 *>(sW2t)[i] = __ldg(reinterpret_cast<const float4 *>(gW2t) + i);
    for (int i = t; i < C; i += kThreads) sb2[i] = __ldg(gb2 + i);
    for (int i = t; i < E; i += kThreads) {
        sgamma[i] = __ldg(ggamma + i);
        sbeta[i] = __ldg(gbeta + i);
    }
    for (int i = t; i < kDensFloats; i += kThreads) sdens[i] = __ldg(gdens + i);
    for (int i = t; i < H; i += kThreads) sb1[i] = __ldg(gb1 + i);
    // first-layer weights: one hidden unit per lane keeps its row in registers; with two units per lane (C = 128) that
    // would push the token phase past 128 registers, so the rows are read from a transposed (conflict-free) smem copy
    constexpr int W1R = HPL == 1 ? 12 : 1;
    float w1[W1R];
    if constexpr (HPL == 1) {
#pragma unroll
        for (int i = 0; i < 12; i++) w1[i] = __ldg(gW1 + lane * 12 + i);
    } else {
        w1[0] = 0.f;
    }
    for (int i = t; i < 12 * H; i += kThreads) sW1t[i] = __ldg(gW1 + (i % H) * 12 + i / H);

    // ---- phase 1: ordered ball scan, thread per centre
    const int j = j0 + lane;
    const bool active = j < p.m;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (active) {
        const float *ctr = p.new_xyz + ((size_t)scene * p.m + j) * 3;
        cx = ctr[0], cy = ctr[1], cz = ctr[2];
    }
    if (t < CTR) {
        sctr[t * 3 + 0] = cx;
        sctr[t * 3 + 1] = cy;
        sctr[t * 3 + 2] = cz;
    }
    __syncthreads();  // parameters staged; the scan scratch is about to be written
    pdab::ball_scan_split_to_smem<kWarps, STRIDE>(p.n, xyz, active, cx, cy, cz, p.r2, ns, tile, slist, scnt, sidx);

    // ---- phase 2: tokens
    const int nctr = min(CTR, p.m - j0);
    const int ntok = nctr * ns;
    const int npass = (ntok + 31) >> 5;
    float *obase = p.out + ((size_t)scene * p.m + j0) * ns * E;
    const float *gbase = p.glob + ((size_t)scene * p.m + j0) * C;
    float *srppe = srppe_all + warp * 32 * 12;
    float *sh = sh_all + warp * 16 * H;

    for (int pass = warp; pass < npass; pass += kWarps) {
        // lane = token: geometry, density scale, encoding
        const int tok = min(pass * 32 + lane, ntok - 1);  // whole neighbourhoods are valid or not (ntok % ns == 0)
        const int jl = tok / ns, s = tok - jl * ns;
        const int k = sidx[s * STRIDE + jl];
        float sc;
        {
            const float gx = __ldg(xyz + (size_t)k * 3 + 0), gy = __ldg(xyz + (size_t)k * 3 + 1),
                        gz = __ldg(xyz + (size_t)k * 3 + 2);
            const float qx = sctr[jl * 3 + 0], qy = sctr[jl * 3 + 1], qz = sctr[jl * 3 + 2];
            const float dx = gx - qx, dy = gy - qy, dz = gz - qz;
            const float dist = sqrtf(dx * dx + dy * dy + dz * dz);  // torch.norm(...)**2, PB/pointnet2_utils.py:592-593
            const float dens = expf(-(dist * dist) / p.two_r2) / p.dens_norm;
            float mx = dens;
            for (int off = ns >> 1; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float d0 = dens / mx;
            // DensityNet: three folded conv+BN layers, ReLU after each (PB/pointnet2_modules.py:958-981)
            float h1[16];
#pragma unroll
            for (int i = 0; i < 16; i++) h1[i] = fmaxf(fmaf(sdens[i], d0, sdens[16 + i]), 0.f);
            float a3 = sdens[32 + 128 + 8 + 8];
#pragma unroll
            for (int o = 0; o < 8; o++) {
                float a = sdens[32 + 128 + o];
#pragma unroll
                for (int i = 0; i < 16; i++) a = fmaf(sdens[32 + o * 16 + i], h1[i], a);
                a3 = fmaf(sdens[32 + 128 + 8 + o], fmaxf(a, 0.f), a3);
            }
            sc = fmaxf(a3, 0.f);
            float4 *r = reinterpret_cast<float4 *>(srppe + lane * 12);
            r[0] = make_float4(qx, qy, qz, gx);
            r[1] = make_float4(gy, gz, qx - gx, qy - gy);
            r[2] = make_float4(qz - gz, dx / p.radius, dy / p.radius, dz / p.radius);
        }
        __syncwarp();

#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            // lane = hidden unit(s): first layer of the position MLP for 16 tokens
#pragma unroll 4
            for (int tk = 0; tk < 16; tk++) {
                const float4 *r = reinterpret_cast<const float4 *>(srppe + (half * 16 + tk) * 12);
                const float4 r0 = r[0], r1 = r[1], r2 = r[2];
                const float rv[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
#pragma unroll
                for (int u = 0; u < HPL; u++) {
                    float a = sb1[lane + 32 * u];
#pragma unroll
                    for (int i = 0; i < 12; i++)
                        a = fmaf(HPL == 1 ? w1[i % W1R] : sW1t[i * H + lane + 32 * u], rv[i], a);
                    sh[tk * H + lane + 32 * u] = fmaxf(a, 0.f);
                }
            }
            __syncwarp();

#pragma unroll 1
            for (int tb = 0; tb < 16 / TB; tb++) {
                const int tbase = half * 16 + tb * TB;           // first token (lane index) of this block of TB
                const int tok0 = pass * 32 + tbase;
                if (tok0 >= ntok) break;                         // warp-uniform: blocks of TB never straddle ntok
                // the TB neighbours' feature rows and the centre's global row: in flight during the second layer
                float feat[TB][CPL], gl[CPL];
#pragma unroll
                for (int q = 0; q < TB; q++) {
                    const int kq = __shfl_sync(0xffffffffu, k, tbase + q);
                    vldg<CPL>(feat[q], features_t + (size_t)kq * C + CPL * lane);
                }
                vldg<CPL>(gl, gbase + (size_t)(tok0 / ns) * C + CPL * lane);
                float acc[TB][CPL];
                {
                    float b[CPL];
                    vload<CPL>(b, sb2 + CPL * lane);
#pragma unroll
                    for (int q = 0; q < TB; q++)
#pragma unroll
                        for (int e = 0; e < CPL; e++) acc[q][e] = b[e];
                }
#pragma unroll 1
                for (int in0 = 0; in0 < H; in0 += 4) {
                    float wv[4][CPL];
#pragma unroll
                    for (int i = 0; i < 4; i++) vload<CPL>(wv[i], sW2t + (in0 + i) * C + CPL * lane);
#pragma unroll
                    for (int q = 0; q < TB; q++) {
                        const float4 hv = *reinterpret_cast<const float4 *>(sh + (tb * TB + q) * H + in0);
#pragma unroll
                        for (int e = 0; e < CPL; e++) {
                            acc[q][e] = fmaf(wv[0][e], hv.x, acc[q][e]);
                            acc[q][e] = fmaf(wv[1][e], hv.y, acc[q][e]);
                            acc[q][e] = fmaf(wv[2][e], hv.z, acc[q][e]);
                            acc[q][e] = fmaf(wv[3][e], hv.w, acc[q][e]);
                        }
                    }
                }
                // rows: [pos | feat * scale | feat | glob], LayerNorm (two-pass, like torch.nn.LayerNorm), store
                float scq[TB], mean[TB], rstd[TB];
#pragma unroll
                for (int q = 0; q < TB; q++) {
                    scq[q] = __shfl_sync(0xffffffffu, sc, tbase + q);
                    float sum = 0.f;
#pragma unroll
                    for (int e = 0; e < CPL; e++) {
                        acc[q][e] = fmaxf(acc[q][e], 0.f);
                        sum += acc[q][e] + feat[q][e] * scq[q] + feat[q][e] + gl[e];
                    }
                    mean[q] = sum;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                    for (int q = 0; q < TB; q++) mean[q] += __shfl_xor_sync(0xffffffffu, mean[q], off);
#pragma unroll
                for (int q = 0; q < TB; q++) {
                    mean[q] *= 1.0f / E;
                    float sq = 0.f;
#pragma unroll
                    for (int e = 0; e < CPL; e++) {
                        const float a = acc[q][e] - mean[q], b = feat[q][e] * scq[q] - mean[q],
                                    c = feat[q][e] - mean[q], d = gl[e] - mean[q];
                        sq += (a * a + b * b) + (c * c + d * d);
                    }
                    rstd[q] = sq;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                    for (int q = 0; q < TB; q++) rstd[q] += __shfl_xor_sync(0xffffffffu, rstd[q], off);
#pragma unroll
                for (int q = 0; q < TB; q++) rstd[q] = rsqrtf(rstd[q] * (1.0f / E) + p.eps);
                float *rows = obase + (size_t)tok0 * E + CPL * lane;
#pragma unroll
                for (int part = 0; part < 4; part++) {
                    float g[CPL], bt[CPL];
                    vload<CPL>(g, sgamma + part * C + CPL * lane);
                    vload<CPL>(bt, sbeta + part * C + CPL * lane);
#pragma unroll
                    for (int q = 0; q < TB; q++) {
                        float y[CPL];
#pragma unroll
                        for (int e = 0; e < CPL; e++) {
                            const float v = part == 0 ? acc[q][e] : part == 1 ? feat[q][e] * scq[q]
                                          : part == 2 ? feat[q][e] : gl[e];
                            y[e] = (v - mean[q]) * rstd[q] * g[e] + bt[e];
                        }
                        vstore<CPL>(rows + (size_t)q * E + part * C, y);
                    }
                }
            }
            __syncwarp();  // sh is rewritten by the next half / pass
        }
    }
}

template <int C>
size_t enc_smem(int nsample) {
    constexpr int H = C / 2, E = 4 * C, CTR = 32;
    const size_t phase2 = sizeof(float) * (kWarps * 32 * 12 + kWarps * 16 * H);
    const size_t phase1 = sizeof(int) * ((size_t)kWarps * nsample * 33 + kWarps * 32);
    return sizeof(float4) * kWarps * pdab::kSplitTile +
           sizeof(float) * ((size_t)H * C + C + 2 * E + kDensFloats + H + 12 * H + 3 * CTR) +
           sizeof(int) * (size_t)nsample * (CTR + 1) + (phase1 > phase2 ? phase1 : phase2);
}

template <int C>
int enc_launch(int b, const EncParams &p, cudaStream_t s) {
    const size_t smem = enc_smem<C>(p.nsample);
    auto kern = pda_encode_ln_kernel<C>;
    static size_t configured = 0;  // per instantiation
    if (smem > configured) {
        PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(pdab::div_up(p.m, 32), b);
    kern<<<grid, kThreads, smem, s>>>(p);
    PDAB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" size_t pdab_pda_encode_param_floats(int c) {
    if (c != 64 && c != 128) return 0;
    const size_t h = c / 2;
    return h * 12 + h + h * c + c + kDensFloats + 8 * (size_t)c;
}

extern "C" int pdab_pda_encode_ln(int b, int c, int n, int m, float radius, int nsample, const float *xyz,
                                  const float *new_xyz, const float *features_t, const float *glob,
                                  const float *params, float eps, float *out, pdab_stream_t stream) {
    if (b < 0 || n < 1 || m < 0 || nsample < 1 || !xyz || !new_xyz || !features_t || !glob || !params || !out)
        return PDAB_EINVAL;
    if (b == 0 || m == 0) return 0;
    if ((c != 64 && c != 128) || (nsample != 16 && nsample != 32) || b > 65535) return PDAB_EUNSUPPORTED;
    EncParams p{};
    p.n = n;
    p.m = m;
    p.nsample = nsample;
    p.radius = radius;
    p.r2 = radius * radius;
    // Python-double scalars are rounded to fp32 once, as torch does for tensor-scalar ops (PB/pointnet2_utils.py:593)
    p.two_r2 = (float)(2.0 * (double)radius * (double)radius);
    p.dens_norm = (float)(2.5 * (double)radius);
    p.eps = eps;
    p.xyz = xyz;
    p.new_xyz = new_xyz;
    p.features_t = features_t;
    p.glob = glob;
    p.params = params;
    p.out = out;
    cudaStream_t s = pdab::to_stream(stream);
    return c == 64 ? enc_launch<64>(b, p, s) : enc_launch<128>(b, p, s);
}
