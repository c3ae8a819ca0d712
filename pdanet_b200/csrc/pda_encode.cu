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
// Mapping.  A CTA owns 32 consecutive centres of one scene.  Phase 1: CTA-wide ordered ball scan (ball_scan.cuh: lane =
// centre, every warp scans its own slice of the cloud; hit lists in shared memory).  Phase 2: a warp takes 32 tokens (32 /
// nsample neighbourhoods) at a time:
//   lane = token            : geometry, density, neighbourhood max (segmented shuffles), DensityNet, 12-float encoding -> smem
//   per 16-token m-tile     : position MLP on the warp-level tensor-core path — layer 1 (12 -> C/2) on mma.sync m16n8k8 TF32 with
//                             the 3x hi/lo compensation, layer 2 (C/2 -> C) on bf16 m16n8k16 split products (the product class
//                             of the split-bf16 GEMMs; its A fragments are the packed C fragments of layer 1, no re-layout).  The
//                             first version pushed the layers through the FMA pipe (lane = 4 output channels x 4 tokens): it was
//                             FMA-issue bound at 43 % pipe utilisation (profiles/r01_ncu_pda_encode_ln_*).
//   per token row           : [pos | feat * scale | feat | glob] in the MMA C-fragment layout (a quad of lanes owns 32 contiguous
//                             bytes of a row): LayerNorm statistics are two shuffles inside the quad, the row is written as
//                             full 32-byte sectors.
// The encoding itself (density, direction, DensityNet, LayerNorm) stays on CUDA cores (north_star); all accumulation in fp32.
#include <cuda_fp16.h>

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
// d -> fp16 hi plane and the fp16 remainders (lo plane): d = hi + lo to ~2^-22 relative
template <int CPL>
__device__ __forceinline__ void vstore_split16(__half *hi, __half *lo, const float (&d)[CPL]) {
    __half2 h[CPL / 2], l[CPL / 2];
#pragma unroll
    for (int i = 0; i < CPL / 2; i++) {
        h[i] = __floats2half2_rn(d[2 * i], d[2 * i + 1]);
        const float2 b = __half22float2(h[i]);
        l[i] = __floats2half2_rn(d[2 * i] - b.x, d[2 * i + 1] - b.y);
    }
    if constexpr (CPL == 4) {
        *reinterpret_cast<uint2 *>(hi) = make_uint2(*reinterpret_cast<unsigned *>(&h[0]), *reinterpret_cast<unsigned *>(&h[1]));
        *reinterpret_cast<uint2 *>(lo) = make_uint2(*reinterpret_cast<unsigned *>(&l[0]), *reinterpret_cast<unsigned *>(&l[1]));
    } else {
        *reinterpret_cast<__half2 *>(hi) = h[0];
        *reinterpret_cast<__half2 *>(lo) = l[0];
    }
}

__host__ __device__ constexpr int p16(int K) { return K / 2 + 4; }   // packed-pair row pitch (words): 4 * odd -> conflict-free
constexpr int kW1Pitch = 20;                                         // layer-1 weights: 12 inputs padded to 16, + 4

__device__ __forceinline__ void split_tf32(float x, unsigned &hi, unsigned &lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void split_bf16x2(float x0, float x1, unsigned &hi, unsigned &lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct EncParams {
    int n, m, nsample;
    float radius, r2, two_r2, dens_norm, eps;
    const float *xyz, *new_xyz, *features_t, *glob, *params;
    float *out;
    // fp16 (hi, lo) plane output (the fp16 single-pass transformer path): out_hi feeds the in_proj GEMM through TMA, hi + lo
    // is the fp32-level residual stream; out == nullptr then
    __half *out_hi, *out_lo;
};

// params (floats): W1 [H][12] | b1 [H] | W2t [H][C] | b2 [C] | dens (kDensFloats) | gamma [4C] | beta [4C],  H = C / 2
template <int C>
__global__ void __launch_bounds__(kThreads, 2) pda_encode_ln_kernel(const EncParams p) {
    constexpr int H = C / 2, E = 4 * C, CTR = 32, STRIDE = CTR + 1, P2 = p16(H);
    constexpr int CPL = C / 32;        // row phase: channels per lane and part
    constexpr int PP = C + 8;          // pitch of a staged position row: conflict-free float2 fragment stores, 16-byte aligned rows
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned *sW2h = reinterpret_cast<unsigned *>(smem_raw);               // C x P2 packed bf16 pairs (hi)
    unsigned *sW2l = sW2h + C * P2;                                        // C x P2 (lo)
    float *sW1 = reinterpret_cast<float *>(sW2l + C * P2);                 // H x kW1Pitch (12 inputs padded to 16 with zeros)
    float *sb1 = sW1 + H * kW1Pitch;                                       // H
    float *sb2 = sb1 + H;                                                  // C
    float *sgamma = sb2 + C;                                               // E
    float *sbeta = sgamma + E;                                             // E
    float *sdens = sbeta + E;                                              // kDensFloats
    float *sctr = sdens + kDensFloats;                                     // 3 * CTR
    int *sidx = reinterpret_cast<int *>(sctr + 3 * CTR);                   // nsample * STRIDE
    // one region for the phase-1 scratch (staging tiles, per-slice hit lists) and the phase-2 scratch (encodings, position rows)
    unsigned char *uni = reinterpret_cast<unsigned char *>(sidx + ((p.nsample * STRIDE + 3) & ~3));
    float4 *tile = reinterpret_cast<float4 *>(uni);                        // kWarps * kSplitTile
    int *slist = reinterpret_cast<int *>(tile + kWarps * pdab::kSplitTile);  // kWarps * nsample * 33
    int *scnt = slist + kWarps * p.nsample * 33;                           // kWarps * 32
    float *srppe_all = reinterpret_cast<float *>(uni);                     // kWarps * 32 * 12
    float *spos_all = srppe_all + kWarps * 32 * 12;                        // kWarps * 8 * PP

    const int scene = blockIdx.y;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int g = lane >> 2, t4 = lane & 3;                                // MMA fragment coordinates
    const int j0 = blockIdx.x * CTR;
    const int ns = p.nsample;
    const float *xyz = p.xyz + (size_t)scene * p.n * 3;
    const float *features_t = p.features_t + (size_t)scene * p.n * C;

    // ---- parameters -> shared memory
    const float *gW1 = p.params, *gb1 = gW1 + H * 12, *gW2t = gb1 + H, *gb2 = gW2t + H * C, *gdens = gb2 + C,
                *ggamma = gdens + kDensFloats, *gbeta = ggamma + E;
#pragma unroll 4
    for (int i = t; i < H * kW1Pitch; i += kThreads) {
        const int r = i / kW1Pitch, q = i - r * kW1Pitch;
        sW1[i] = q < 12 ? __ldg(gW1 + r * 12 + q) : 0.f;
    }
    // W2 (C x H) = W2t^T, packed: word (n, j) = inputs 2j, 2j + 1 of output n.  Consecutive threads take consecutive n: rows of
    // W2t are read whole (the j-fastest order strode 2 C floats between lanes — 32 sectors per load, 8 % of the kernel's
    // warp samples sat behind these loads), several loads in flight.
#pragma unroll 4
    for (int i = t; i < C * (H / 2); i += kThreads) {
        const int j = i / C, n = i - j * C;
        unsigned hi, lo;
        split_bf16x2(__ldg(gW2t + (2 * j) * C + n), __ldg(gW2t + (2 * j + 1) * C + n), hi, lo);
        sW2h[n * P2 + j] = hi;
        sW2l[n * P2 + j] = lo;
    }
    for (int i = t; i < H; i += kThreads) sb1[i] = __ldg(gb1 + i);
    for (int i = t; i < C; i += kThreads) sb2[i] = __ldg(gb2 + i);
    for (int i = t; i < E; i += kThreads) {
        sgamma[i] = __ldg(ggamma + i);
        sbeta[i] = __ldg(gbeta + i);
    }
    for (int i = t; i < kDensFloats; i += kThreads) sdens[i] = __ldg(gdens + i);

    // ---- phase 1: ordered ball scan, lane = centre, every warp scans its slice of the cloud
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
    const size_t obase_off = ((size_t)scene * p.m + j0) * ns * E;
    float *obase = p.out + obase_off;
    const float *gbase = p.glob + ((size_t)scene * p.m + j0) * C;
    float *srppe = srppe_all + warp * 32 * 12;
    float *spos = spos_all + warp * 8 * PP;

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
        for (int mt = 0; mt < 2; mt++) {                       // 16 tokens = one m-tile, inside one neighbourhood (16 | ns)
            const int tok0 = pass * 32 + 16 * mt;
            if (tok0 >= ntok) break;                           // warp-uniform
            // ---- layer 1: h1 = relu(W1 . rppe + b1), 12 (padded to 16) -> H, TF32 with the 3x compensation
            float h1[H / 8][4];
#pragma unroll
            for (int nt = 0; nt < H / 8; nt++) {
                const float b0 = sb1[nt * 8 + 2 * t4], b1 = sb1[nt * 8 + 2 * t4 + 1];
                h1[nt][0] = b0, h1[nt][1] = b1, h1[nt][2] = b0, h1[nt][3] = b1;
            }
#pragma unroll
            for (int ks = 0; ks < 2; ks++) {
                const float *r0 = srppe + (16 * mt + g) * 12 + 8 * ks + t4, *r1 = r0 + 8 * 12;
                unsigned ah[4], al[4];
                split_tf32(r0[0], ah[0], al[0]);                                   // (row g,     k = t)
                split_tf32(r1[0], ah[1], al[1]);                                   // (row g + 8, k = t)
                split_tf32(ks == 0 ? r0[4] : 0.f, ah[2], al[2]);                   // (row g,     k = t + 4): inputs 12..15 are padding
                split_tf32(ks == 0 ? r1[4] : 0.f, ah[3], al[3]);
                unsigned bh[H / 8][2], bl[H / 8][2];
#pragma unroll
                for (int nt = 0; nt < H / 8; nt++) {
                    const float *w = sW1 + (nt * 8 + g) * kW1Pitch + 8 * ks + t4;
                    split_tf32(w[0], bh[nt][0], bl[nt][0]);
                    split_tf32(w[4], bh[nt][1], bl[nt][1]);
                }
#pragma unroll
                for (int ps = 0; ps < 3; ps++)                                       // small terms first, tiles interleaved
#pragma unroll
                    for (int nt = 0; nt < H / 8; nt++)
                        mma_tf32(h1[nt], ps == 1 ? al : ah, ps == 0 ? bl[nt][0] : bh[nt][0], ps == 0 ? bl[nt][1] : bh[nt][1]);
            }
#pragma unroll
            for (int nt = 0; nt < H / 8; nt++)
#pragma unroll
                for (int e = 0; e < 4; e++) h1[nt][e] = fmaxf(h1[nt][e], 0.f);
            // ---- layer 2: pos = relu(W2 . h1 + b2), H -> C, bf16 m16n8k16 split products; A = packed C fragments of layer 1
            float pos[C / 8][4];
#pragma unroll
            for (int nt = 0; nt < C / 8; nt++) {
                const float b0 = sb2[nt * 8 + 2 * t4], b1 = sb2[nt * 8 + 2 * t4 + 1];
                pos[nt][0] = b0, pos[nt][1] = b1, pos[nt][2] = b0, pos[nt][3] = b1;
            }
#pragma unroll
            for (int ks = 0; ks < H / 16; ks++) {
                unsigned ah[4], al[4];
                split_bf16x2(h1[2 * ks][0], h1[2 * ks][1], ah[0], al[0]);
                split_bf16x2(h1[2 * ks][2], h1[2 * ks][3], ah[1], al[1]);
                split_bf16x2(h1[2 * ks + 1][0], h1[2 * ks + 1][1], ah[2], al[2]);
                split_bf16x2(h1[2 * ks + 1][2], h1[2 * ks + 1][3], ah[3], al[3]);
#pragma unroll
                for (int ps = 0; ps < 3; ps++)
#pragma unroll
                    for (int nt = 0; nt < C / 8; nt++) {
                        const int w = (nt * 8 + g) * P2 + ks * 8 + t4;
                        const unsigned *sw = ps == 0 ? sW2l : sW2h;
                        mma_bf16(pos[nt], ps == 1 ? al : ah, sw[w], sw[w + 4]);
                    }
            }
            // ---- rows: the position block goes through this warp's smem tile, 8 tokens (rows g, then rows g + 8) at a time, into
            // the row layout: lane = CPL consecutive channels of each of the four parts [pos | feat * scale | feat | glob], so a
            // row is four coalesced 16-byte accesses per lane; LayerNorm (two-pass, like torch.nn.LayerNorm) = two warp sums.
            float gl[CPL];
            vldg<CPL>(gl, gbase + (size_t)(tok0 / ns) * C + CPL * lane);
#pragma unroll
            for (int r = 0; r < 2; r++) {
                __syncwarp();  // the previous half has been read
#pragma unroll
                for (int nt = 0; nt < C / 8; nt++)
                    *reinterpret_cast<float2 *>(spos + g * PP + 8 * nt + 2 * t4) =
                        make_float2(fmaxf(pos[nt][2 * r], 0.f), fmaxf(pos[nt][2 * r + 1], 0.f));
                __syncwarp();
#pragma unroll 1
                for (int qb = 0; qb < 8; qb += 4) {            // 4 token rows at a time: their shuffle sums interleave
                    float pv[4][CPL], fv[4][CPL], scq[4], mean[4], rstd[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int tl = 16 * mt + 8 * r + qb + q;                    // the row's token, as a lane of stage A
                        const int kq = __shfl_sync(0xffffffffu, k, tl);
                        scq[q] = __shfl_sync(0xffffffffu, sc, tl);
                        vldg<CPL>(fv[q], features_t + (size_t)kq * C + CPL * lane);
                        vload<CPL>(pv[q], spos + (qb + q) * PP + CPL * lane);
                    }
                    // LayerNorm on packed fp32 pairs (pdab::f32x2): the same operations per element as the scalar two-pass form,
                    // two elements per FADD2 / FMUL2 / FFMA2 — the row phase was fp32-issue bound (FADD + FFMA + FMUL = 48 % of
                    // the kernel's instructions)
                    constexpr int NP = CPL / 2;
                    pdab::f32x2 pv2[4][NP], fv2[4][NP], fs2[4][NP], gl2[NP];
#pragma unroll
                    for (int e = 0; e < NP; e++) gl2[e] = pdab::pack2(gl[2 * e], gl[2 * e + 1]);
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const pdab::f32x2 sc2 = pdab::pack2(scq[q], scq[q]);
                        pdab::f32x2 acc = pdab::pack2(0.f, 0.f);
#pragma unroll
                        for (int e = 0; e < NP; e++) {
                            pv2[q][e] = pdab::pack2(pv[q][2 * e], pv[q][2 * e + 1]);
                            fv2[q][e] = pdab::pack2(fv[q][2 * e], fv[q][2 * e + 1]);
                            fs2[q][e] = pdab::mul2(fv2[q][e], sc2);                       // feat * scale (part 1)
                            acc = pdab::add2(acc, pdab::add2(pdab::add2(pv2[q][e], fs2[q][e]), pdab::add2(fv2[q][e], gl2[e])));
                        }
                        float s0, s1;
                        pdab::unpack2(acc, s0, s1);
                        mean[q] = s0 + s1;
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                        for (int q = 0; q < 4; q++) mean[q] += __shfl_xor_sync(0xffffffffu, mean[q], off);
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        mean[q] *= 1.0f / E;
                        const pdab::f32x2 m2 = pdab::pack2(mean[q], mean[q]);
                        pdab::f32x2 sq = pdab::pack2(0.f, 0.f);
#pragma unroll
                        for (int e = 0; e < NP; e++) {
                            const pdab::f32x2 a = pdab::sub2(pv2[q][e], m2), b = pdab::sub2(fs2[q][e], m2),
                                              c = pdab::sub2(fv2[q][e], m2), d = pdab::sub2(gl2[e], m2);
                            sq = pdab::fma2(a, a, pdab::fma2(b, b, pdab::fma2(c, c, pdab::fma2(d, d, sq))));
                        }
                        float s0, s1;
                        pdab::unpack2(sq, s0, s1);
                        rstd[q] = s0 + s1;
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                        for (int q = 0; q < 4; q++) rstd[q] += __shfl_xor_sync(0xffffffffu, rstd[q], off);
#pragma unroll
                    for (int q = 0; q < 4; q++) rstd[q] = rsqrtf(rstd[q] * (1.0f / E) + p.eps);
                    const size_t row_off = (size_t)(tok0 + 8 * r + qb) * E + CPL * lane;
                    float *rows = obase + row_off;
#pragma unroll
                    for (int part = 0; part < 4; part++) {
                        float gm[CPL], bt[CPL];
                        vload<CPL>(gm, sgamma + part * C + CPL * lane);
                        vload<CPL>(bt, sbeta + part * C + CPL * lane);
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const pdab::f32x2 m2 = pdab::pack2(mean[q], mean[q]), r2 = pdab::pack2(rstd[q], rstd[q]);
                            float y[CPL];
#pragma unroll
                            for (int e = 0; e < NP; e++) {
                                const pdab::f32x2 v = part == 0 ? pv2[q][e] : part == 1 ? fs2[q][e] : part == 2 ? fv2[q][e] : gl2[e];
                                const pdab::f32x2 o = pdab::fma2(pdab::mul2(pdab::sub2(v, m2), r2), pdab::pack2(gm[2 * e], gm[2 * e + 1]),
                                                                 pdab::pack2(bt[2 * e], bt[2 * e + 1]));
                                pdab::unpack2(o, y[2 * e], y[2 * e + 1]);
                            }
                            if (p.out_hi)
                                vstore_split16<CPL>(p.out_hi + obase_off + row_off + (size_t)q * E + part * C,
                                                    p.out_lo + obase_off + row_off + (size_t)q * E + part * C, y);
                            else
                                vstore<CPL>(rows + (size_t)q * E + part * C, y);
                        }
                    }
                }
            }
        }
        __syncwarp();  // srppe is rewritten by the next pass
    }
}

template <int C>
size_t enc_smem(int nsample) {
    constexpr int H = C / 2, E = 4 * C, CTR = 32, PP = C + 8;
    const size_t phase1 = sizeof(float4) * kWarps * pdab::kSplitTile + sizeof(int) * ((size_t)kWarps * nsample * 33 + kWarps * 32);
    const size_t phase2 = sizeof(float) * (kWarps * 32 * 12 + kWarps * 8 * PP);
    return sizeof(float) * ((size_t)2 * C * p16(H) + H * kW1Pitch + H + C + 2 * E + kDensFloats + 3 * CTR) +
           sizeof(int) * (((size_t)nsample * (CTR + 1) + 3) & ~(size_t)3) + (phase1 > phase2 ? phase1 : phase2);
}

template <int C>
int enc_launch(int b, const EncParams &p, cudaStream_t s) {
    const size_t smem = enc_smem<C>(p.nsample);
    auto kern = pda_encode_ln_kernel<C>;
    // per device and cheap: set on every launch (a cached flag would skip the second GPU of a process)
    PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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

static int encode_ln(int b, int c, int n, int m, float radius, int nsample, const float *xyz, const float *new_xyz,
                     const float *features_t, const float *glob, const float *params, float eps, float *out,
                     void *out_hi, void *out_lo, pdab_stream_t stream) {
    if (b < 0 || n < 1 || m < 0 || nsample < 1 || !xyz || !new_xyz || !features_t || !glob || !params ||
        (!out && !(out_hi && out_lo)))
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
    p.out_hi = reinterpret_cast<__half *>(out_hi);
    p.out_lo = reinterpret_cast<__half *>(out_lo);
    cudaStream_t s = pdab::to_stream(stream);
    return c == 64 ? enc_launch<64>(b, p, s) : enc_launch<128>(b, p, s);
}

extern "C" int pdab_pda_encode_ln(int b, int c, int n, int m, float radius, int nsample, const float *xyz,
                                  const float *new_xyz, const float *features_t, const float *glob,
                                  const float *params, float eps, float *out, pdab_stream_t stream) {
    if (!out) return PDAB_EINVAL;
    return encode_ln(b, c, n, m, radius, nsample, xyz, new_xyz, features_t, glob, params, eps, out, nullptr, nullptr, stream);
}

extern "C" int pdab_pda_encode_ln_h(int b, int c, int n, int m, float radius, int nsample, const float *xyz,
                                    const float *new_xyz, const float *features_t, const float *glob,
                                    const float *params, float eps, void *out_hi, void *out_lo, pdab_stream_t stream) {
    if (!out_hi || !out_lo) return PDAB_EINVAL;
    return encode_ln(b, c, n, m, radius, nsample, xyz, new_xyz, features_t, glob, params, eps, nullptr, out_hi, out_lo, stream);
}
