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
// One CTA (4 warps) per (neighbourhood, head).  Q, K, V tiles (NS x HD) are staged in shared memory with coalesced
// float4 loads; the two contractions are NS x NS x HD — far below a tcgen05 tile (M = 128) — so they run on the
// warp-level tensor-core path, mma.sync m16n8k8 TF32, with the same error compensation as the big GEMMs:
//   x = x_hi + x_lo (hi = top 19 bits, exact in TF32),  acc += x_hi y_lo + x_lo y_hi + x_hi y_hi     (3 MMAs)
// i.e. fp32-level scores and outputs (|err| ~ 1e-6 relative).  A first version on the FMA pipe was bound by shared
// memory bandwidth (97.8 % l1tex: 9.2 k wavefronts per CTA, profiles/r01_ncu_group_attention_*); fragments cut that ~6x.
//   S = (Q / sqrt(hd)) K^T : warp w owns m-tile (w % MT) and NS/8/NW n-tiles; A / B fragments by conflict-free LDS.32
//                            (row pitch HD+4: bank = 4 g + t).
//   softmax               : S goes through shared memory once; thread (row, lane block) as before, max / sum by
//                            shuffles over the TPR lanes of a row, full-precision expf (tracks torch.softmax to ~1e-7).
//   O = P V               : A fragments from P (pitch NS+4), B fragments from V (pitch HD+8: bank = 8 t + g);
//                            each warp owns HD/8/NW output n-tiles and writes 32-byte row segments.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;

// BF: bf16 m16n8k16 split products (hi / lo bf16 pairs, three MMAs per k-step: the product class of the split-bf16 GEMMs, half
// the mma.sync count of the TF32 path).  Its fragments are 8-byte pairs, which want other conflict-free pitches.
template <int NS, int HD, bool BF = false>
struct AttnSmem {
    static constexpr int QP = BF ? HD + 8 : HD + 4;   // Q, K row pitch (floats)
    static constexpr int VP = BF ? HD + 4 : HD + 8;   // V row pitch
    static constexpr int PP = BF ? NS + 8 : NS + 4;   // S / P row pitch
    static constexpr int FLOATS = 2 * NS * QP + NS * VP + NS * PP;
    static constexpr int BYTES = FLOATS * 4;
};

__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// D (16x8, fp32) += A (16x8, row) . B (8x8, col), TF32 operands
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void mma_3x(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                       const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    mma_tf32(d, ah, bl);  // small terms first
    mma_tf32(d, al, bh);
    mma_tf32(d, ah, bh);
}

__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma16_3x(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                         const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    mma_bf16(d, ah, bl);  // small terms first
    mma_bf16(d, al, bh);
    mma_bf16(d, ah, bh);
}
// packed (hi, lo) pair of two consecutive floats at p
__device__ __forceinline__ void pair_at(const float *p, uint32_t &hi, uint32_t &lo) {
    const float2 v = *reinterpret_cast<const float2 *>(p);
    split_bf16x2(v.x, v.y, hi, lo);
}

template <int NS, int HD, bool BF>
__global__ void __launch_bounds__(kThreads) group_attention_kernel(long long groups, int heads, const float *__restrict__ qkv,
                                                                   float *__restrict__ ctx) {
    using SM = AttnSmem<NS, HD, BF>;
    constexpr int QP = SM::QP, VP = SM::VP, PP = SM::PP;
    constexpr int MT = NS / 16;             // m-tiles (16 query rows each): 1 or 2
    constexpr int NW = 4 / MT;              // warps along n
    constexpr int S_NT = (NS / 8 + NW - 1) / NW;   // score n-tiles per warp (NS=16: warps 2,3 idle in this phase)
    constexpr int O_NT = HD / 8 / NW;       // output n-tiles per warp
    constexpr int TPR = kThreads / NS;      // softmax: threads per query row
    constexpr int KPT = NS / TPR;           // softmax: keys per thread
    extern __shared__ __align__(16) float sm[];
    float *sQ = sm, *sK = sQ + NS * QP, *sV = sK + NS * QP, *sP = sV + NS * VP;

    const long long gh = blockIdx.x;
    const long long grp = gh / heads;
    const int h = (int)(gh - grp * heads);
    const int E = heads * HD;
    const float *base = qkv + grp * NS * 3LL * E + h * HD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;  // mma fragment coordinates

    // ---- stage Q, K, V (rows of HD contiguous floats) with cp.async: 16-byte copies straight into shared memory, all
    // of a thread's copies in flight at once (no register staging), so the stage-in costs one memory latency.  The
    // head_dim ** -0.5 scaling PyTorch applies to q is applied to the scores instead (same product, one rounding later).
    constexpr float scaling = HD == 64 ? 0.125f : 0.08838834764831845f;
    constexpr int F4 = HD / 4;
#pragma unroll 4
    for (int i = tid; i < 3 * NS * F4; i += kThreads) {
        const int part = i / (NS * F4);
        const int rem = i - part * (NS * F4);
        const int r = rem / F4, c4 = rem - r * F4;
        const float *src = base + (long long)r * 3 * E + part * E + c4 * 4;
        float *dst = part == 0 ? sQ + r * QP + c4 * 4 : part == 1 ? sK + r * QP + c4 * 4 : sV + r * VP + c4 * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int mt = warp % MT, nw = warp / MT;
    const int m0 = mt * 16;

    // ---- S = Q K^T on the tensor cores (3xTF32)
    if (nw * S_NT * 8 < NS) {
        float acc[S_NT][4];
#pragma unroll
        for (int n = 0; n < S_NT; n++) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
        if constexpr (BF) {
#pragma unroll 4
            for (int k0 = 0; k0 < HD; k0 += 16) {   // A: rows g, g + 8, columns k0 + 2t (+1) and k0 + 8 + 2t (+1)
                uint32_t ah[4], al[4];
                pair_at(sQ + (m0 + g) * QP + k0 + 2 * t, ah[0], al[0]);
                pair_at(sQ + (m0 + g + 8) * QP + k0 + 2 * t, ah[1], al[1]);
                pair_at(sQ + (m0 + g) * QP + k0 + 8 + 2 * t, ah[2], al[2]);
                pair_at(sQ + (m0 + g + 8) * QP + k0 + 8 + 2 * t, ah[3], al[3]);
#pragma unroll
                for (int n = 0; n < S_NT; n++) {
                    const int n0 = (nw * S_NT + n) * 8;
                    uint32_t bh[2], bl[2];
                    pair_at(sK + (n0 + g) * QP + k0 + 2 * t, bh[0], bl[0]);
                    pair_at(sK + (n0 + g) * QP + k0 + 8 + 2 * t, bh[1], bl[1]);
                    mma16_3x(acc[n], ah, al, bh, bl);
                }
            }
        } else {
#pragma unroll 4
        for (int k0 = 0; k0 < HD; k0 += 8) {
            uint32_t ah[4], al[4];
            split_tf32(sQ[(m0 + g) * QP + k0 + t], ah[0], al[0]);
            split_tf32(sQ[(m0 + g + 8) * QP + k0 + t], ah[1], al[1]);
            split_tf32(sQ[(m0 + g) * QP + k0 + t + 4], ah[2], al[2]);
            split_tf32(sQ[(m0 + g + 8) * QP + k0 + t + 4], ah[3], al[3]);
#pragma unroll
            for (int n = 0; n < S_NT; n++) {
                const int n0 = (nw * S_NT + n) * 8;
                uint32_t bh[2], bl[2];
                split_tf32(sK[(n0 + g) * QP + k0 + t], bh[0], bl[0]);
                split_tf32(sK[(n0 + g) * QP + k0 + t + 4], bh[1], bl[1]);
                mma_3x(acc[n], ah, al, bh, bl);
            }
        }
        }
#pragma unroll
        for (int n = 0; n < S_NT; n++) {
            const int n0 = (nw * S_NT + n) * 8;
            *reinterpret_cast<float2 *>(sP + (m0 + g) * PP + n0 + 2 * t) =
                make_float2(acc[n][0] * scaling, acc[n][1] * scaling);
            *reinterpret_cast<float2 *>(sP + (m0 + g + 8) * PP + n0 + 2 * t) =
                make_float2(acc[n][2] * scaling, acc[n][3] * scaling);
        }
    }
    __syncthreads();

    // ---- softmax over the NS keys of each row (the TPR lanes of a row are adjacent lanes of one warp)
    {
        const int i = tid / TPR, jb = tid % TPR;
        float s[KPT];
        float mx = -3.4e38f;
#pragma unroll
        for (int jj = 0; jj < KPT; jj++) {
            s[jj] = sP[i * PP + jb + jj * TPR];
            mx = fmaxf(mx, s[jj]);
        }
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
    }
    __syncthreads();

    // ---- O = P V on the tensor cores (3xTF32)
    {
        float acc[O_NT][4];
#pragma unroll
        for (int n = 0; n < O_NT; n++) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
        if constexpr (BF) {
#pragma unroll
            for (int k0 = 0; k0 < NS; k0 += 16) {
                uint32_t ah[4], al[4];
                pair_at(sP + (m0 + g) * PP + k0 + 2 * t, ah[0], al[0]);
                pair_at(sP + (m0 + g + 8) * PP + k0 + 2 * t, ah[1], al[1]);
                pair_at(sP + (m0 + g) * PP + k0 + 8 + 2 * t, ah[2], al[2]);
                pair_at(sP + (m0 + g + 8) * PP + k0 + 8 + 2 * t, ah[3], al[3]);
#pragma unroll
                for (int n = 0; n < O_NT; n++) {
                    const int n0 = (nw * O_NT + n) * 8;
                    uint32_t bh[2], bl[2];   // B: keys k0 + 2t, + 1 (and + 8, + 9), channel n0 + g
                    split_bf16x2(sV[(k0 + 2 * t) * VP + n0 + g], sV[(k0 + 2 * t + 1) * VP + n0 + g], bh[0], bl[0]);
                    split_bf16x2(sV[(k0 + 8 + 2 * t) * VP + n0 + g], sV[(k0 + 9 + 2 * t) * VP + n0 + g], bh[1], bl[1]);
                    mma16_3x(acc[n], ah, al, bh, bl);
                }
            }
        } else {
#pragma unroll
        for (int k0 = 0; k0 < NS; k0 += 8) {
            uint32_t ah[4], al[4];
            split_tf32(sP[(m0 + g) * PP + k0 + t], ah[0], al[0]);
            split_tf32(sP[(m0 + g + 8) * PP + k0 + t], ah[1], al[1]);
            split_tf32(sP[(m0 + g) * PP + k0 + t + 4], ah[2], al[2]);
            split_tf32(sP[(m0 + g + 8) * PP + k0 + t + 4], ah[3], al[3]);
#pragma unroll
            for (int n = 0; n < O_NT; n++) {
                const int n0 = (nw * O_NT + n) * 8;
                uint32_t bh[2], bl[2];
                split_tf32(sV[(k0 + t) * VP + n0 + g], bh[0], bl[0]);
                split_tf32(sV[(k0 + t + 4) * VP + n0 + g], bh[1], bl[1]);
                mma_3x(acc[n], ah, al, bh, bl);
            }
        }
        }
        float *orow0 = ctx + (grp * NS + m0 + g) * (long long)E + h * HD;
        float *orow1 = orow0 + 8LL * E;
#pragma unroll
        for (int n = 0; n < O_NT; n++) {
            const int n0 = (nw * O_NT + n) * 8;
            *reinterpret_cast<float2 *>(orow0 + n0 + 2 * t) = make_float2(acc[n][0], acc[n][1]);
            *reinterpret_cast<float2 *>(orow1 + n0 + 2 * t) = make_float2(acc[n][2], acc[n][3]);
        }
    }
}

template <int NS, int HD, bool BF>
int launch(long long groups, int heads, const float *qkv, float *ctx, cudaStream_t s) {
    using SM = AttnSmem<NS, HD, BF>;
    auto kern = group_attention_kernel<NS, HD, BF>;
    // per device and cheap: set on every launch (a cached flag would skip the second GPU of a process)
    PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::BYTES));
    const long long blocks = groups * heads;
    if (blocks > 2147483647LL) return PDAB_EUNSUPPORTED;
    kern<<<(unsigned)blocks, kThreads, SM::BYTES, s>>>(groups, heads, qkv, ctx);
    PDAB_LAUNCH_CHECK();
    return 0;
}


// ------------------------------------------------------------------------------------------------------------------
// fp16 form (the fp16 single-pass GEMM mode, tc_gemm.cu NPASS = 4): qkv and ctx are fp16 matrices — half the bytes of a
// kernel that is HBM-bound — and both contractions run as ONE fp16 m16n8k16 MMA per k-step (fp32 accumulation, fp32
// softmax).  Fragments come straight out of shared memory as packed half pairs: A / B of Q K^T by 4-byte loads (row pitch
// HD + 8 halves: bank = 4 g + t), B of P V by ldmatrix.trans (V is stored [key][channel]; its transpose is the col-major
// operand).  A warp owns one 16-row m-tile and a slice of the output channels; its scores and probabilities never leave
// registers (the C fragments of Q K^T are the A fragments of P V).
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo_half, float hi_half) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_half), "f"(lo_half));
    return r;
}

template <int NS, int HD>
__global__ void __launch_bounds__(kThreads) group_attention_h_kernel(long long groups, int heads,
                                                                     const unsigned short *__restrict__ qkv,
                                                                     unsigned short *__restrict__ ctx) {
    constexpr int P = HD + 8;               // row pitch in halves (272 B / 144 B: 16-byte aligned, conflict-free fragments)
    constexpr int MT = NS / 16;             // m-tiles
    constexpr int NW = 4 / MT;              // warps sharing an m-tile, each owning HD / NW output channels
    constexpr int O_NT = HD / 8 / NW;       // output n-tiles per warp
    constexpr int S_NT = NS / 8;            // score n-tiles (all keys) per warp
    __shared__ __align__(16) unsigned short sQ[NS * P], sK[NS * P], sV[NS * P];

    const long long gh = blockIdx.x;
    const long long grp = gh / heads;
    const int h = (int)(gh - grp * heads);
    const int E = heads * HD;
    const unsigned short *base = qkv + grp * NS * 3LL * E + h * HD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    constexpr int C8 = HD / 8;              // 16-byte chunks per row
#pragma unroll 4
    for (int i = tid; i < 3 * NS * C8; i += kThreads) {
        const int part = i / (NS * C8);
        const int rem = i - part * (NS * C8);
        const int r = rem / C8, c8 = rem - r * C8;
        const unsigned short *src = base + (long long)r * 3 * E + part * E + c8 * 8;
        unsigned short *dst = (part == 0 ? sQ : part == 1 ? sK : sV) + r * P + c8 * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int mt = warp % MT, nw = warp / MT;
    const int m0 = mt * 16;
    constexpr float scaling = HD == 64 ? 0.125f : 0.08838834764831845f;

    // ---- S = Q K^T for this warp's 16 rows, all NS keys (the NW warps of an m-tile repeat it: 16 x NS x HD, cheap)
    float sc[S_NT][4];
#pragma unroll
    for (int n = 0; n < S_NT; n++) sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
#pragma unroll 4
    for (int k0 = 0; k0 < HD; k0 += 16) {
        uint32_t a[4];
        a[0] = *reinterpret_cast<const uint32_t *>(sQ + (m0 + g) * P + k0 + 2 * t);
        a[1] = *reinterpret_cast<const uint32_t *>(sQ + (m0 + g + 8) * P + k0 + 2 * t);
        a[2] = *reinterpret_cast<const uint32_t *>(sQ + (m0 + g) * P + k0 + 8 + 2 * t);
        a[3] = *reinterpret_cast<const uint32_t *>(sQ + (m0 + g + 8) * P + k0 + 8 + 2 * t);
#pragma unroll
        for (int n = 0; n < S_NT; n++) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t *>(sK + (8 * n + g) * P + k0 + 2 * t);
            const uint32_t b1 = *reinterpret_cast<const uint32_t *>(sK + (8 * n + g) * P + k0 + 8 + 2 * t);
            mma_f16(sc[n], a, b0, b1);
        }
    }
    // ---- softmax over the keys of rows g (values [.][0..1]) and g + 8 ([.][2..3]); a row lives in one quad of lanes
#pragma unroll
    for (int r = 0; r < 2; r++) {
        float mx = -3.4e38f;
#pragma unroll
        for (int n = 0; n < S_NT; n++) {
            sc[n][2 * r] *= scaling;
            sc[n][2 * r + 1] *= scaling;
            mx = fmaxf(mx, fmaxf(sc[n][2 * r], sc[n][2 * r + 1]));
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float sum = 0.f;
#pragma unroll
        for (int n = 0; n < S_NT; n++) {
            sc[n][2 * r] = expf(sc[n][2 * r] - mx);
            sc[n][2 * r + 1] = expf(sc[n][2 * r + 1] - mx);
            sum += sc[n][2 * r] + sc[n][2 * r + 1];
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float inv = 1.0f / sum;
#pragma unroll
        for (int n = 0; n < S_NT; n++) {
            sc[n][2 * r] *= inv;
            sc[n][2 * r + 1] *= inv;
        }
    }
    // ---- O = P V for this warp's channel slice: keys 16 s .. 16 s + 15 per k-step
    float o[O_NT][4];
#pragma unroll
    for (int n = 0; n < O_NT; n++) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int s16 = 0; s16 < NS / 16; s16++) {
        uint32_t a[4];
        a[0] = pack_f16x2(sc[2 * s16][0], sc[2 * s16][1]);
        a[1] = pack_f16x2(sc[2 * s16][2], sc[2 * s16][3]);
        a[2] = pack_f16x2(sc[2 * s16 + 1][0], sc[2 * s16 + 1][1]);
        a[3] = pack_f16x2(sc[2 * s16 + 1][2], sc[2 * s16 + 1][3]);
#pragma unroll
        for (int n = 0; n < O_NT; n++) {
            const int n0 = (nw * O_NT + n) * 8;
            // ldmatrix.x2.trans: lanes 0-15 address the 16 key rows of the (16 keys x 8 channels) block
            const uint32_t addr = (uint32_t)__cvta_generic_to_shared(sV + (16 * s16 + (lane & 15)) * P + n0);
            uint32_t b0, b1;
            asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(addr));
            mma_f16(o[n], a, b0, b1);
        }
    }
    unsigned short *orow0 = ctx + (grp * NS + m0 + g) * (long long)E + h * HD;
    unsigned short *orow1 = orow0 + 8LL * E;
#pragma unroll
    for (int n = 0; n < O_NT; n++) {
        const int n0 = (nw * O_NT + n) * 8;
        *reinterpret_cast<uint32_t *>(orow0 + n0 + 2 * t) = pack_f16x2(o[n][0], o[n][1]);
        *reinterpret_cast<uint32_t *>(orow1 + n0 + 2 * t) = pack_f16x2(o[n][2], o[n][3]);
    }
}

template <int NS, int HD>
int launch_h(long long groups, int heads, const void *qkv, void *ctx, cudaStream_t s) {
    const long long blocks = groups * heads;
    if (blocks > 2147483647LL) return PDAB_EUNSUPPORTED;
    group_attention_h_kernel<NS, HD><<<(unsigned)blocks, kThreads, 0, s>>>(
        groups, heads, reinterpret_cast<const unsigned short *>(qkv), reinterpret_cast<unsigned short *>(ctx));
    PDAB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int pdab_group_attention_h(long long groups, int nsample, int heads, int head_dim, const void *qkv, void *ctx,
                                      pdab_stream_t stream) {
    if (groups < 0 || heads < 1 || !qkv || !ctx) return PDAB_EINVAL;
    if (groups == 0) return 0;
    cudaStream_t s = pdab::to_stream(stream);
    if (nsample == 16 && head_dim == 64) return launch_h<16, 64>(groups, heads, qkv, ctx, s);
    if (nsample == 32 && head_dim == 64) return launch_h<32, 64>(groups, heads, qkv, ctx, s);
    if (nsample == 16 && head_dim == 128) return launch_h<16, 128>(groups, heads, qkv, ctx, s);
    if (nsample == 32 && head_dim == 128) return launch_h<32, 128>(groups, heads, qkv, ctx, s);
    return PDAB_EUNSUPPORTED;
}

extern "C" int pdab_group_attention(long long groups, int nsample, int heads, int head_dim, int npass, const float *qkv,
                                    float *ctx, pdab_stream_t stream) {
    if (groups < 0 || heads < 1 || !qkv || !ctx || npass < 1 || npass > 3) return PDAB_EINVAL;
    if (groups == 0) return 0;
    cudaStream_t s = pdab::to_stream(stream);
    const bool bf = npass == 2;
#define PDAB_ATTN(NS_, HD_)                                                        \
    if (nsample == NS_ && head_dim == HD_)                                         \
        return bf ? launch<NS_, HD_, true>(groups, heads, qkv, ctx, s) : launch<NS_, HD_, false>(groups, heads, qkv, ctx, s);
    PDAB_ATTN(16, 64)
    PDAB_ATTN(32, 64)
    PDAB_ATTN(16, 128)
    PDAB_ATTN(32, 128)
#undef PDAB_ATTN
    return PDAB_EUNSUPPORTED;
}
