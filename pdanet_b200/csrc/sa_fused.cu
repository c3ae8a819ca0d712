// Fused plain set-abstraction scale for sm_100a:
//   ball query -> gather (xyz centred, features) -> 3 x (1x1 conv + folded BN + ReLU)
//   -> max over nsample,  one kernel, grouped tensor and activations never in HBM.
// Replaces, per scale, 1 ball-query + 2 group + cat + 3 x (conv, BN, ReLU) + max-pool
// launches of the reference (PB/pointnet2_utils.py:681-704,
// PB/pointnet2_modules.py:1655-1672).
//
// This file holds the CUDA-core variant for NARROW layers (first SA layer of
// PDA-SSD: 4 -> 16 -> 16 -> 32 and 4 -> 32 -> 32 -> 64): contraction depth 4..32
// is too shallow to feed tcgen05 tiles, the layer is bound by the neighbour
// search and gather, not by FLOPs.  One thread owns one query centre: it scans
// the cloud (ball_scan.cuh), then pushes its nsample neighbours through the MLP
// with the weights broadcast from shared memory (LDS.128: four weights per load,
// four FMAs per load) and keeps the running channel maxima in registers.
// Wide layers (centroid aggregation, 259 -> 256 -> ... -> 1024) are served by the
// tensor-core variant; unsupported shapes return PDAB_EUNSUPPORTED and the host
// layer composes the unfused CUDA ops instead.
#include "ball_scan.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kStride = kThreads + 1;

template <int CIN, int COUT, bool RELU_TO_ARRAY>
__device__ __forceinline__ void dense(const float *__restrict__ W, const float *__restrict__ bias,
                                      const float (&in)[CIN], float (&out)[COUT]) {
    static_assert(CIN % 4 == 0, "rows are read as float4");
#pragma unroll
    for (int r = 0; r < COUT; r++) {
        float acc = bias[r];
        const float4 *w4 = reinterpret_cast<const float4 *>(W + r * CIN);
#pragma unroll
        for (int q = 0; q < CIN / 4; q++) {
            const float4 w = w4[q];
            acc = fmaf(w.x, in[4 * q + 0], acc);
            acc = fmaf(w.y, in[4 * q + 1], acc);
            acc = fmaf(w.z, in[4 * q + 2], acc);
            acc = fmaf(w.w, in[4 * q + 3], acc);
        }
        if (RELU_TO_ARRAY)
            out[r] = fmaxf(acc, 0.f);
        else
            out[r] = fmaxf(out[r], acc);  // running max; out starts at 0 == ReLU floor
    }
}

constexpr int kMlpThreads = 256;  // 128 scan threads + 128 helpers; all 8 warps run the MLP phase

// ---- Phase 2 of the fused kernels on the warp-level tensor-core path --------------------------------------------------
// The three layers of a narrow SA scale are 16..32 x {4..8, 16..32} x {16..64} contractions per neighbourhood: far below
// a tcgen05 tile, but a perfect fit for mma.sync m16n8k8 (TF32 operands, fp32 accumulate) with the same 3x hi/lo error
// compensation as the big GEMMs (fp32-level results: x = x_hi + x_lo, acc += x_hi w_lo + x_lo w_hi + x_hi w_hi).
// A warp takes 32 (centre, neighbour) rows per pass = two m-tiles of 16 rows: one centre when nsample > 16, two centres
// otherwise (rows beyond nsample repeat the last neighbour, harmless for a max).  Activations never leave registers: the
// C fragments of one layer become the A fragments of the next through 8 shuffles per k-step; the max over the
// neighbourhood is 3 shuffles per output column.  (The first version pushed one neighbour per lane through the layers
// on the FMA pipe with broadcast weights: ~5400 instructions per pass against ~900 here.)
// Weights live in shared memory with row pitch wpitch(K) floats so that B-fragment loads (row g, column t) are
// conflict-free.
__host__ __device__ constexpr int wpitch(int K) { return K == 4 ? 4 : K + 4; }

__device__ __forceinline__ void split_tf32(float x, unsigned &hi, unsigned &lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32_3x(float (&d)[4], const float (&a)[4], float b0, float b1) {
    unsigned ah[4], al[4], bh[2], bl[2];
#pragma unroll
    for (int i = 0; i < 4; i++) split_tf32(a[i], ah[i], al[i]);
    split_tf32(b0, bh[0], bl[0]);
    split_tf32(b1, bh[1], bl[1]);
#define PDAB_MMA(A, B)                                                                                              \
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "     \
                 "{%0,%1,%2,%3};"                                                                                   \
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])                                                   \
                 : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(B[0]), "r"(B[1]))
    PDAB_MMA(ah, bl);  // small terms first
    PDAB_MMA(al, bh);
    PDAB_MMA(ah, bh);
#undef PDAB_MMA
}

// One layer on one m-tile: out[nt] (C fragments, N/8 tiles) = relu(A (16 x K) . W^T + bias).
// A fragments come from `afrag(s, a)`, which fills a[4] for k-step s.
template <int K, int N, class AFrag>
__device__ __forceinline__ void mma_layer(const float *sW, const float *sb, int g, int t, AFrag afrag,
                                          float (&out)[N / 8][4]) {
    constexpr int WP = wpitch(K);
#pragma unroll
    for (int nt = 0; nt < N / 8; nt++) {
        const float b0 = sb[nt * 8 + 2 * t], b1 = sb[nt * 8 + 2 * t + 1];
        out[nt][0] = b0;
        out[nt][1] = b1;
        out[nt][2] = b0;
        out[nt][3] = b1;
    }
#pragma unroll
    for (int s = 0; s < (K + 7) / 8; s++) {
        float a[4];
        afrag(s, a);
#pragma unroll
        for (int nt = 0; nt < N / 8; nt++) {
            const float *w = sW + (nt * 8 + g) * WP + s * 8 + t;
            const float w0 = w[0];
            const float w1 = (s * 8 + 4 < K) ? w[4] : 0.f;   // K = 4: columns 4..7 do not exist
            mma_tf32_3x(out[nt], a, w0, w1);
        }
    }
#pragma unroll
    for (int nt = 0; nt < N / 8; nt++)
#pragma unroll
        for (int e = 0; e < 4; e++) out[nt][e] = fmaxf(out[nt][e], 0.f);
}

// A fragments of k-step s from the previous layer's C fragments h[s] (row g: h[s][0..1], row g+8: h[s][2..3]).
__device__ __forceinline__ void c_to_a(const float (&h)[4], int g, int t, float (&a)[4]) {
    const int src = g * 4 + (t >> 1);
    const bool odd = t & 1;
    const float x0 = __shfl_sync(0xffffffffu, h[0], src), x1 = __shfl_sync(0xffffffffu, h[1], src);
    const float x2 = __shfl_sync(0xffffffffu, h[2], src), x3 = __shfl_sync(0xffffffffu, h[3], src);
    const float y0 = __shfl_sync(0xffffffffu, h[0], src + 2), y1 = __shfl_sync(0xffffffffu, h[1], src + 2);
    const float y2 = __shfl_sync(0xffffffffu, h[2], src + 2), y3 = __shfl_sync(0xffffffffu, h[3], src + 2);
    a[0] = odd ? x1 : x0;   // (row g,   col 8s + t)
    a[1] = odd ? x3 : x2;   // (row g+8, col 8s + t)
    a[2] = odd ? y1 : y0;   // (row g,   col 8s + t + 4)
    a[3] = odd ? y3 : y2;   // (row g+8, col 8s + t + 4)
}

// ---- layers 2 and 3 on bf16 m16n8k16 split products ("bf16x3": x = hi + lo in bf16, acc += hi w_lo + lo w_hi + hi w_hi,
// ~2^-17 relative per product, fp32 accumulation — the product class of the big split-bf16 GEMMs).  Two reasons: half the
// mma.sync count of m16n8k8 (the phase is bound by the legacy tensor path, ~18 cycles per instruction), and the C fragments
// of the previous layer (row g: channels 2t, 2t+1 of each 8-channel tile) ARE the k16 A fragments once packed — the eight
// shuffles per k-step of the TF32 re-layout (c_to_a) disappear.  Weights are split once per CTA into packed (hi, lo) pairs.
__host__ __device__ constexpr int p16(int K) { return K / 2 + 4; }   // row pitch in 32-bit words: 4 * odd -> conflict-free B loads

__device__ __forceinline__ void split_bf16x2(float x0, float x1, unsigned &hi, unsigned &lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// out (C fragments, N/8 tiles) = relu(in (C fragments, K/8 tiles) . W^T + bias); sWh / sWl: packed bf16 pairs, pitch p16(K)
// RELU = false: the last layer leaves its ReLU to the max-pool (max and ReLU commute: one FMNMX3 per pooled value instead of
// one FMNMX per accumulator element)
template <int K, int N, bool RELU = true>
__device__ __forceinline__ void mma_layer16(const unsigned *sWh, const unsigned *sWl, const float *sb, int g, int t,
                                            const float (&in)[K / 8][4], float (&out)[N / 8][4]) {
    constexpr int P = p16(K);
#pragma unroll
    for (int nt = 0; nt < N / 8; nt++) {
        const float b0 = sb[nt * 8 + 2 * t], b1 = sb[nt * 8 + 2 * t + 1];
        out[nt][0] = b0;
        out[nt][1] = b1;
        out[nt][2] = b0;
        out[nt][3] = b1;
    }
#pragma unroll
    for (int s = 0; s < K / 16; s++) {
        unsigned ah[4], al[4];
        split_bf16x2(in[2 * s][0], in[2 * s][1], ah[0], al[0]);          // row g,     k = 2t, 2t+1
        split_bf16x2(in[2 * s][2], in[2 * s][3], ah[1], al[1]);          // row g + 8
        split_bf16x2(in[2 * s + 1][0], in[2 * s + 1][1], ah[2], al[2]);  // row g,     k = 2t+8, 2t+9
        split_bf16x2(in[2 * s + 1][2], in[2 * s + 1][3], ah[3], al[3]);  // row g + 8
#pragma unroll
        for (int pass = 0; pass < 3; pass++)                              // small terms first; independent tiles interleaved
#pragma unroll
            for (int nt = 0; nt < N / 8; nt++) {
                const int w = (nt * 8 + g) * P + s * 8 + t;
                const unsigned *sw = pass == 0 ? sWl : sWh;
                mma_bf16(out[nt], pass == 1 ? al : ah, sw[w], sw[w + 4]);
            }
    }
    if constexpr (RELU)
#pragma unroll
    for (int nt = 0; nt < N / 8; nt++)
#pragma unroll
        for (int e = 0; e < 4; e++) out[nt][e] = fmaxf(out[nt][e], 0.f);
}

// ---- the fp16 single-pass form of the three layers (the `_h` entry point; the product class of the big GEMMs' fp16 mode:
// 11-bit operands, fp32 accumulation — what the reference's cuDNN convolutions run as TF32).  ONE mma.sync m16n8k16 per
// (k-step, n-tile) instead of three, no operand splitting; layer 1 pads its 4..8 inputs to one k16 step (lanes t < C0P / 2
// hold the channel pair 2t, 2t+1; the upper eight k are zero, so the second B register is a literal 0).
__device__ __forceinline__ unsigned pack_f16x2(float x0, float x1) {
    unsigned r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
    return r;
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__host__ __device__ constexpr int p1h() { return 4; }   // layer-1 fp16 image: words 0..3 of a row = channel pairs (k < 8)

// layer 1: a0 / a1 = the packed channel pair (2t, 2t+1) of rows g / g + 8
template <int N>
__device__ __forceinline__ void mma_layer1_h(const unsigned *sW, const float *sb, int g, int t, unsigned a0, unsigned a1,
                                             float (&out)[N / 8][4]) {
    const unsigned a[4] = {a0, a1, 0u, 0u};
#pragma unroll
    for (int nt = 0; nt < N / 8; nt++) {
        const float b0 = sb[nt * 8 + 2 * t], b1 = sb[nt * 8 + 2 * t + 1];
        out[nt][0] = b0;
        out[nt][1] = b1;
        out[nt][2] = b0;
        out[nt][3] = b1;
        mma_f16(out[nt], a, sW[(nt * 8 + g) * p1h() + t], 0u);
#pragma unroll
        for (int e = 0; e < 4; e++) out[nt][e] = fmaxf(out[nt][e], 0.f);
    }
}

template <int K, int N, bool RELU = true>
__device__ __forceinline__ void mma_layer16_h(const unsigned *sW, const float *sb, int g, int t, const float (&in)[K / 8][4],
                                              float (&out)[N / 8][4]) {
    constexpr int P = p16(K);
#pragma unroll
    for (int nt = 0; nt < N / 8; nt++) {
        const float b0 = sb[nt * 8 + 2 * t], b1 = sb[nt * 8 + 2 * t + 1];
        out[nt][0] = b0;
        out[nt][1] = b1;
        out[nt][2] = b0;
        out[nt][3] = b1;
    }
#pragma unroll
    for (int s = 0; s < K / 16; s++) {
        const unsigned a[4] = {pack_f16x2(in[2 * s][0], in[2 * s][1]), pack_f16x2(in[2 * s][2], in[2 * s][3]),
                               pack_f16x2(in[2 * s + 1][0], in[2 * s + 1][1]), pack_f16x2(in[2 * s + 1][2], in[2 * s + 1][3])};
#pragma unroll
        for (int nt = 0; nt < N / 8; nt++) {
            const int w = (nt * 8 + g) * P + s * 8 + t;
            mma_f16(out[nt], a, sW[w], sW[w + 4]);
        }
    }
    if constexpr (RELU)
#pragma unroll
    for (int nt = 0; nt < N / 8; nt++)
#pragma unroll
        for (int e = 0; e < 4; e++) out[nt][e] = fmaxf(out[nt][e], 0.f);
}

template <int C0P, int C1, int C2, int C3, bool H = false>
__device__ __forceinline__ void mlp_phase(int c, int n, int nsample, int nctr, const float *__restrict__ xyz,
                                          const float *__restrict__ features, const float *sW1, const float *sb1,
                                          const unsigned *sW2, const float *sb2, const unsigned *sW3, const float *sb3,
                                          const float *sctr, const int *sidx, float *sout) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const bool whole = nsample > 16;                // one centre per pass (32 rows) or two (16 rows each)
    const int cpw = whole ? 1 : 2;
    for (int base = warp * cpw; base < nctr; base += (kMlpThreads / 32) * cpw) {
        float best[C3 / 8][2];                      // running max over the m-tiles of one centre
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            const int jl = min(base + (whole ? 0 : mt), nctr - 1);  // tail: duplicate the last centre (idempotent store)
            // layer-1 A fragments straight from the cloud: rows r0 = 16 mt + g and r0 + 8, columns t and t + 4 (TF32 k8) or
            // the pair 2t, 2t + 1 (fp16 k16)
            float in[2][2];
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const int row = (whole ? 16 * mt : 0) + g + 8 * rr;        // neighbour slot inside the centre
                const int k = sidx[min(row, nsample - 1) * kStride + jl];
#pragma unroll
                for (int cc = 0; cc < 2; cc++) {
                    const int col = H ? 2 * t + cc : t + 4 * cc;
                    float v = 0.f;
                    if (col < 3) v = __ldg(xyz + (size_t)k * 3 + col) - sctr[jl * 3 + col];  // grouped_xyz -= new_xyz
                    else if (col - 3 < c && col < C0P) v = __ldg(features + (size_t)(col - 3) * n + k);
                    in[rr][cc] = v;
                }
            }
            float h1[C1 / 8][4], h2[C2 / 8][4], h3[C3 / 8][4];
            if constexpr (H) {
                mma_layer1_h<C1>(reinterpret_cast<const unsigned *>(sW1), sb1, g, t, pack_f16x2(in[0][0], in[0][1]),
                                 pack_f16x2(in[1][0], in[1][1]), h1);
                mma_layer16_h<C1, C2>(sW2, sb2, g, t, h1, h2);
                mma_layer16_h<C2, C3, false>(sW3, sb3, g, t, h2, h3);
            } else {
                mma_layer<C0P, C1>(sW1, sb1, g, t, [&](int, float (&a)[4]) {
                    a[0] = in[0][0];
                    a[1] = in[1][0];
                    a[2] = in[0][1];
                    a[3] = in[1][1];
                }, h1);
                mma_layer16<C1, C2>(sW2, sW2 + C2 * p16(C1), sb2, g, t, h1, h2);
                mma_layer16<C2, C3, false>(sW3, sW3 + C3 * p16(C2), sb3, g, t, h2, h3);
            }
            // max over the rows of the centre, ReLU included (floor 0): rows g / g+8 in the thread — for a 32-row centre both
            // m-tiles first — then over g by shuffles, once per centre
#pragma unroll
            for (int nt = 0; nt < C3 / 8; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    float x = fmaxf(fmaxf(h3[nt][e], h3[nt][2 + e]), (whole && mt == 1) ? best[nt][e] : 0.f);
                    if (!whole || mt == 1) {
                        x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 4));
                        x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 8));
                        x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 16));
                    }
                    best[nt][e] = x;
                }
            if ((!whole || mt == 1) && g == 0) {
#pragma unroll
                for (int nt = 0; nt < C3 / 8; nt++) {
                    sout[(nt * 8 + 2 * t) * kStride + jl] = best[nt][0];
                    sout[(nt * 8 + 2 * t + 1) * kStride + jl] = best[nt][1];
                }
            }
        }
    }
}


// W (rows x kreal, row-major, global) -> shared memory rows of pitch wpitch(K), zero padded to K columns
template <int K>
__device__ __forceinline__ void stage_weights(float *sW, const float *__restrict__ W, int rows, int kreal) {
    constexpr int WP = wpitch(K);
    for (int i = threadIdx.x; i < rows * WP; i += kMlpThreads) {
        const int r = i / WP, q = i - r * WP;
        sW[i] = q < kreal ? W[r * kreal + q] : 0.f;
    }
}

// W (rows x K, row-major, global) -> packed bf16 (hi | lo) images of pitch p16(K): word (r, j) = columns 2j, 2j+1 of row r
template <int K>
__device__ __forceinline__ void stage_weights16(unsigned *sW, const float *__restrict__ W, int rows) {
    constexpr int P = p16(K);
    unsigned *sWl = sW + rows * P;
    for (int i = threadIdx.x; i < rows * (K / 2); i += kMlpThreads) {
        const int r = i / (K / 2), j = i - r * (K / 2);
        unsigned hi, lo;
        split_bf16x2(W[r * K + 2 * j], W[r * K + 2 * j + 1], hi, lo);
        sW[r * P + j] = hi;
        sWl[r * P + j] = lo;
    }
}

// fp16 images (the `_h` kernels).  Layer 1: W (rows x kreal) -> rows of p1h() words, word j = channels 2j, 2j+1 (zero padded
// to 8 channels).  Layers 2, 3: W (rows x K) -> rows of pitch p16(K) words (the hi image's layout; no lo image).
__device__ __forceinline__ void stage_weights1_h(unsigned *sW, const float *__restrict__ W, int rows, int kreal) {
    for (int i = threadIdx.x; i < rows * p1h(); i += kMlpThreads) {
        const int r = i / p1h(), j = i - r * p1h();
        const float w0 = 2 * j < kreal ? W[r * kreal + 2 * j] : 0.f, w1 = 2 * j + 1 < kreal ? W[r * kreal + 2 * j + 1] : 0.f;
        sW[i] = pack_f16x2(w0, w1);
    }
}
template <int K>
__device__ __forceinline__ void stage_weights16_h(unsigned *sW, const float *__restrict__ W, int rows) {
    constexpr int P = p16(K);
    for (int i = threadIdx.x; i < rows * (K / 2); i += kMlpThreads) {
        const int r = i / (K / 2), j = i - r * (K / 2);
        sW[r * P + j] = pack_f16x2(W[r * K + 2 * j], W[r * K + 2 * j + 1]);
    }
}

// C0P: input width padded to a multiple of 4 (3 + C real channels, rest zero weights/inputs)
//
// Phase 1: threads 0..127 each scan the cloud for one centre (hit lists in shared memory).
// Phase 2: one LANE per (centre, neighbour): a warp takes one centre (nsample = 32) or two (nsample = 16) at a
// time, every lane pushes its neighbour through the three layers with the weights broadcast from shared memory,
// and the channel maximum over the neighbourhood is a shuffle butterfly.  (One thread per centre, the first
// version, left the FMA pipe idle: 2 warps per scheduler each on a 4-cycle dependent chain; per-sample lanes
// give 8 resident warps per scheduler of independent work.)
template <int C0P, int C1, int C2, int C3>
__global__ void __launch_bounds__(kMlpThreads, 2)
sa_fused_narrow_kernel(int c, int n, int m, float r2, int nsample, const float *__restrict__ xyz,
                       const float *__restrict__ new_xyz, const float *__restrict__ features,
                       const float *__restrict__ W1, const float *__restrict__ b1, const float *__restrict__ W2,
                       const float *__restrict__ b2, const float *__restrict__ W3, const float *__restrict__ b3,
                       float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *tile = reinterpret_cast<float4 *>(smem_raw);
    float *sW1 = reinterpret_cast<float *>(tile + pdab::kScanTile + 8);  // C1 x C0P
    float *sW2 = sW1 + C1 * wpitch(C0P);                              // C2 x wpitch(C1)
    float *sW3 = sW2 + C2 * 2 * p16(C1);                              // packed (hi | lo) bf16 pairs, C3 x 2 p16(C2) words
    float *sb1 = sW3 + C3 * 2 * p16(C2);
    float *sb2 = sb1 + C1;
    float *sb3 = sb2 + C2;
    float *sctr = sb3 + C3;                                           // 3 x kThreads
    float *sout = sctr + 3 * kThreads;                                // C3 x (kThreads + 1)
    int *sidx = reinterpret_cast<int *>(sout + C3 * kStride);         // nsample x kStride

    const int scene = blockIdx.y;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int j0 = blockIdx.x * kThreads;
    const int j = j0 + t;
    const bool active = t < kThreads && j < m;
    const int c0 = 3 + c;
    xyz += (size_t)scene * n * 3;
    if (c > 0) features += (size_t)scene * c * n;

    stage_weights<C0P>(sW1, W1, C1, c0);
    stage_weights16<C1>(reinterpret_cast<unsigned *>(sW2), W2, C2);
    stage_weights16<C2>(reinterpret_cast<unsigned *>(sW3), W3, C3);
    for (int i = t; i < C1; i += kMlpThreads) sb1[i] = b1[i];
    for (int i = t; i < C2; i += kMlpThreads) sb2[i] = b2[i];
    for (int i = t; i < C3; i += kMlpThreads) sb3[i] = b3[i];

    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (active) {
        const float *ctr = new_xyz + ((size_t)scene * m + j) * 3;
        cx = ctr[0];
        cy = ctr[1];
        cz = ctr[2];
    }
    if (t < kThreads) {
        sctr[t * 3 + 0] = cx;
        sctr[t * 3 + 1] = cy;
        sctr[t * 3 + 2] = cz;
    }
    pdab::ball_scan_to_smem<kThreads, kStride>(n, xyz, active, cx, cy, cz, r2, nsample, tile, sidx);

    const int nctr = min(kThreads, m - j0);
    mlp_phase<C0P, C1, C2, C3>(c, n, nsample, nctr, xyz, features, sW1, sb1, reinterpret_cast<const unsigned *>(sW2), sb2,
                               reinterpret_cast<const unsigned *>(sW3), sb3, sctr, sidx, sout);
    __syncthreads();
    for (int i = t; i < C3 * nctr; i += kMlpThreads) {
        const int r = i / nctr, jl = i - r * nctr;
        __stcs(out + ((size_t)scene * C3 + r) * m + j0 + jl, sout[r * kStride + jl]);
    }
}

template <int C0P, int C1, int C2, int C3>
int launch_narrow(int b, int c, int n, int m, float radius, int nsample, const float *xyz, const float *new_xyz,
                  const float *features, const float *const *W, const float *const *B, float *out,
                  cudaStream_t stream) {
    const size_t smem = sizeof(float4) * (pdab::kScanTile + 8) +
                        sizeof(float) * (C1 * wpitch(C0P) + C2 * 2 * p16(C1) + C3 * 2 * p16(C2) + C1 + C2 + C3 +
                                         3 * kThreads + C3 * kStride) +
                        sizeof(int) * (size_t)nsample * kStride;
    auto kern = sa_fused_narrow_kernel<C0P, C1, C2, C3>;
    PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(pdab::div_up(m, kThreads), b);
    kern<<<grid, kMlpThreads, smem, stream>>>(c, n, m, radius * radius, nsample, xyz, new_xyz, features, W[0], B[0], W[1],
                                           B[1], W[2], B[2], out);
    PDAB_LAUNCH_CHECK();
    return 0;
}

// Both scales of an SA layer in one kernel: ONE scan of the cloud fills the two hit lists (ball_scan2_to_smem), then the
// two MLP + max-pool phases run back to back; the outputs land in one (B, A3 + B3, M) tensor, i.e. already concatenated
// along the channel axis as PB/pointnet2_modules.py:1674 (torch.cat of the scales) wants it.
// GRID: the hit lists come from the scene's hashed cell list (grid_build_kernel, ball_scan.cuh) instead of a scan of the
// whole cloud — 27 cells per centre instead of N points; same lists, bit-identical outputs.
// H: the fp16 single-pass form of the MLP phase (same shared-memory carve-up; the fp16 images use the front of each slot).
template <int C0P, int A1, int A2, int A3, int B1, int B2, int B3, bool GRID, bool H>
__global__ void __launch_bounds__(kMlpThreads, 2)
sa_fused_pair_kernel(int c, int n, int m, float r2a, int ns_a, float r2b, int ns_b, const unsigned char *__restrict__ ws,
                     float inv_edge, const float *__restrict__ xyz,
                     const float *__restrict__ new_xyz, const float *__restrict__ features,
                     const float *__restrict__ Wa1, const float *__restrict__ ba1, const float *__restrict__ Wa2,
                     const float *__restrict__ ba2, const float *__restrict__ Wa3, const float *__restrict__ ba3,
                     const float *__restrict__ Wb1, const float *__restrict__ bb1, const float *__restrict__ Wb2,
                     const float *__restrict__ bb2, const float *__restrict__ Wb3, const float *__restrict__ bb3,
                     float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *tile = reinterpret_cast<float4 *>(smem_raw);
    float *sWa1 = reinterpret_cast<float *>(tile + pdab::kScanTile + 8);
    float *sWa2 = sWa1 + A1 * wpitch(C0P), *sWa3 = sWa2 + A2 * 2 * p16(A1);      // layers 2, 3: packed bf16 (hi | lo)
    float *sWb1 = sWa3 + A3 * 2 * p16(A2), *sWb2 = sWb1 + B1 * wpitch(C0P), *sWb3 = sWb2 + B2 * 2 * p16(B1);
    float *sba1 = sWb3 + B3 * 2 * p16(B2), *sba2 = sba1 + A1, *sba3 = sba2 + A2;
    float *sbb1 = sba3 + A3, *sbb2 = sbb1 + B1, *sbb3 = sbb2 + B2;
    float *sctr = sbb3 + B3;                                          // 3 x kThreads
    float *sout = sctr + 3 * kThreads;                                // (A3 + B3) x kStride
    int *sidx_a = reinterpret_cast<int *>(sout + (A3 + B3) * kStride);  // ns_a x kStride
    int *sidx_b = sidx_a + ns_a * kStride;                              // ns_b x kStride

    const int scene = blockIdx.y;
    const int t = threadIdx.x;
    const int j0 = blockIdx.x * kThreads;
    const int j = j0 + t;
    const bool active = t < kThreads && j < m;
    const int c0 = 3 + c;
    xyz += (size_t)scene * n * 3;
    if (c > 0) features += (size_t)scene * c * n;

    if constexpr (H) {
        static_assert(C0P <= 8 && p1h() <= wpitch(C0P), "layer-1 fp16 image fits the fp32 slot");
        stage_weights1_h(reinterpret_cast<unsigned *>(sWa1), Wa1, A1, c0);
        stage_weights1_h(reinterpret_cast<unsigned *>(sWb1), Wb1, B1, c0);
        stage_weights16_h<A1>(reinterpret_cast<unsigned *>(sWa2), Wa2, A2);
        stage_weights16_h<A2>(reinterpret_cast<unsigned *>(sWa3), Wa3, A3);
        stage_weights16_h<B1>(reinterpret_cast<unsigned *>(sWb2), Wb2, B2);
        stage_weights16_h<B2>(reinterpret_cast<unsigned *>(sWb3), Wb3, B3);
    } else {
        stage_weights<C0P>(sWa1, Wa1, A1, c0);
        stage_weights<C0P>(sWb1, Wb1, B1, c0);
        stage_weights16<A1>(reinterpret_cast<unsigned *>(sWa2), Wa2, A2);
        stage_weights16<A2>(reinterpret_cast<unsigned *>(sWa3), Wa3, A3);
        stage_weights16<B1>(reinterpret_cast<unsigned *>(sWb2), Wb2, B2);
        stage_weights16<B2>(reinterpret_cast<unsigned *>(sWb3), Wb3, B3);
    }
    for (int i = t; i < A1; i += kMlpThreads) sba1[i] = ba1[i];
    for (int i = t; i < A2; i += kMlpThreads) sba2[i] = ba2[i];
    for (int i = t; i < A3; i += kMlpThreads) sba3[i] = ba3[i];
    for (int i = t; i < B1; i += kMlpThreads) sbb1[i] = bb1[i];
    for (int i = t; i < B2; i += kMlpThreads) sbb2[i] = bb2[i];
    for (int i = t; i < B3; i += kMlpThreads) sbb3[i] = bb3[i];

    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (active) {
        const float *ctr = new_xyz + ((size_t)scene * m + j) * 3;
        cx = ctr[0];
        cy = ctr[1];
        cz = ctr[2];
    }
    if (t < kThreads) {
        sctr[t * 3 + 0] = cx;
        sctr[t * 3 + 1] = cy;
        sctr[t * 3 + 2] = cz;
    }
    if (GRID) {
        const unsigned char *wscene = ws + (size_t)scene * pdab::grid_scene_bytes(n);
        __syncthreads();  // weights / centres staged
        pdab::grid_scan2_to_smem<kThreads, kStride>(reinterpret_cast<const int *>(wscene),
                                                    reinterpret_cast<const float4 *>(wscene + pdab::grid_scene_ints() * sizeof(int)),
                                                    inv_edge, active, cx, cy, cz, r2a, ns_a, r2b, ns_b, sidx_a, sidx_b);
    } else {
        pdab::ball_scan2_to_smem<kThreads, kStride>(n, xyz, active, cx, cy, cz, r2a, ns_a, r2b, ns_b, tile, sidx_a, sidx_b);
    }

    const int nctr = min(kThreads, m - j0);
    mlp_phase<C0P, A1, A2, A3, H>(c, n, ns_a, nctr, xyz, features, sWa1, sba1, reinterpret_cast<const unsigned *>(sWa2), sba2,
                               reinterpret_cast<const unsigned *>(sWa3), sba3, sctr, sidx_a, sout);
    mlp_phase<C0P, B1, B2, B3, H>(c, n, ns_b, nctr, xyz, features, sWb1, sbb1, reinterpret_cast<const unsigned *>(sWb2), sbb2,
                               reinterpret_cast<const unsigned *>(sWb3), sbb3, sctr, sidx_b, sout + A3 * kStride);
    __syncthreads();
    for (int i = t; i < (A3 + B3) * nctr; i += kMlpThreads) {
        const int r = i / nctr, jl = i - r * nctr;
        __stcs(out + ((size_t)scene * (A3 + B3) + r) * m + j0 + jl, sout[r * kStride + jl]);
    }
}

template <int C0P, int A1, int A2, int A3, int B1, int B2, int B3, bool H>
int launch_pair(int b, int c, int n, int m, float ra, int ns_a, float rb, int ns_b, const float *xyz,
                const float *new_xyz, const float *features, const float *const *W, const float *const *B, float *out,
                void *workspace, cudaStream_t stream) {
    const size_t smem = sizeof(float4) * (pdab::kScanTile + 8) +
                        sizeof(float) * (A1 * wpitch(C0P) + A2 * 2 * p16(A1) + A3 * 2 * p16(A2) + B1 * wpitch(C0P) +
                                         B2 * 2 * p16(B1) + B3 * 2 * p16(B2) + A1 + A2 + A3 + B1 + B2 + B3 + 3 * kThreads +
                                         (A3 + B3) * kStride) +
                        sizeof(int) * (size_t)(ns_a + ns_b) * kStride;
    dim3 grid(pdab::div_up(m, kThreads), b);
    if (workspace) {
        const float inv_edge = 1.0f / (pdab::kCellSlack * fmaxf(ra, rb));
        unsigned char *ws = static_cast<unsigned char *>(workspace);
        pdab::grid_build_kernel<<<b, pdab::kBuildThreads, 0, stream>>>(n, inv_edge, xyz, ws);
        PDAB_LAUNCH_CHECK();
        auto kern = sa_fused_pair_kernel<C0P, A1, A2, A3, B1, B2, B3, true, H>;
        PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kMlpThreads, smem, stream>>>(c, n, m, ra * ra, ns_a, rb * rb, ns_b, ws, inv_edge, xyz, new_xyz, features,
                                                  W[0], B[0], W[1], B[1], W[2], B[2], W[3], B[3], W[4], B[4], W[5], B[5], out);
        PDAB_LAUNCH_CHECK();
        return 0;
    }
    auto kern = sa_fused_pair_kernel<C0P, A1, A2, A3, B1, B2, B3, false, H>;
    PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kMlpThreads, smem, stream>>>(c, n, m, ra * ra, ns_a, rb * rb, ns_b, nullptr, 0.f, xyz, new_xyz, features,
                                              W[0], B[0], W[1], B[1], W[2], B[2], W[3], B[3], W[4], B[4], W[5], B[5], out);
    PDAB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int pdab_sa_fused(int b, int c, int n, int m, float radius, int nsample, const float *xyz,
                             const float *new_xyz, const float *features, int nlayers, const int *dims_host,
                             const float *const *weights_host, const float *const *biases_host, float *out,
                             pdab_stream_t stream) {
    if (b < 0 || c < 0 || n < 1 || m < 0 || nsample < 1 || !xyz || !new_xyz || !out || !dims_host || !weights_host ||
        !biases_host || (c > 0 && !features))
        return PDAB_EINVAL;
    if (b == 0 || m == 0) return 0;
    if (dims_host[0] != 3 + c) return PDAB_EINVAL;
    if (nlayers != 3 || nsample > 32 || b > 65535) return PDAB_EUNSUPPORTED;
    for (int l = 0; l < 3; l++)
        if (!weights_host[l] || !biases_host[l]) return PDAB_EINVAL;
    cudaStream_t s = pdab::to_stream(stream);
    const int d0 = dims_host[0], d1 = dims_host[1], d2 = dims_host[2], d3 = dims_host[3];
#define PDAB_NARROW(C0P, C1, C2, C3)                                                                              \
    if (d0 <= C0P && d0 > C0P - 4 && d1 == C1 && d2 == C2 && d3 == C3)                                            \
        return launch_narrow<C0P, C1, C2, C3>(b, c, n, m, radius, nsample, xyz, new_xyz, features, weights_host,  \
                                              biases_host, out, s);
    PDAB_NARROW(4, 16, 16, 32)
    PDAB_NARROW(4, 32, 32, 64)
    PDAB_NARROW(8, 16, 16, 32)
    PDAB_NARROW(8, 32, 32, 64)
#undef PDAB_NARROW
    return PDAB_EUNSUPPORTED;
}

template <bool H>
static int sa_fused_pair_entry(int b, int c, int n, int m, float radius_a, int nsample_a, float radius_b, int nsample_b,
                               const float *xyz, const float *new_xyz, const float *features, const int *dims_a_host,
                               const int *dims_b_host, const float *const *weights_host, const float *const *biases_host,
                               float *out, void *workspace, pdab_stream_t stream) {
    if (b < 0 || c < 0 || n < 1 || m < 0 || nsample_a < 1 || nsample_b < 1 || !xyz || !new_xyz || !out || !dims_a_host ||
        !dims_b_host || !weights_host || !biases_host || (c > 0 && !features))
        return PDAB_EINVAL;
    if (b == 0 || m == 0) return 0;
    if (dims_a_host[0] != 3 + c || dims_b_host[0] != 3 + c) return PDAB_EINVAL;
    if (nsample_a > 32 || nsample_b > 32 || b > 65535) return PDAB_EUNSUPPORTED;
    for (int l = 0; l < 6; l++)
        if (!weights_host[l] || !biases_host[l]) return PDAB_EINVAL;
    cudaStream_t s = pdab::to_stream(stream);
    const int d0 = dims_a_host[0];
    const bool a_small = dims_a_host[1] == 16 && dims_a_host[2] == 16 && dims_a_host[3] == 32;
    const bool b_large = dims_b_host[1] == 32 && dims_b_host[2] == 32 && dims_b_host[3] == 64;
    if (a_small && b_large && d0 <= 4)
        return launch_pair<4, 16, 16, 32, 32, 32, 64, H>(b, c, n, m, radius_a, nsample_a, radius_b, nsample_b, xyz, new_xyz,
                                                         features, weights_host, biases_host, out, workspace, s);
    if (a_small && b_large && d0 <= 8)
        return launch_pair<8, 16, 16, 32, 32, 32, 64, H>(b, c, n, m, radius_a, nsample_a, radius_b, nsample_b, xyz, new_xyz,
                                                         features, weights_host, biases_host, out, workspace, s);
    return PDAB_EUNSUPPORTED;
}

extern "C" int pdab_sa_fused_pair(int b, int c, int n, int m, float radius_a, int nsample_a, float radius_b,
                                  int nsample_b, const float *xyz, const float *new_xyz, const float *features,
                                  const int *dims_a_host, const int *dims_b_host, const float *const *weights_host,
                                  const float *const *biases_host, float *out, void *workspace, pdab_stream_t stream) {
    return sa_fused_pair_entry<false>(b, c, n, m, radius_a, nsample_a, radius_b, nsample_b, xyz, new_xyz, features, dims_a_host,
                                      dims_b_host, weights_host, biases_host, out, workspace, stream);
}

extern "C" int pdab_sa_fused_pair_h(int b, int c, int n, int m, float radius_a, int nsample_a, float radius_b,
                                    int nsample_b, const float *xyz, const float *new_xyz, const float *features,
                                    const int *dims_a_host, const int *dims_b_host, const float *const *weights_host,
                                    const float *const *biases_host, float *out, void *workspace, pdab_stream_t stream) {
    return sa_fused_pair_entry<true>(b, c, n, m, radius_a, nsample_a, radius_b, nsample_b, xyz, new_xyz, features, dims_a_host,
                                     dims_b_host, weights_host, biases_host, out, workspace, stream);
}

extern "C" size_t pdab_sa_grid_workspace_bytes(int b, int n) {
    if (b < 1 || n < 1) return 0;
    return (size_t)b * pdab::grid_scene_bytes(n);
}
