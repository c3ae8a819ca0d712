// Fused second half of the PDA transformer block for d_model = 256 (sm_100a only), ONE persistent kernel:
//
//   z      = LayerNorm2( y + ctx . Wo^T + bo )                       out_proj + residual + norm2   (PB/PointFormer.py:31-34)
//   h      = relu( z . W1^T + b1 )                                   linear1 + activation          (PB/PointFormer.py:35)
//   pooled = max over the nsample tokens of a neighbourhood of ( z + h . W2^T + b2 )
//                                                                    linear2 + residual + max-pool  (PB/PointFormer.py:35-36,
//                                                                                                   PB/pointnet2_modules.py:931)
//
// z (tokens x 256) and h (tokens x 128) never reach HBM: per 128-token tile the LayerNorm epilogue writes z as an fp16
// K-major SWIZZLE_128B tile into shared memory, where it IS the A operand of the next tcgen05.mma, and parks the fp32 z in
// the accumulator's own TMEM columns for the last residual; h takes the same route.  Per token the kernel reads ctx (fp16,
// TMA) and the residual y ((hi, lo) fp16 planes, TMA) and writes 1 / nsample of a pooled row: 6 E bytes of HBM traffic
// against 18 E for the three-launch chain it replaces (pdab_tc_linear_h ADD_LN -> RELU -> ADD_MAXPOOL).
//
// Products: fp16 x fp16 single pass, fp32 accumulation in TMEM (tc_gemm.cu, NPASS = 4); residual streams at fp32 level.
//
// Mapping.  CTA pairs (cta_group::2, M = 256 tokens per pair, 128 per CTA), persistent over tiles.  Warps per CTA:
//   0      loader          : ctx k-atoms by TMA + the pre-packed weight k-atoms of all three GEMMs by cp.async.bulk, one ring
//   1      MMA issuer      : leader CTA issues the three GEMMs of a tile back to back, each gated by the epilogue that
//                            produces its A operand; peer CTA relays "my stage landed" to the leader
//   2      residual loader : y boxes (128 rows x 32 columns, hi + lo) by TMA into a 3-deep ring
//   4-11   epilogue        : two warps per TMEM lane quadrant, 16 token rows each (row statistics need no exchange)
// TMEM (512 columns): [0, 256) accumulator 1 -> fp32 z (parked in place) | [256, 384) accumulator 2 | [256, 512) accumulator 3.
// Shared memory: 3-stage ring {A 16 KB, W 16 KB} | z / h operand tiles 64 KB | residual ring 48 KB | barriers, parameters.
#include "tc_common.cuh"
#ifdef PDAB_FFN_TIMERS
#include <cstdio>
#endif

namespace {

constexpr int E = 256, F = 128;
constexpr int KA1 = E / 64, KA2 = E / 64, KA3 = F / 64;    // k-atoms (64 fp16 = one 128-byte swizzle row) of the three GEMMs
constexpr int S = 3;                                       // ring stages
constexpr int kStage = 2 * A_TILE_BYTES;                   // A slot + W slot
constexpr int kRing = S * kStage;
constexpr int kZBytes = KA2 * A_TILE_BYTES;                // z_hi (4 atoms); h reuses atoms 0, 1
constexpr int NR = 3;                                      // residual ring stages
constexpr int kRPlane = BM * 64;                           // one plane of a box: 128 rows x 32 fp16 columns (SWIZZLE_64B)
constexpr int kRStage = 2 * kRPlane;
constexpr int kZOff = kRing, kROff = kZOff + kZBytes, kBarOff = kROff + NR * kRStage, kParamOff = kBarOff + 512;
constexpr int kParamFloats = E + F + E + E + E;            // bo, b1, b2, gamma, beta
constexpr int kXchgOff = kParamOff + kParamFloats * 4;     // max-pool exchange: [2 slots][4 quadrants][32 columns]
constexpr int kSmemBytes = kXchgOff + 2 * 4 * 32 * 4 + 1024 /*align slack*/;
constexpr int EW = 8, kFirstEpi = 4, kThreads = 32 * (kFirstEpi + EW);
static_assert(kSmemBytes <= 215 * 1024 + 256, "shared memory budget");

struct FfnParams {
    long long T;
    long long n_tiles;          // 256-row pair tiles
    int ns;
    const uint8_t *Wo, *W1, *W2;   // packed fp16 images (pdab_tc_pack_weights, npass = 4; bn = 256 / 128 / 256)
    const float *bo, *b1, *b2, *gamma, *beta;
    float eps;
    float *out;                 // (T / ns, ldo) fp32
    int ldo;
    alignas(64) CUtensorMap tmCtx;   // fp16 (T, E): box 64 x 128, SWIZZLE_128B
    alignas(64) CUtensorMap tmYh;    // fp16 (T, E) planes: box 32 x 128, SWIZZLE_64B
    alignas(64) CUtensorMap tmYl;
};

__global__ void __launch_bounds__(kThreads, 1) ffn_fused_kernel(const __grid_constant__ FfnParams p) {
    const u32 rank = cluster_ctarank();
    const bool leader = rank == 0;
    const long long pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    extern __shared__ uint8_t smem_raw[];
    const u32 smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const u32 bar = smem_base + kBarOff;
    auto full = [&](int s) { return bar + 8u * s; };
    auto peer_full = [&](int s) { return bar + 8u * (S + s); };
    auto empty = [&](int s) { return bar + 8u * (2 * S + s); };
    auto acc_full = [&](int a) { return bar + 8u * (3 * S + a); };        // a = 0, 1, 2
    const u32 z_ready = bar + 8u * (3 * S + 3), h_ready = bar + 8u * (3 * S + 4), tile_free = bar + 8u * (3 * S + 5);
    auto r_full = [&](int s) { return bar + 8u * (3 * S + 6 + s); };
    auto r_empty = [&](int s) { return bar + 8u * (3 * S + 6 + NR + s); };
    const u32 tmem_slot = bar + 8u * (3 * S + 6 + 2 * NR);
    volatile u32 *tmem_slot_ptr = reinterpret_cast<volatile u32 *>(smem + kBarOff + 8 * (3 * S + 6 + 2 * NR));
    float *sbo = reinterpret_cast<float *>(smem + kParamOff), *sb1 = sbo + E, *sb2 = sb1 + F, *sgamma = sb2 + E,
          *sbeta = sgamma + E;
    float *sxchg = reinterpret_cast<float *>(smem + kXchgOff);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(full(s), 1);
            mbar_init(peer_full(s), 1);
            mbar_init(empty(s), 1);
        }
        for (int a = 0; a < 3; a++) mbar_init(acc_full(a), 1);
        mbar_init(z_ready, EW * 2);
        mbar_init(h_ready, EW * 2);
        mbar_init(tile_free, EW * 2);
        for (int r = 0; r < NR; r++) {
            mbar_init(r_full(r), 1);
            mbar_init(r_empty(r), EW);
        }
        fence_barrier_init();
    }
    cluster_sync_all();
    if (warp == 1) tmem_alloc<2>(tmem_slot, kTmemCols);
    for (int i = threadIdx.x; i < E; i += kThreads) {
        sbo[i] = __ldg(p.bo + i);
        sb2[i] = __ldg(p.b2 + i);
        sgamma[i] = __ldg(p.gamma + i);
        sbeta[i] = __ldg(p.beta + i);
        if (i < F) sb1[i] = __ldg(p.b1 + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const u32 tmem_base = *tmem_slot_ptr;
    auto row0_of = [&](long long tile) { return tile * (2 * BM) + (long long)rank * BM; };

    if (warp == 0) {
        // ===================================================================== loader
        if (lane == 0) {
            tma_prefetch_desc(&p.tmCtx);
            Pipe pipe;
            auto w_fill = [&](const uint8_t *src, u32 bytes) {     // this CTA's half of a weight k-atom -> the stage's W slot
                const u32 dst = smem_base + (u32)pipe.stage * kStage + A_TILE_BYTES;
                for (u32 o = 0; o < bytes; o += 8192) bulk_g2s(dst + o, src + o, 8192, full(pipe.stage));
            };
            for (long long tile = pair0; tile < p.n_tiles; tile += npairs) {
                const int arow0 = (int)row0_of(tile);
                for (int ka = 0; ka < KA1; ka++) {
                    mbar_wait(empty(pipe.stage), pipe.phase ^ 1);
                    mbar_arrive_expect_tx(full(pipe.stage), 2 * A_TILE_BYTES);
                    tma_load_2d(smem_base + (u32)pipe.stage * kStage, &p.tmCtx, ka * 64, arow0, full(pipe.stage));
                    w_fill(p.Wo + (size_t)ka * (E * 128) + (size_t)rank * (E / 2 * 128), E / 2 * 128);
                    pipe.advance<S>();
                }
                for (int ka = 0; ka < KA2; ka++) {
                    mbar_wait(empty(pipe.stage), pipe.phase ^ 1);
                    mbar_arrive_expect_tx(full(pipe.stage), F / 2 * 128);
                    w_fill(p.W1 + (size_t)ka * (F * 128) + (size_t)rank * (F / 2 * 128), F / 2 * 128);
                    pipe.advance<S>();
                }
                for (int ka = 0; ka < KA3; ka++) {
                    mbar_wait(empty(pipe.stage), pipe.phase ^ 1);
                    mbar_arrive_expect_tx(full(pipe.stage), E / 2 * 128);
                    w_fill(p.W2 + (size_t)ka * (E * 128) + (size_t)rank * (E / 2 * 128), E / 2 * 128);
                    pipe.advance<S>();
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (leader) / stage relay (peer)
        if (lane == 0) {
            Pipe pipe;
            constexpr int kFills = KA1 + KA2 + KA3;
            if (!leader) {
                for (long long tile = pair0; tile < p.n_tiles; tile += npairs)
                    for (int f = 0; f < kFills; f++) {
                        mbar_wait(full(pipe.stage), pipe.phase);
                        mbar_arrive_remote(peer_full(pipe.stage), 0);
                        pipe.advance<S>();
                    }
            } else {
                constexpr u32 idesc256 = umma_idesc<256, 0, 2>(), idesc128 = umma_idesc<128, 0, 2>();   // fp16 operands
                u32 tp = 0;                                       // tile parity of the once-per-tile barriers
                auto gemm = [&](int katoms, bool a_from_ring, u32 d, u32 idesc) {
                    for (int ka = 0; ka < katoms; ka++) {
                        mbar_wait(full(pipe.stage), pipe.phase);
                        mbar_wait(peer_full(pipe.stage), pipe.phase);
                        tc_fence_after();
                        const u32 a0 = a_from_ring ? smem_base + (u32)pipe.stage * kStage : smem_base + kZOff + (u32)ka * A_TILE_BYTES;
                        const u32 b0 = smem_base + (u32)pipe.stage * kStage + A_TILE_BYTES;
#pragma unroll
                        for (int kk = 0; kk < 4; kk++)
                            umma_bf16<2>(d, umma_desc(a0 + kk * 32), umma_desc(b0 + kk * 32), idesc, (ka | kk) ? 1u : 0u);
                        umma_commit<2>(empty(pipe.stage));
                        pipe.advance<S>();
                    }
                };
                // Phase timers of the MMA thread (-DPDAB_FFN_TIMERS, never in the library build).  Measured, cycles per tile of
                // ~21.6 k: wait tile_free (E3) 4.3 k | issue GEMM1 3.0 k (ring fills) | wait z_ready (GEMM1 + E1) 9.6 k |
                // issue GEMM2 1.8 k | wait h_ready (GEMM2 + E2) 1.4 k | issue GEMM3 0.9 k.
#ifdef PDAB_FFN_TIMERS
                long long ph[6] = {0, 0, 0, 0, 0, 0}, t0 = clock64(), t1;
                int ntile = 0;
#define FFN_TICK(i) t1 = clock64(); ph[i] += t1 - t0; t0 = t1;
#else
#define FFN_TICK(i)
#endif
                for (long long tile = pair0; tile < p.n_tiles; tile += npairs) {
                    mbar_wait(tile_free, tp ^ 1);                 // previous tile's last epilogue has left TMEM
                    tc_fence_after();
                    FFN_TICK(0)
                    gemm(KA1, true, tmem_base, idesc256);         // acc1 = ctx . Wo^T
                    umma_commit<2>(acc_full(0));
                    FFN_TICK(1)
                    mbar_wait(z_ready, tp);                       // both CTAs' z tiles are in shared memory
                    tc_fence_after();
                    FFN_TICK(2)
                    gemm(KA2, false, tmem_base + 256, idesc128);  // acc2 = z . W1^T
                    umma_commit<2>(acc_full(1));
                    FFN_TICK(3)
                    mbar_wait(h_ready, tp);
                    tc_fence_after();
                    FFN_TICK(4)
                    gemm(KA3, false, tmem_base + 256, idesc256);  // acc3 = h . W2^T
                    umma_commit<2>(acc_full(2));
                    FFN_TICK(5)
                    tp ^= 1;
#ifdef PDAB_FFN_TIMERS
                    ntile++;
#endif
                }
#ifdef PDAB_FFN_TIMERS
                if (blockIdx.x == 0 && ntile > 0)
                    printf("FFN cycles/tile (leader MMA thread, %d tiles): wait tile_free %lld | issue GEMM1 %lld | wait z_ready %lld | "
                           "issue GEMM2 %lld | wait h_ready %lld | issue GEMM3 %lld\n", ntile, ph[0] / ntile, ph[1] / ntile,
                           ph[2] / ntile, ph[3] / ntile, ph[4] / ntile, ph[5] / ntile);
#endif
            }
        }
    } else if (warp == 2) {
        // ===================================================================== residual loader
        if (lane == 0) {
            tma_prefetch_desc(&p.tmYh);
            tma_prefetch_desc(&p.tmYl);
            Pipe rp;
            for (long long tile = pair0; tile < p.n_tiles; tile += npairs) {
                const int rrow0 = (int)row0_of(tile);
                // (An L2 prefetch of the tile's residual slab and ctx atoms from here — cp.async.bulk.prefetch.tensor, half a tile
                // ahead of their use — shortened E1 by 12 % in an isolated launch and did nothing in the pipelined step.)
                for (int cb = 0; cb < E / 32; cb++) {
                    mbar_wait(r_empty(rp.stage), rp.phase ^ 1);
                    mbar_arrive_expect_tx(r_full(rp.stage), (u32)kRStage);
                    const u32 dst = smem_base + kROff + (u32)rp.stage * kRStage;
                    tma_load_2d(dst, &p.tmYh, 32 * cb, rrow0, r_full(rp.stage));
                    tma_load_2d(dst + kRPlane, &p.tmYl, 32 * cb, rrow0, r_full(rp.stage));
                    rp.advance<NR>();
                }
            }
        }
    } else if (warp >= kFirstEpi) {
        // ===================================================================== epilogue (8 warps, 16 rows each)
        const int q = warp & 3, half = (warp - kFirstEpi) >> 2;
        const int fr = lane >> 2, fc = (lane & 3) * 2;
        const int rbase = q * 32 + 16 * half;                     // this warp's first row inside the CTA's 128-row tile
        const u32 tlane = (u32)rbase << 16;
        // v[k * 4 + j * 2 + e]  <->  row rbase + 8 j + fr, column 8 k + fc + e of a 32-column block
        auto ld = [&](u32 col, float (&v)[16]) {
            tmem_ld_16x256b_x4(tmem_base + tlane + col, v);
            tmem_wait_ld();
        };
        auto st = [&](u32 col, const float (&v)[16]) {
            tmem_st_16x256b_x4(tmem_base + tlane + col, v);
            tmem_wait_st();
        };
        auto add_bias = [&](float (&v)[16], const float *b) {    // b -> the block's first column
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float2 b2 = *reinterpret_cast<const float2 *>(b + 8 * k + fc);
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    v[k * 4 + j * 2] += b2.x;
                    v[k * 4 + j * 2 + 1] += b2.y;
                }
            }
        };
        // fp16 pairs of a 32-column block -> the K-major SWIZZLE_128B operand tile (atom = 64 columns) of the next GEMM
        auto to_operand = [&](const float (&v)[16], int cb) {
            uint8_t *atom = smem + kZOff + (cb >> 1) * A_TILE_BYTES;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                uint8_t *rowp = atom + (rbase + 8 * j + fr) * 128 + fc * 2;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    *reinterpret_cast<u32 *>(rowp + ((((cb & 1) * 4 + k) ^ fr) << 4)) = f16x2(v[k * 4 + j * 2], v[k * 4 + j * 2 + 1]);
            }
        };
        auto signal = [&](u32 barrier) {                          // "this warp's part is done" -> the leader's MMA thread
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(barrier);
                else mbar_arrive_remote(barrier, 0);
            }
        };
        Pipe rp;
        u32 tp = 0;
        const int ns = p.ns;
        for (long long tile = pair0; tile < p.n_tiles; tile += npairs) {
            const long long grow = row0_of(tile) + rbase;         // global row of this warp's first token
            // ---------------------------------------------------------------- E1: LayerNorm(acc1 + bo + y) -> z
            mbar_wait(acc_full(0), tp);
            tc_fence_after();
            float sum[2] = {0.f, 0.f};
            for (int cb = 0; cb < E / 32; cb++) {
                float v[16];
                ld(32 * cb, v);
                add_bias(v, sbo + 32 * cb);
                mbar_wait(r_full(rp.stage), rp.phase);
                const uint8_t *rs = smem + kROff + (size_t)rp.stage * kRStage;
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int rr = rbase + 8 * j + fr;
                    const uint8_t *rowp = rs + rr * 64 + fc * 2;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int off = (k ^ ((fr >> 1) & 3)) << 4;
                        const float2 a = f16x2_to_float2(*reinterpret_cast<const u32 *>(rowp + off));
                        const float2 b = f16x2_to_float2(*reinterpret_cast<const u32 *>(rowp + kRPlane + off));
                        v[k * 4 + j * 2] += a.x + b.x;
                        v[k * 4 + j * 2 + 1] += a.y + b.y;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(r_empty(rp.stage));
                rp.advance<NR>();
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int k = 0; k < 4; k++) sum[j] += v[k * 4 + j * 2] + v[k * 4 + j * 2 + 1];
                st(32 * cb, v);
            }
            float mean[2], rstd[2], sq[2] = {0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 2; j++) {
                float t = sum[j];
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                t += __shfl_xor_sync(0xffffffffu, t, 2);
                mean[j] = t * (1.0f / E);
            }
            for (int cb = 0; cb < E / 32; cb++) {
                float v[16];
                ld(32 * cb, v);
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int k = 0; k < 4; k++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const float d = v[k * 4 + j * 2 + e] - mean[j];
                            sq[j] = fmaf(d, d, sq[j]);
                        }
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                float t = sq[j];
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                t += __shfl_xor_sync(0xffffffffu, t, 2);
                rstd[j] = rsqrtf(t * (1.0f / E) + p.eps);
            }
            for (int cb = 0; cb < E / 32; cb++) {
                float v[16];
                ld(32 * cb, v);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float2 g2 = *reinterpret_cast<const float2 *>(sgamma + 32 * cb + 8 * k + fc);
                    const float2 b2 = *reinterpret_cast<const float2 *>(sbeta + 32 * cb + 8 * k + fc);
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        v[k * 4 + j * 2] = (v[k * 4 + j * 2] - mean[j]) * rstd[j] * g2.x + b2.x;
                        v[k * 4 + j * 2 + 1] = (v[k * 4 + j * 2 + 1] - mean[j]) * rstd[j] * g2.y + b2.y;
                    }
                }
                st(32 * cb, v);            // fp32 z stays in TMEM for the last residual
                to_operand(v, cb);         // fp16 z = A operand of linear1
            }
            fence_proxy_async();
            tc_fence_before();
            signal(z_ready);
            // ---------------------------------------------------------------- E2: h = relu(acc2 + b1)
            mbar_wait(acc_full(1), tp);
            tc_fence_after();
            for (int cb = 0; cb < F / 32; cb++) {
                float v[16];
                ld(256 + 32 * cb, v);
                add_bias(v, sb1 + 32 * cb);
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = fmaxf(v[i], 0.f);
                to_operand(v, cb);
            }
            fence_proxy_async();
            tc_fence_before();
            signal(h_ready);
            // ---------------------------------------------------------------- E3: max over the neighbourhood of (acc3 + b2 + z)
            mbar_wait(acc_full(2), tp);
            tc_fence_after();
            for (int cb = 0; cb < E / 32; cb++) {
                float v[16], z[16];
                tmem_ld_16x256b_x4(tmem_base + tlane + 256 + 32 * cb, v);
                tmem_ld_16x256b_x4(tmem_base + tlane + 32 * cb, z);
                tmem_wait_ld();
                add_bias(v, sb2 + 32 * cb);
                float m[8];
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        float x = fmaxf(v[k * 4 + e] + z[k * 4 + e], v[k * 4 + 2 + e] + z[k * 4 + 2 + e]);   // rows fr, fr + 8
                        x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 4));
                        x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 8));
                        x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 16));
                        m[k * 2 + e] = x;                          // max over this warp's 16 rows
                    }
                if (ns == 16) {
                    if (fr == 0 && grow < p.T) {
                        float *o = p.out + (grow / 16) * p.ldo + 32 * cb + fc;
#pragma unroll
                        for (int k = 0; k < 4; k++) *reinterpret_cast<float2 *>(o + 8 * k) = make_float2(m[k * 2], m[k * 2 + 1]);
                    }
                } else {   // 32 rows = both warps of the quadrant: the upper half hands its maxima over through shared memory
                    float *slot = sxchg + ((cb & 1) * 4 + q) * 32;
                    if (half == 1 && fr == 0) {
#pragma unroll
                        for (int k = 0; k < 4; k++) *reinterpret_cast<float2 *>(slot + 8 * k + fc) = make_float2(m[k * 2], m[k * 2 + 1]);
                    }
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
                    if (half == 0 && fr == 0 && grow < p.T) {
                        float *o = p.out + (grow / 32) * p.ldo + 32 * cb + fc;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const float2 s2 = *reinterpret_cast<const float2 *>(slot + 8 * k + fc);
                            *reinterpret_cast<float2 *>(o + 8 * k) = make_float2(fmaxf(m[k * 2], s2.x), fmaxf(m[k * 2 + 1], s2.y));
                        }
                    }
                }
            }
            tc_fence_before();
            signal(tile_free);
            tp ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // the leader's MMAs read the peer's shared memory: nobody leaves before both are done
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2>(tmem_base, kTmemCols);
    }
}

// fp16 (rows, ld) row-major matrix -> boxes of `box_cols` columns x 128 rows, swizzle = the box's row bytes, zero fill
int make_box_map(CUtensorMap *map, const void *a, long long rows, int cols, int ld, int box_cols) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return PDAB_EUNSUPPORTED;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(a), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : PDAB_EINVAL;
}

}  // namespace

// See include/pdab.h for the contract.
extern "C" int pdab_tc_ffn_h(long long rows, int e, int nsample, const void *ctx, int ldc, const void *y_hi, const void *y_lo,
                             int ldy, const float *wo_packed, const float *bo, const float *gamma, const float *beta, float eps,
                             const float *w1_packed, const float *b1, const float *w2_packed, const float *b2, float *out,
                             int ldo, pdab_stream_t stream) {
    if (rows < 0 || !ctx || !y_hi || !y_lo || !wo_packed || !bo || !gamma || !beta || !w1_packed || !b1 || !w2_packed ||
        !b2 || !out)
        return PDAB_EINVAL;
    if (e != E || (nsample != 16 && nsample != 32)) return PDAB_EUNSUPPORTED;
    if (rows == 0) return 0;
    if (rows % nsample || rows > 0x7fffffffLL || (ldc & 7) || (ldy & 7) || ldc < e || ldy < e || (ldo & 1) || ldo < e ||
        ((reinterpret_cast<uintptr_t>(ctx) | reinterpret_cast<uintptr_t>(y_hi) | reinterpret_cast<uintptr_t>(y_lo)) & 15))
        return PDAB_EINVAL;
    FfnParams p{};
    p.T = rows;
    p.n_tiles = (rows + 2 * BM - 1) / (2 * BM);
    p.ns = nsample;
    p.Wo = reinterpret_cast<const uint8_t *>(wo_packed);
    p.W1 = reinterpret_cast<const uint8_t *>(w1_packed);
    p.W2 = reinterpret_cast<const uint8_t *>(w2_packed);
    p.bo = bo;
    p.b1 = b1;
    p.b2 = b2;
    p.gamma = gamma;
    p.beta = beta;
    p.eps = eps;
    p.out = out;
    p.ldo = ldo;
    int rc = make_box_map(&p.tmCtx, ctx, rows, e, ldc, 64);
    if (!rc) rc = make_box_map(&p.tmYh, y_hi, rows, e, ldy, 32);
    if (!rc) rc = make_box_map(&p.tmYl, y_lo, rows, e, ldy, 32);
    if (rc) return rc;
    long long grid = pdab::persistent_ctas();      // the calling thread's launch policy (pdab_set_persistent_ctas)
    if (p.n_tiles * 2 < grid) grid = p.n_tiles * 2;
    grid &= ~1LL;
    if (grid < 2) return PDAB_EUNSUPPORTED;
    cudaStream_t s = pdab::to_stream(stream);
    PDAB_CUDA(cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PDAB_CUDA(cudaLaunchKernelEx(&cfg, ffn_fused_kernel, p));
    PDAB_LAUNCH_CHECK();
    return 0;
}
