// tcgen05 / TMEM GEMM for the dense contractions of the PDA-SSD hot path (sm_100a only).
//
//   out = epilogue( A[T,K] . W[Nout,K]^T + bias )
//
// with A produced on the fly by a fused PROLOGUE (plain rows, or the ball-query gather of a set-abstraction scale)
// and the accumulator consumed in TMEM by a fused EPILOGUE (bias, ReLU, residual + LayerNorm, residual / ReLU +
// neighbourhood max-pool).  It replaces, for the wide plain-SA scales, the reference's group_points -> cat ->
// 3 x (cuDNN 1x1 conv, BN, ReLU) -> max_pool2d chain (PB/pointnet2_modules.py:1655-1672) and, for
// the PDA block, the in_proj / out_proj / linear1 / linear2 GEMMs of TransformerEncoderLayerPreNorm with the
// LayerNorm / residual / ReLU / max-pool launches between them (PB/PointFormer.py:28-38, PB/pointnet2_modules.py:929-933).
//
// Arithmetic.  kind::tf32 tensor-core products accumulated in fp32 in TMEM.
//   NPASS = 1 : operands rounded to TF32 (cvt.rna) — the precision class of the reference's cuDNN convolutions.
//   NPASS = 3 : error-compensated "3xTF32": x = x_hi + x_lo, w = w_hi + w_lo with the hi parts exactly representable
//               in TF32 (top 19 bits), acc += x_hi w_lo + x_lo w_hi + x_hi w_hi.  The dropped x_lo w_lo term is
//               O(2^-22) relative: fp32-level results (what nn.Linear computes in the reference) at 3 MMAs / k-step.
//   NPASS = 2 : the same compensation with bf16 halves (kind::f16, K = 16 per MMA): ~1e-5 relative, 3 MMAs / k-step.
//   NPASS = 4 : fp16 x fp16 single pass (kind::f16, fp32 accumulation): 11-bit significands, the TF32 precision class at
//               one MMA per k-step.  In this mode activations travel BETWEEN kernels as fp16 (half the HBM bytes) and the
//               residual streams of the transformer as (hi, lo) fp16 plane pairs (fp32-level, 22 bits): measured, the
//               residual stream is what sets the block's output error (5e-4 with rounded residuals, 8e-5 with exact
//               ones), the single-pass products contribute 4e-5 (tools/precision_study.py).
//               A operand: fp32 rows converted by the producer warps (A_ROWS / A_GATHER), or — A_TMA — an fp16 row-major
//               matrix loaded by the TMA engine: cp.async.bulk.tensor.2d (UTMALDG) boxes of 128 rows x 64 columns through a
//               SWIZZLE_128B tensor map land as the canonical K-major tile the UMMA descriptor expects; no producer warps,
//               no register round trip, no shared-memory stores from the LSU.
//
// Mapping to the SM.  One CTA per SM, persistent over (row tile, column group) work items.  Large problems run as
// CTA PAIRS (cta_group::2, thread-block clusters of 2): one tcgen05.mma covers M = 256 rows across the two SMs, each SM
// stages its own 128 rows of A and HALF of the weight tile, so per SM the tensor core reads 8 KB instead of 12 KB of
// shared memory per MMA and half the weight bytes arrive from L2 — the kernel is shared-memory-bandwidth bound
// (profiles/r01_ncu_tc_gemm_*).  The leader CTA issues the MMAs and multicasts its commits to both CTAs' barriers; the
// peer's producers / epilogue arrive on the leader's barriers remotely (mapa + mbarrier.arrive.release.cluster).
// Warps (single CTA: 14, pair: 12 per CTA):
//   warp 0      W loader   : one lane issues cp.async.bulk (TMA engine, UBLKCP) of pre-packed weight tiles -> smem
//   warp 1      MMA issuer : one lane issues tcgen05.mma (M=128, N=BN, K=8 per instruction), tcgen05.commit -> mbarriers
//   warps 2-9   A producers: global -> registers -> (hi, lo) split -> 128B-swizzled K-major smem tiles
//               (4 groups of 2 warps, group g owns every 4th k-atom: 4 k-atoms = 64 KB of loads in flight per SM)
//   warps 10-13 epilogue   : tcgen05.ld (TMEM -> registers) -> fused epilogue -> global
// Tiles: 128 rows (= TMEM lanes) x BN columns (fp32 accumulator columns) x 32-float k-atoms (one 128-byte swizzle
// row).  Smem ring of S stages {A_hi, A_lo, W_hi, W_lo}; TMEM holds two accumulator stages (or one 512-column full
// row for the LayerNorm epilogue) so the epilogue of item i overlaps the main loop of item i+1.
// Weights are packed once (host side, pdanet_b200/tc_pack.py) into the exact smem image of each (column chunk, k-atom)
// tile — canonical K-major SWIZZLE_128B layout — so a stage's weights are ONE contiguous bulk copy.
#include <type_traits>

#include "tc_common.cuh"

namespace {

constexpr int BK = 32;             // fp32 / tf32 elements per k-atom (128 B swizzle row); bf16 mode: 64 elements
__host__ __device__ constexpr bool BK_IS_64(int npass) { return npass == 2 || npass == 4; }
__host__ __device__ constexpr int bk_of(int npass) { return (npass == 2 || npass == 4) ? 64 : 32; }
// operand planes per k-atom: the split modes stage (hi, lo), TF32 / fp16 single-pass one plane
__host__ __device__ constexpr int parts_of(int npass) { return (npass == 2 || npass == 3) ? 2 : 1; }
constexpr int kStgPitch = 36;       // floats per row of an epilogue staging tile (32 + 4: 16-byte aligned, conflict-free)

enum ALoad { A_ROWS = 0, A_GATHER = 1, A_TMA = 2 };
enum Epi { E_STORE = 0, E_RELU = 1, E_ADD_LN = 2, E_ADD_MAXPOOL = 3, E_RELU_MAXPOOL = 4, E_ATTN = 5 };
constexpr int kAttnVP = 68;         // floats per row of the attention epilogue's V staging tile (64 + 4: conflict-free B fragments)

struct GemmParams {
    // A operand
    const float *A;      // (T, lda) rows                                   [A_ROWS]
    int lda;
    long long T;         // rows
    int K;               // logical K (columns of A used)
    int KA;              // k-atoms = ceil(K / 32)
    // gather prologue [A_GATHER]: row t -> point idx[t] of scene t / (M*ns); A row = [feat_t[b, i, 0:C], xyz[b,i]-new_xyz[b,j]]
    const int *idx;      // (B, M, ns) flat
    const float *feat_t; // (B, Nsrc, C) point-major
    const float *xyz;    // (B, Nsrc, 3)
    const float *new_xyz;  // (B, M, 3)
    int C, Nsrc, M, ns;
    // weights
    const float *Wp;     // packed tiles [chunk][k-atom][hi|lo][BN*32]
    const float *bias;   // (Nout) or null
    int Nout;            // logical output columns
    int n_groups;        // column groups (each NCH chunks of BN)
    // output
    float *out;
    int ldo;
    // epilogue extras
    const float *R;      // residual (T, ldr)
    int ldr;
    const float *gamma, *beta;
    float eps;
    long long n_items;
    // NPASS = 4 i/o formats.  out16: 0 = fp32 `out`; 1 = fp16 `out`; 2 = (hi, lo) fp16 planes `out` / `out_lo`
    // (E_ADD_LN: the next residual stream).  res16: 0 = fp32 residual R; 1 = R / R_lo are fp16 (hi, lo) planes.
    // Pooled outputs (E_*_MAXPOOL) are always fp32.
    int out16, res16;
    void *out_lo;
    const void *R_lo;
    int persistent_ctas;   // grid cap of this launch (<= SM count)
    // A_TMA: fp16 (T, lda) row-major A, box = 64 columns x 128 rows, SWIZZLE_128B, out-of-range elements read as zero
    alignas(64) CUtensorMap tmA;
    // the residual's (hi, lo) fp16 planes, same box shape (res_tma != 0: the epilogue reads the residual from a smem ring)
    alignas(64) CUtensorMap tmRh;
    alignas(64) CUtensorMap tmRl;
    int res_tma;
};

// ------------------------------------------------------------------------------------------- kernel

// CG = 1: one CTA per SM, M = 128.  CG = 2: CTA pair (cta_group::2), M = 256: each CTA produces its own 128 rows of A and
// holds HALF of the W tile (rows [rank * BN/2, +BN/2) of the chunk); the tensor cores of the pair read both halves, so
// per SM the MMA reads 8 KB of shared memory instead of 12 KB and receives half the weight bytes from L2.
// EW: epilogue warps per CTA.  4 = one per TMEM lane quadrant.  8 (CTA pairs, attention epilogue) = two per quadrant,
// each owning one 16-row m-tile; the CTA then has 16 warps = 4 warpgroups and the register file is re-divided with
// setmaxnreg: the two warpgroups that hold the loader / MMA / A-producer warps grow to kProdRegs, the two epilogue
// warpgroups shrink to kEpiRegs (128 * 2 * 152 + 128 * 2 * 104 = 65536).
constexpr int kProdRegs = 152, kEpiRegs = 104;
__host__ __device__ constexpr bool epi_has_res(int epi) { return epi == 2 || epi == 3; }   // E_ADD_LN, E_ADD_MAXPOOL
template <int EPI, int CG, int ALOAD = A_ROWS>
struct EpiWarps {
    // (tried for the residual + max-pool epilogues of the producer-fed modes too: slower — they wait for residual tiles, not
    // for issue slots, and 104 registers leave one tile in flight per warp instead of two.)
    // A_TMA has no producer warps and its residual comes from shared memory: the epilogue is then a dependent chain of
    // TMEM load -> arithmetic -> store per 32-column block with ONE warp per scheduler and nothing to hide its latency
    // behind (measured: 14 k cycles per 128 x 256 tile against 1 k cycles of MMAs), so it always runs two warps per quadrant.
    static constexpr int value = (ALOAD == A_TMA || (CG == 2 && EPI == 5)) ? 8 : 4;
};

// RES: the epilogue adds a residual (E_ADD_LN, E_ADD_MAXPOOL).  With the TMA-fed fp16 mode (A_TMA) the residual is staged
// too: a ring of kResStages x {hi, lo} fp16 boxes of 128 rows x 64 columns (SWIZZLE_128B), filled by a dedicated warp while the
// main loop of the same item still runs, so the epilogue reads it from shared memory at shared-memory latency.
constexpr int kMaxSmem = 227 * 1024;   // opt-in dynamic shared memory per CTA on sm_100
constexpr int kResStages = 2;
constexpr int kResStageBytes = 2 * A_TILE_BYTES;     // hi + lo box: 128 rows x 64 fp16 columns each
constexpr int kBarBytes = 512;
template <int NPASS, int BN, int CG, int EW = 4, int ALOAD = A_ROWS, bool RES = false>
struct Cfg {
    static constexpr int PARTS = parts_of(NPASS);
    static constexpr int W_TILE_BYTES = BN * 128 / CG;                     // one of {hi, lo}, per CTA
    static constexpr int STAGE_BYTES = PARTS * (A_TILE_BYTES + W_TILE_BYTES);
    static constexpr bool RES_TMA = RES && ALOAD == A_TMA;
    static constexpr int RES_BYTES = RES_TMA ? kResStages * kResStageBytes : 0;
    // BN = 192 is the attention epilogue (one head of [Q | K | V], head_dim 64, per chunk): + a V staging tile per warp
    static constexpr int ATTN_BYTES = BN == 192 ? 4 * 32 * kAttnVP * 4 : 0;
    static constexpr int BIAS_BYTES = EW * 512 * 4;     // per-warp bias staging (an item has at most 512 columns)
    static constexpr int MISC_BYTES = 1024 /*align slack*/ + kBarBytes + BIAS_BYTES + 2 * 512 * 4 /*gamma, beta*/ +
                                      2 * 256 * 4 /*LayerNorm partial sums of the two warps of a quadrant*/;
    // rings get 200 KB (A_TMA: minus the residual ring and the attention tile, which the producer-fed modes fit on top)
    static constexpr int STAGES_RAW = ALOAD == A_TMA ? (200 * 1024 - ATTN_BYTES - RES_BYTES) / STAGE_BYTES
                                                     : (200 * 1024) / STAGE_BYTES;
    // A-producer groups: group g owns k-atom steps g, g+G, ...  G must divide STAGES so that every smem stage belongs
    // to ONE group, which then sees every phase of that stage's `empty` barrier (a parity wait can only tell two
    // consecutive phases apart).
    // (A_TMA has no producer groups: any depth from 2 to 6 works)
    static constexpr int STAGES = ALOAD == A_TMA ? (STAGES_RAW >= 6 ? 6 : (STAGES_RAW < 2 ? 2 : STAGES_RAW))
                                  : CG == 2      ? (STAGES_RAW >= 6 ? 6 : 3)
                                                 : (STAGES_RAW >= 4 ? 4 : 2);
    static constexpr int GROUPS = CG == 2 ? 3 : STAGES;
    static constexpr int PRODUCER_WARPS = ALOAD == A_TMA ? 0 : (CG == 2 ? 6 : 8);   // A_TMA: the TMA engine is the producer
    static constexpr int EPI_WARPS = EW;
    // warps before the epilogue warps: loader, MMA, producers; A_TMA: loader, MMA, residual loader, one idle warp (so that
    // epilogue warp e sits in TMEM lane quadrant e % 4)
    static constexpr int FIRST_EPI_WARP = ALOAD == A_TMA ? 4 : 2 + PRODUCER_WARPS;
    static constexpr int THREADS = 32 * (FIRST_EPI_WARP + EW);
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFF = RING_BYTES + RES_BYTES;                 // residual ring right behind the main ring (1024-aligned)
    static constexpr int PARAM_OFF = BAR_OFF + kBarBytes;
    static constexpr int SMEM_BYTES = RING_BYTES + RES_BYTES + MISC_BYTES + ATTN_BYTES;
    static_assert(SMEM_BYTES <= kMaxSmem, "shared memory budget");
    static_assert((4 * STAGES + 4 + 2 * kResStages) * 8 + 8 <= kBarBytes, "barrier block");
    static_assert(STAGE_BYTES % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
};


template <int NPASS, int BN, int NCH, int ALOAD, int EPI, int CG>
__global__ void __launch_bounds__(Cfg<NPASS, BN, CG, EpiWarps<EPI, CG, ALOAD>::value, ALOAD, epi_has_res(EPI)>::THREADS, 1)
    tc_gemm_kernel(const __grid_constant__ GemmParams p) {
    using C = Cfg<NPASS, BN, CG, EpiWarps<EPI, CG, ALOAD>::value, ALOAD, epi_has_res(EPI)>;
    static_assert(ALOAD != A_TMA || NPASS == 4, "the TMA-fed A operand is the fp16 single-pass mode");
    constexpr int EW = C::EPI_WARPS;
    constexpr int S = C::STAGES;
    constexpr int kThreads = C::THREADS, kProducerWarps = C::PRODUCER_WARPS;
    // CTA pair: rank 0 (leader) issues the MMAs for both SMs; work items are 256-row tiles, 128 rows per CTA
    const u32 rank = CG == 2 ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const long long item0 = CG == 2 ? (blockIdx.x >> 1) : blockIdx.x;
    const long long item_step = CG == 2 ? (gridDim.x >> 1) : gridDim.x;
    auto row0_of = [&](long long item) { return (item / p.n_groups) * (long long)(BM * CG) + (long long)rank * BM; };
    constexpr int ACC_STAGES = (NCH * BN * 2 <= kTmemCols) ? 2 : 1;
    constexpr int ACC_COLS = kTmemCols / ACC_STAGES;  // column stride between accumulator stages
    // LayerNorm over a 512-column row fills TMEM: no second accumulator stage.  The two 256-column chunks are then handed
    // over one by one (acc_full / acc_empty slot c = chunk c), so the first statistics pass of chunk 0 overlaps the MMAs of
    // chunk 1 and the last normalise pass of chunk 1 overlaps the MMAs of the next tile's chunk 0.
    constexpr bool CHUNKED = EPI == E_ADD_LN && NCH == 2;

    extern __shared__ uint8_t smem_raw[];
    const u32 smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const u32 bar_base = smem_base + C::BAR_OFF;
    // barriers (8 B each): full_w[S], full_a[S], empty[S], acc_full[2], acc_empty[2], peer_w[S]; then the TMEM base address.
    // CTA pair: full_a / acc_empty / peer_w of the LEADER also receive the peer's arrivals (remote mbarrier.arrive);
    // empty / acc_full are signalled in both CTAs by the leader's multicast tcgen05.commit.
    auto full_w = [&](int s) { return bar_base + 8u * s; };
    auto full_a = [&](int s) { return bar_base + 8u * (S + s); };
    auto empty = [&](int s) { return bar_base + 8u * (2 * S + s); };
    auto acc_full = [&](int a) { return bar_base + 8u * (3 * S + a); };
    auto acc_empty = [&](int a) { return bar_base + 8u * (3 * S + 2 + a); };
    auto peer_w = [&](int s) { return bar_base + 8u * (3 * S + 4 + s); };
    auto r_full = [&](int s) { return bar_base + 8u * (4 * S + 4 + s); };               // residual ring (RES_TMA)
    auto r_empty = [&](int s) { return bar_base + 8u * (4 * S + 4 + kResStages + s); };
    const u32 tmem_slot = bar_base + 8u * (4 * S + 4 + 2 * kResStages);
    volatile u32 *tmem_slot_ptr = reinterpret_cast<volatile u32 *>(smem + C::BAR_OFF + 8 * (4 * S + 4 + 2 * kResStages));
    // the residual ring is used when the host built the residual tensor maps (fp16 plane pair, nout % 64 == 0)
    const bool res_tma = C::RES_TMA && p.res_tma != 0;

    auto a_hi = [&](int s) { return smem_base + (u32)s * C::STAGE_BYTES; };
    auto a_lo = [&](int s) { return a_hi(s) + A_TILE_BYTES; };                       // split modes only
    auto w_hi = [&](int s) { return a_hi(s) + C::PARTS * A_TILE_BYTES; };
    auto w_lo = [&](int s) { return w_hi(s) + C::W_TILE_BYTES; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(full_w(s), 1);
            mbar_init(full_a(s), ALOAD == A_TMA ? 1 : CG * (kProducerWarps / C::GROUPS));   // (unused with A_TMA)
            mbar_init(empty(s), 1);
            mbar_init(peer_w(s), 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(acc_full(a), 1);
            mbar_init(acc_empty(a), EW * CG);
        }
        for (int r = 0; r < kResStages; r++) {
            mbar_init(r_full(r), 1);
            mbar_init(r_empty(r), EW);
        }
        fence_barrier_init();
    }
    if (CG == 2) cluster_sync_all();  // both CTAs of the pair resident, barriers initialised before any remote arrive
    if (warp == 1) tmem_alloc<CG>(tmem_slot, kTmemCols);
    if (EPI == E_ADD_LN) {  // LayerNorm scale / shift -> smem once (Nout = NCH * BN <= 512 columns)
        float *sg = reinterpret_cast<float *>(smem + C::PARAM_OFF) + EW * 512;
        for (int i = threadIdx.x; i < NCH * BN; i += kThreads) {
            sg[i] = __ldg(p.gamma + i);
            sg[512 + i] = __ldg(p.beta + i);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const u32 tmem_base = *tmem_slot_ptr;
    const int KA = p.KA;
    const long long n_items = p.n_items;

    // EW = 8: warps 0-7 (loader, MMA, 6 producers) and warps 8-15 (epilogue) are two pairs of warpgroups; each side
    // re-sizes its registers INSIDE its own branch, so that the setmaxnreg dominates the code it governs (ptxas bounds
    // the registers of a region by the setmaxnreg that dominates it) and a warpgroup executes one and the same instruction.
    // (A_TMA has no producer warps: 2 + EW warps, nothing to re-divide)
    constexpr bool kSetMaxNReg = EW == 8 && C::PRODUCER_WARPS == 6;
    static_assert(EW != 8 || C::PRODUCER_WARPS == 6 || C::PRODUCER_WARPS == 0, "register re-division assumes 4 warpgroups");
    if (warp < C::FIRST_EPI_WARP) {
    if constexpr (kSetMaxNReg) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kProdRegs));
    if (warp == 0) {
        // ===================================================================== W loader
        if (lane == 0) {
            Pipe pipe;
            constexpr int PARTS = C::PARTS;
            // bytes this CTA receives per stage (A_TMA: + its own 128-row A tile, on the same barrier)
            constexpr u32 BYTES = PARTS * C::W_TILE_BYTES + (ALOAD == A_TMA ? A_TILE_BYTES : 0);
            constexpr size_t TILE = (size_t)PARTS * BN * 128;        // packed bytes of one (chunk, k-atom) tile
            if (ALOAD == A_TMA) tma_prefetch_desc(&p.tmA);
            constexpr u32 PIECE = C::W_TILE_BYTES < 8192 ? C::W_TILE_BYTES : (C::W_TILE_BYTES % 8192 ? 4096 : 8192);
            for (long long item = item0; item < n_items; item += item_step) {
                const int n_group = (int)(item % p.n_groups);
                const int arow0 = (int)row0_of(item);
                for (int c = 0; c < NCH; c++) {
                    const int chunk = n_group * NCH + c;
                    const uint8_t *src = reinterpret_cast<const uint8_t *>(p.Wp) + (size_t)chunk * KA * TILE;
                    for (int ka = 0; ka < KA; ka++) {
                        mbar_wait(empty(pipe.stage), pipe.phase ^ 1);
                        mbar_arrive_expect_tx(full_w(pipe.stage), BYTES);
                        if (ALOAD == A_TMA) tma_load_2d(a_hi(pipe.stage), &p.tmA, ka * 64, arow0, full_w(pipe.stage));
                        // several smaller bulk copies: more requests in flight per SM than one big copy.  CTA pair:
                        // this CTA takes rows [rank * BN/2, +BN/2) of the hi and of the lo image.
#pragma unroll
                        for (int part = 0; part < PARTS; part++) {
                            const uint8_t *ps = src + (size_t)ka * TILE + (size_t)part * BN * 128 + (size_t)rank * C::W_TILE_BYTES;
#pragma unroll
                            for (u32 o = 0; o < C::W_TILE_BYTES; o += PIECE)
                                bulk_g2s(w_hi(pipe.stage) + part * C::W_TILE_BYTES + o, ps + o, PIECE, full_w(pipe.stage));
                        }
                        pipe.advance<S>();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (leader) / W relay (peer)
        if (CG == 2 && !leader) {
            // the peer's weight half lands on the peer's own full_w barrier (a bulk copy signals a barrier of the CTA
            // it writes to); one thread forwards each completion to the leader, in stage order
            if (lane == 0) {
                Pipe pipe;
                for (long long item = item0; item < n_items; item += item_step)
                    for (int st = 0; st < NCH * KA; st++) {
                        mbar_wait(full_w(pipe.stage), pipe.phase);
                        mbar_arrive_remote(peer_w(pipe.stage), 0);
                        pipe.advance<S>();
                    }
            }
        } else if (lane == 0) {
            Pipe pipe;
            int as = 0;
            u32 aphase = 0;
            constexpr u32 idesc = umma_idesc<BN, NPASS == 2 ? 1 : (NPASS == 4 ? 0 : 2), CG>();   // bf16 / fp16 / tf32 operands
            constexpr int BKE = bk_of(NPASS), KSTEP = BKE / 4;               // elements per k-atom / per MMA (32 bytes of K)
            for (long long item = item0; item < n_items; item += item_step) {
                if (!CHUNKED) {
                    if (CG == 2) mbar_wait_cluster(acc_empty(as), aphase ^ 1);
                    else mbar_wait(acc_empty(as), aphase ^ 1);  // epilogue has drained this accumulator stage
                    tc_fence_after();
                }
                for (int c = 0; c < NCH; c++) {
                    if (CHUNKED) {
                        mbar_wait(acc_empty(c), aphase ^ 1);    // epilogue has drained chunk c of the previous tile
                        tc_fence_after();
                    }
                    const u32 d = tmem_base + (u32)(as * ACC_COLS + c * BN);
                    for (int ka = 0; ka < KA; ka++) {
                        mbar_wait(full_w(pipe.stage), pipe.phase);   // (A_TMA: this CTA's A tile arrives on it too)
                        if (CG == 2) {
                            if (ALOAD != A_TMA) mbar_wait_cluster(full_a(pipe.stage), pipe.phase);   // both CTAs' A tiles
                            mbar_wait_cluster(peer_w(pipe.stage), pipe.phase);   // the peer's W half (A_TMA: and A tile)
                        } else if (ALOAD != A_TMA) {
                            mbar_wait(full_a(pipe.stage), pipe.phase);
                        }
                        tc_fence_after();
                        // k-steps of 8 inside the atom; the last atom may be partially filled (zero padded)
                        const int ksteps = (ka == KA - 1) ? ((p.K - ka * BKE + KSTEP - 1) / KSTEP) : (BKE / KSTEP);
                        for (int kk = 0; kk < ksteps; kk++) {
                            const u32 acc = (ka | kk) ? 1u : 0u;
                            const u64 ah = umma_desc(a_hi(pipe.stage) + kk * 32);
                            const u64 wh = umma_desc(w_hi(pipe.stage) + kk * 32);
                            if (NPASS == 3) {
                                const u64 al = umma_desc(a_lo(pipe.stage) + kk * 32);
                                const u64 wl = umma_desc(w_lo(pipe.stage) + kk * 32);
                                umma_tf32<CG>(d, ah, wl, idesc, acc);   // small terms first
                                umma_tf32<CG>(d, al, wh, idesc, 1u);
                                umma_tf32<CG>(d, ah, wh, idesc, 1u);
                            } else if (NPASS == 2) {
                                const u64 al = umma_desc(a_lo(pipe.stage) + kk * 32);
                                const u64 wl = umma_desc(w_lo(pipe.stage) + kk * 32);
                                umma_bf16<CG>(d, ah, wl, idesc, acc);
                                umma_bf16<CG>(d, al, wh, idesc, 1u);
                                umma_bf16<CG>(d, ah, wh, idesc, 1u);
                            } else if (NPASS == 4) {
                                umma_bf16<CG>(d, ah, wh, idesc, acc);    // kind::f16, fp16 operands (idesc)
                            } else {
                                umma_tf32<CG>(d, ah, wh, idesc, acc);
                            }
                        }
                        umma_commit<CG>(empty(pipe.stage));  // frees the smem stage (both CTAs) once these MMAs have read it
                        pipe.advance<S>();
                    }
                    if (CHUNKED) umma_commit<CG>(acc_full(c));   // chunk c complete -> epilogue (both CTAs)
                }
                if (!CHUNKED) umma_commit<CG>(acc_full(as));     // accumulator complete -> epilogue (both CTAs)
                if (ACC_STAGES == 2) {
                    as ^= 1;
                    if (as == 0) aphase ^= 1;
                } else {
                    aphase ^= 1;
                }
            }
        }
    } else if constexpr (ALOAD == A_TMA) {
        // ===================================================================== residual loader (warp 2; warp 3 idles)
        // Runs ahead of the epilogue by the depth of the ring: the first boxes of an item land while its MMAs are still
        // being issued.  Consumption order = the epilogue's: 64-column blocks of the item, left to right.
        if constexpr (C::RES_TMA) {
            if (warp == 2 && lane == 0 && res_tma) {
                tma_prefetch_desc(&p.tmRh);
                tma_prefetch_desc(&p.tmRl);
                Pipe rp;
                // (An L2 prefetch of the rest of the item's slab and of the next item's — cp.async.bulk.prefetch.tensor — was
                // measured: no shorter epilogue, and 25-60 % more DRAM reads per launch, the prefetched lines being evicted by
                // the kernel's own output stream before the ring got to them; profiles/r02_ncu_tc_gemm_step_launches.csv.)
                for (long long item = item0; item < n_items; item += item_step) {
                    const int rrow0 = (int)row0_of(item);
                    const int col0 = (int)(item % p.n_groups) * NCH * BN;
                    for (int cb = 0; cb < NCH * BN / 64; cb++) {
                        mbar_wait(r_empty(rp.stage), rp.phase ^ 1);
                        mbar_arrive_expect_tx(r_full(rp.stage), (u32)kResStageBytes);
                        const u32 dst = smem_base + C::RING_BYTES + (u32)rp.stage * kResStageBytes;
                        tma_load_2d(dst, &p.tmRh, col0 + 64 * cb, rrow0, r_full(rp.stage));
                        tma_load_2d(dst + A_TILE_BYTES, &p.tmRl, col0 + 64 * cb, rrow0, r_full(rp.stage));
                        rp.advance<kResStages>();
                    }
                }
            }
        }
    } else {
        // ===================================================================== A producers (G groups of 8/G warps)
        // Group g owns k-atom steps g, g+G, g+2G, ... of this CTA's flattened (item, chunk pass, k-atom) sequence and
        // has exactly ONE batch of loads in flight per thread; the G groups together keep G k-atoms (16 KB each) in
        // flight per SM, loaded before their smem stage is free.  (A register ring inside one warp
        // does not work: ptxas folds the ring's loads onto shared scoreboard slots, so waiting for the oldest batch
        // waits for the newest.)  Thread -> (16-byte chunk of the 128-byte row, R rows RSTEP apart): a warp-wide access
        // covers 4 rows x 128 B, coalesced in global memory and conflict-free in the swizzled tile.
        constexpr int G = C::GROUPS;                      // 2 or 4
        constexpr int WPG = kProducerWarps / G;           // warps per group
        constexpr int RSTEP = 4 * WPG;                    // rows covered by one group-wide access
        constexpr int R = BM / RSTEP;                     // rows per thread
        constexpr int V = BK_IS_64(NPASS) ? 2 : 1;        // float4 loads per row: 16-bit modes pack 8 fp32 -> one 16 B chunk
        constexpr int BKE = bk_of(NPASS);                 // elements per k-atom
        const int pw = warp - 2;
        const int g = pw / WPG, h = pw % WPG;
        const int chunk = lane & 7;                       // 16-byte chunk inside the 128-byte smem row
        const int r0 = h * 4 + (lane >> 3);               // rows r0 + RSTEP i, i = 0..R-1
        float4 v[R][V];
        int src_row[R];                       // A_GATHER: flat source point row (b*Nsrc + i) per owned row, -1 = padding
        long long meta_item = -1;

        const long long my_items = n_items > item0 ? (n_items - item0 + item_step - 1) / item_step : 0;
        const int steps_per_item = NCH * KA;
        const long long total = my_items * steps_per_item;

        for (long long step = g; step < total; step += G) {
            const long long li = step / steps_per_item;
            const int ka = (int)(step - li * steps_per_item) % KA;
            const long long item = item0 + li * item_step;
            const long long m0 = row0_of(item);
            const int k = ka * BKE + chunk * 4 * V;       // first fp32 column this thread converts
            if (ALOAD == A_ROWS) {
#pragma unroll
                for (int i = 0; i < R; i++) {
                    const long long t = m0 + r0 + RSTEP * i;
#pragma unroll
                    for (int u = 0; u < V; u++) {
                        v[i][u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (t < p.T && k + 4 * u < p.K)
                            v[i][u] = __ldg(reinterpret_cast<const float4 *>(p.A + t * p.lda + k + 4 * u));
                    }
                }
            } else {
                if (item != meta_item) {
                    meta_item = item;
                    // scene of the tile's first row: ONE 64-bit division per item (it used to be one per row — sixteen
                    // emulated divisions per thread and item were 44 % of the producer warps' samples); a 128-row tile
                    // crosses a scene boundary at most every M * ns rows, walked with a compare
                    const long long rps = (long long)p.M * p.ns;
                    const long long b0 = m0 / rps;
                    const long long next0 = (b0 + 1) * rps;        // first row of the next scene
#pragma unroll
                    for (int i = 0; i < R; i++) {
                        const long long t = m0 + r0 + RSTEP * i;
                        int b = (int)b0;
                        for (long long nx = next0; t >= nx; nx += rps) b++;   // rps < 128 only for toy shapes
                        src_row[i] = t < p.T ? b * p.Nsrc + __ldg(p.idx + t) : -1;
                    }
                }
#pragma unroll
                for (int i = 0; i < R; i++) {
#pragma unroll
                    for (int u = 0; u < V; u++) {
                        const int kc = k + 4 * u;
                        v[i][u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (src_row[i] >= 0) {
                            if (kc < p.C) {
                                v[i][u] = __ldg(reinterpret_cast<const float4 *>(p.feat_t + (size_t)src_row[i] * p.C + kc));
                            } else if (kc == p.C) {  // the three centred coordinates follow the C features
                                const long long grp = (m0 + r0 + RSTEP * i) / p.ns;  // flat (b, centre)
                                const float *q = p.xyz + (size_t)src_row[i] * 3;
                                const float *c = p.new_xyz + grp * 3;
                                v[i][u] = make_float4(__ldg(q) - __ldg(c), __ldg(q + 1) - __ldg(c + 1),
                                                      __ldg(q + 2) - __ldg(c + 2), 0.f);
                            }
                        }
                    }
                }
            }
            const int stage = (int)(step % S);
            const u32 phase = (u32)((step / S) & 1);
            mbar_wait(empty(stage), phase ^ 1);
            uint8_t *ah = smem + (size_t)stage * C::STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < R; i++) {
                const int r = r0 + RSTEP * i;
                const u32 off = (u32)r * 128u + (u32)((chunk ^ (r & 7)) << 4);
                if (NPASS == 3) {
                    const float4 x = v[i][0];
                    const float4 hh = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
                    *reinterpret_cast<float4 *>(ah + off) = hh;
                    *reinterpret_cast<float4 *>(ah + A_TILE_BYTES + off) =
                        make_float4(x.x - hh.x, x.y - hh.y, x.z - hh.z, x.w - hh.w);
                } else if (NPASS == 2) {
                    uint4 hi, lo;
                    bf16_split8(v[i][0], v[i][V - 1], hi, lo);
                    *reinterpret_cast<uint4 *>(ah + off) = hi;
                    *reinterpret_cast<uint4 *>(ah + A_TILE_BYTES + off) = lo;
                } else if (NPASS == 4) {
                    *reinterpret_cast<uint4 *>(ah + off) = f16_pack8(v[i][0], v[i][V - 1]);
                } else {
                    const float4 x = v[i][0];
                    *reinterpret_cast<float4 *>(ah + off) =
                        make_float4(tf32_rna(x.x), tf32_rna(x.y), tf32_rna(x.z), tf32_rna(x.w));
                }
            }
            fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) {
                if (CG == 2 && !leader) mbar_arrive_remote(full_a(stage), 0);  // the leader's MMA thread waits for both tiles
                else mbar_arrive(full_a(stage));
            }
        }
    }
    } else {
        if constexpr (kSetMaxNReg) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
        // ===================================================================== epilogue (128 threads)
        // The accumulator is read with the 16x256b TMEM load shape, which hands out an MMA-C-fragment layout: for a
        // 32-row x 32-column block, lane (fr = lane / 4, fc = 2 * (lane % 4)) holds rows 16 h + 8 j + fr (h, j in {0,1})
        // and columns 8 k + fc + {0, 1} (k = 0..3).  A quad of lanes therefore owns 32 contiguous bytes of a row: every
        // global load / store instruction moves whole 32-byte sectors with no shared-memory transpose, a row statistic
        // is two shuffles inside the quad, and a maximum over the rows of a neighbourhood is three shuffles.  (The first
        // version used the 32x32b shape — thread = row — and staged every block through shared memory twice; it was
        // instruction- and latency-bound: profiles/r01_ncu_tc_gemm_*.)
        const int q = warp & 3;                  // TMEM lane quadrant this warp may read
        const int ew = warp - C::FIRST_EPI_WARP;      // 0..EW-1 (EW = 8: ew >> 2 = the warp's m-tile)
        const int fr = lane >> 2, fc = (lane & 3) * 2;
        // Per-column parameters live in shared memory: with ~200 KB of it carved out the L1 is a few KB and thrashed by
        // the A stream, so an __ldg of bias / gamma / beta inside the block loop was an L2 round trip on the critical path.
        // The bias of ALL column groups is staged once per kernel when it fits the EW x 512 floats set aside for it (every
        // shape of the model); an item then only moves a pointer.  (It used to be re-read per item by each warp, "while the
        // MMAs still run" — but a rolled loop of dependent LDG -> STS pairs is ~6 L2 round trips, and in the epilogue-bound
        // kernels (attention) nothing was waiting for the MMAs: 12 % of the kernel's warp samples sat on that one STS.)
        float *const sbias_all = reinterpret_cast<float *>(smem + C::PARAM_OFF);
        const bool bias_resident = p.n_groups * NCH * BN <= EW * 512;
        const float *sbias = sbias_all + ew * 512;
        if (bias_resident) {
            for (int i = ew * 32 + lane; i < p.n_groups * NCH * BN; i += EW * 32)
                sbias_all[i] = (p.bias && i < p.Nout) ? __ldg(p.bias + i) : 0.f;
            asm volatile("bar.sync 8, %0;" ::"n"(EW * 32) : "memory");   // the epilogue warps only
        }
        const float *sgamma = reinterpret_cast<const float *>(smem + C::PARAM_OFF) + EW * 512;
        const float *sbeta = sgamma + 512;
        float *sstat = const_cast<float *>(sbeta) + 512;   // [2 exchanges][4 quadrants][2 halves][32 rows]
        int as = 0;
        u32 aphase = 0;

        // v[h][k*4 + j*2 + e]  <->  row 16h + 8j + fr, column 8k + fc + e   (register order of tcgen05.ld.16x256b.x4)
        auto frag_ld = [&](u32 taddr, float (&v)[2][16]) {
            tmem_ld_16x256b_x4(taddr, v[0]);
            tmem_ld_16x256b_x4(taddr + (16u << 16), v[1]);
            tmem_wait_ld();
        };
        auto frag_st = [&](u32 taddr, const float (&v)[2][16]) {
            tmem_st_16x256b_x4(taddr, v[0]);
            tmem_st_16x256b_x4(taddr + (16u << 16), v[1]);
            tmem_wait_st();
        };
        // columns [c0, c0+32) of this item (c0 relative to the item's first column): bias staged per item below
        auto add_bias = [&](float (&v)[2][16], int c0) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float2 b2 = *reinterpret_cast<const float2 *>(sbias + c0 + 8 * k + fc);
#pragma unroll
                for (int h = 0; h < 2; h++)
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        v[h][k * 4 + j * 2] += b2.x;
                        v[h][k * 4 + j * 2 + 1] += b2.y;
                    }
            }
        };
        // global rows [grow0, grow0+32) x columns [n0, n0+32) in the fragment layout; zeros outside the matrix.
        // res16 (NPASS = 4): the residual is a (hi, lo) pair of fp16 planes; a tile register then carries the two packed
        // half2 words (.x = hi pair, .y = lo pair) and add_tile adds hi + lo — the same 8 bytes per element pair in flight.
        const bool res16 = NPASS == 4 && p.res16 != 0;
        const int out16 = NPASS == 4 ? p.out16 : 0;
        auto issue_tile = [&](float2 (&t)[2][8], const float *base, int ld, long long grow0, long long nrows, int n0) {
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const long long r = grow0 + 16 * h + 8 * j + fr;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int col = n0 + 8 * k + fc;
                        t[h][k * 2 + j] = make_float2(0.f, 0.f);
                        if (r < nrows && col < p.Nout) {
                            if (res16) {
                                const size_t o = (size_t)r * ld + col;
                                t[h][k * 2 + j] = make_float2(
                                    __uint_as_float(__ldg(reinterpret_cast<const u32 *>(reinterpret_cast<const __half *>(base) + o))),
                                    __uint_as_float(__ldg(reinterpret_cast<const u32 *>(reinterpret_cast<const __half *>(p.R_lo) + o))));
                            } else {
                                t[h][k * 2 + j] = __ldg(reinterpret_cast<const float2 *>(base + r * ld + col));
                            }
                        }
                    }
                }
        };
        auto add_tile = [&](float (&v)[2][16], const float2 (&t)[2][8]) {
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float2 r = t[h][k * 2 + j];
                        if (res16) {
                            const float2 a = f16x2_to_float2(__float_as_uint(r.x)), b = f16x2_to_float2(__float_as_uint(r.y));
                            r = make_float2(a.x + b.x, a.y + b.y);
                        }
                        v[h][k * 4 + j * 2] += r.x;
                        v[h][k * 4 + j * 2 + 1] += r.y;
                    }
        };
        // out16 = 1: fp16 rows; 2: (hi, lo) fp16 planes (p.out, p.out_lo); 0: fp32 rows
        // One row pointer per (h, j) row and plane, the four column pairs at constant offsets from it; the column bound is
        // tested once per block (every shape of the model has Nout % 32 == 0).  (The first version rebuilt a 64-bit address
        // and re-tested both bounds for each of the 16 stores: address arithmetic and branches were half of the
        // instructions of the LayerNorm kernel's last pass.)
        auto store_tile = [&](const float (&v)[2][16], float *base, int ld, long long grow0, long long nrows, int n0) {
            const bool full = n0 + 32 <= p.Nout;
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const long long r = grow0 + 16 * h + 8 * j + fr;
                    if (r >= nrows) continue;
                    const size_t o = (size_t)r * ld + n0 + fc;
                    if (out16 == 0) {
                        float *rowp = base + o;
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (full || n0 + 8 * k + fc < p.Nout)
                                *reinterpret_cast<float2 *>(rowp + 8 * k) = make_float2(v[h][k * 4 + j * 2], v[h][k * 4 + j * 2 + 1]);
                    } else if (out16 == 1) {
                        __half *rowp = reinterpret_cast<__half *>(base) + o;
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (full || n0 + 8 * k + fc < p.Nout)
                                *reinterpret_cast<u32 *>(rowp + 8 * k) = f16x2(v[h][k * 4 + j * 2], v[h][k * 4 + j * 2 + 1]);
                    } else {
                        __half *rowh = reinterpret_cast<__half *>(base) + o, *rowl = reinterpret_cast<__half *>(p.out_lo) + o;
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (full || n0 + 8 * k + fc < p.Nout) {
                                u32 hi, lo;
                                f16_split2(v[h][k * 4 + j * 2], v[h][k * 4 + j * 2 + 1], hi, lo);
                                *reinterpret_cast<u32 *>(rowh + 8 * k) = hi;
                                *reinterpret_cast<u32 *>(rowl + 8 * k) = lo;
                            }
                    }
                }
        };
        // Staged residual (RES_TMA): the (hi, lo) fp16 boxes of the current 64-column block sit in ring stage rp.stage as
        // SWIZZLE_128B tiles (row pitch 128 B, 16-byte chunk c of row r at slot c ^ (r & 7)); the 8 rows x 4 lanes of a warp-wide
        // 4-byte read then cover all 32 banks.  `sub` = which 32-column half of the box.  The first read of a box waits for
        // its TMA, the second half's read releases the stage to the loader warp.
        Pipe rp;
        auto add_res_smem = [&](float (&v)[2][16], int sub) {
            // one warp per quadrant walks both halves of a box; with two warps per quadrant each owns one half
            constexpr bool kBothHalves = EW == 4;
            if (!kBothHalves || sub == 0) mbar_wait(r_full(rp.stage), rp.phase);
            const uint8_t *rs = smem + C::RING_BYTES + (size_t)rp.stage * kResStageBytes + (size_t)(q * 32) * 128;
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const uint8_t *rowp = rs + (16 * h + 8 * j + fr) * 128 + fc * 2;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int off = ((sub * 4 + k) ^ fr) << 4;
                        const float2 a = f16x2_to_float2(*reinterpret_cast<const u32 *>(rowp + off));
                        const float2 b = f16x2_to_float2(*reinterpret_cast<const u32 *>(rowp + A_TILE_BYTES + off));
                        v[h][k * 4 + j * 2] += a.x + b.x;
                        v[h][k * 4 + j * 2 + 1] += a.y + b.y;
                    }
                }
            if (!kBothHalves || sub == 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(r_empty(rp.stage));
                rp.advance<kResStages>();
            }
        };
        // max over the ns (16 or 32) consecutive rows of each neighbourhood -> out[group, n0 + ...]
        auto pool_store = [&](const float (&v)[2][16], long long grow0, int n0, int ns) {
            float m[2][8];
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        float x = fmaxf(v[h][k * 4 + e], v[h][k * 4 + 2 + e]);       // rows fr and fr + 8 of half h
                        x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 4));
                        x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 8));
                        x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 16));
                        m[h][k * 2 + e] = x;
                    }
            if (fr == 0) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int col = n0 + 8 * k + fc;
                    if (col >= p.Nout) continue;
                    if (ns == 64) {
                        // a neighbourhood spans two quadrants (two warps): combine through atomicMax on the bit patterns —
                        // E_RELU_MAXPOOL only (values >= 0 order like their int bits), `out` zeroed by the host wrapper
                        if (grow0 < p.T) {
                            int *o = reinterpret_cast<int *>(p.out + (grow0 / 64) * p.ldo + col);
                            atomicMax(o, __float_as_int(fmaxf(m[0][k * 2], m[1][k * 2])));
                            atomicMax(o + 1, __float_as_int(fmaxf(m[0][k * 2 + 1], m[1][k * 2 + 1])));
                        }
                    } else if (ns == 32) {
                        if (grow0 < p.T)
                            *reinterpret_cast<float2 *>(p.out + (grow0 / 32) * p.ldo + col) =
                                make_float2(fmaxf(m[0][k * 2], m[1][k * 2]), fmaxf(m[0][k * 2 + 1], m[1][k * 2 + 1]));
                    } else {
                        if (grow0 < p.T)
                            *reinterpret_cast<float2 *>(p.out + (grow0 / 16) * p.ldo + col) =
                                make_float2(m[0][k * 2], m[0][k * 2 + 1]);
                        if (grow0 + 16 < p.T)
                            *reinterpret_cast<float2 *>(p.out + (grow0 / 16 + 1) * p.ldo + col) =
                                make_float2(m[1][k * 2], m[1][k * 2 + 1]);
                    }
                }
            }
        };

        for (long long item = item0; item < n_items; item += item_step) {
            const long long m0 = row0_of(item);
            const int n_group = (int)(item % p.n_groups);
            const long long wrow0 = m0 + q * 32;     // first row of this warp's 32-row slab
            if (bias_resident) {
                sbias = sbias_all + n_group * NCH * BN;
            } else {   // this item's bias -> the warp's own slot: independent loads first, then the stores
                __syncwarp();
                constexpr int NBI = (NCH * BN + 31) / 32;
                float bv[NBI];
#pragma unroll
                for (int u = 0; u < NBI; u++) {
                    const int n = n_group * NCH * BN + lane + 32 * u;
                    bv[u] = (p.bias && lane + 32 * u < NCH * BN && n < p.Nout) ? __ldg(p.bias + n) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < NBI; u++)
                    if (lane + 32 * u < NCH * BN) sbias_all[ew * 512 + lane + 32 * u] = bv[u];
                __syncwarp();
            }
            // residual tiles: two register sets, each refilled as soon as it has been consumed, so a tile has ~1.5 column
            // blocks (and, for the first two, the whole accumulator wait) to arrive from L2 / HBM
            // (EW = 8: the two warps of a quadrant take the even / odd 32-column blocks, one register set each)
            constexpr int NW = EW / 4;                       // warps per quadrant
            const int half = EW == 8 ? (ew >> 2) : 0;        // which of them this warp is
            float2 ra[2][8], rb[2][8];
            if ((EPI == E_ADD_MAXPOOL || EPI == E_ADD_LN) && !res_tma) {
                const int nb0 = n_group * NCH * BN;
                // pull this warp's whole residual slab (32 rows x NCH * BN columns) towards L2 while the MMAs still run:
                // the register tiles below are then fed at L2 latency, not DRAM latency
                if (wrow0 + lane < p.T) {
                    if (res16) {   // two fp16 planes: 128 B = 64 columns of each
                        const __half *rh = reinterpret_cast<const __half *>(p.R) + (wrow0 + lane) * p.ldr + nb0;
                        const __half *rl = reinterpret_cast<const __half *>(p.R_lo) + (wrow0 + lane) * p.ldr + nb0;
#pragma unroll
                        for (int c = 0; c < NCH * BN; c += 64)
                            if (nb0 + c < p.Nout) {
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(rh + c));
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(rl + c));
                            }
                    } else {
                        const float *rrow = p.R + (wrow0 + lane) * p.ldr + nb0;
#pragma unroll
                        for (int c = 0; c < NCH * BN; c += 32)
                            if (nb0 + c < p.Nout) asm volatile("prefetch.global.L2 [%0];" ::"l"(rrow + c));
                    }
                }
                issue_tile(ra, p.R, p.ldr, wrow0, p.T, nb0 + 32 * half);
                if (NW == 1) issue_tile(rb, p.R, p.ldr, wrow0, p.T, nb0 + 32);
            }
            if (!CHUNKED) {
                mbar_wait(acc_full(as), aphase);
                tc_fence_after();
            }
            const u32 tacc = tmem_base + ((u32)(q * 32) << 16) + (u32)(as * ACC_COLS);

            if (EPI == E_STORE || EPI == E_RELU) {
                for (int c = 0; c < NCH; c++) {
                    for (int j = 32 * half; j < BN; j += 32 * NW) {
                        const int n0 = (n_group * NCH + c) * BN + j;
                        if (n0 >= p.Nout) break;
                        float v[2][16];
                        frag_ld(tacc + c * BN + j, v);
                        add_bias(v, c * BN + j);
                        if (EPI == E_RELU) {
#pragma unroll
                            for (int e = 0; e < 16; e++) {
                                v[0][e] = fmaxf(v[0][e], 0.f);
                                v[1][e] = fmaxf(v[1][e], 0.f);
                            }
                        }
                        store_tile(v, p.out, p.ldo, wrow0, p.T, n0);
                    }
                }
            } else if (EPI == E_ADD_MAXPOOL || EPI == E_RELU_MAXPOOL) {
                // rows of a tile beyond T cannot occur inside a neighbourhood (T % ns == 0): whole groups are skipped
                const int ns = p.ns;
                const int nblk = NCH * (BN / 32);
                auto n0_of = [&](int b) { return (n_group * NCH + b / (BN / 32)) * BN + (b % (BN / 32)) * 32; };
                auto block = [&](int b, float2 (&r)[2][8]) {
                    const int n0 = n0_of(b);
                    if (n0 >= p.Nout) return;
                    float v[2][16];
                    frag_ld(tacc + (b / (BN / 32)) * BN + (b % (BN / 32)) * 32, v);
                    add_bias(v, b * 32);
                    if (EPI == E_ADD_MAXPOOL) {
                        if (res_tma) {
                            add_res_smem(v, b & 1);
                        } else {
                            add_tile(v, r);  // r is dead now: refill it with the tile two blocks ahead
                            if (b + 2 < nblk) issue_tile(r, p.R, p.ldr, wrow0, p.T, n0_of(b + 2));
                        }
                    }
                    if (EPI == E_RELU_MAXPOOL) {
#pragma unroll
                        for (int e = 0; e < 16; e++) {
                            v[0][e] = fmaxf(v[0][e], 0.f);
                            v[1][e] = fmaxf(v[1][e], 0.f);
                        }
                    }
                    pool_store(v, wrow0, n0, ns);
                };
                if (NW == 1) {
                    for (int b = 0; b < nblk; b += 2) {
                        block(b, ra);
                        block(b + 1, rb);
                    }
                } else {
                    for (int b = half; b < nblk; b += 2) block(b, ra);
                }
            } else if (EPI == E_ADD_LN) {
                // out = LayerNorm(acc + bias + R) over the full row of E = NCH * BN columns (n_groups == 1).
                // Pass 1 parks v = acc + bias + R back in TMEM and sums it; pass 2: centred variance; pass 3: normalise.
                // A thread owns 4 rows (h, j); a row's 8 values per block are summed in the thread, then over the quad.
                // (Software-pipelining the passes over two fragment buffers — TMEM load of block b + 1 in flight while
                // block b is processed — was measured slower: the extra 32 registers spill.)
                constexpr int E = NCH * BN;
                // NW = 2: the two warps of a quadrant take alternate 32-column blocks (= the two halves of a staged residual
                // box) and add up their partial row statistics through shared memory (one named barrier per exchange).
                auto exchange = [&](float (&part)[4], int slot) {   // quad-reduced partials of rows (h, j) -> totals
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        part[i] += __shfl_xor_sync(0xffffffffu, part[i], 1);
                        part[i] += __shfl_xor_sync(0xffffffffu, part[i], 2);
                    }
                    if (NW == 2) {
                        float *mine = sstat + ((slot * 4 + q) * 2 + half) * 32;
                        const float *other = sstat + ((slot * 4 + q) * 2 + (half ^ 1)) * 32;
                        if ((lane & 3) == 0) {
#pragma unroll
                            for (int i = 0; i < 4; i++) mine[16 * (i >> 1) + 8 * (i & 1) + fr] = part[i];
                        }
                        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
#pragma unroll
                        for (int i = 0; i < 4; i++) part[i] += other[16 * (i >> 1) + 8 * (i & 1) + fr];
                    }
                };
                float sum[4] = {0.f, 0.f, 0.f, 0.f};
                auto block = [&](int j0, float2 (&r)[2][8]) {
                    float v[2][16];
                    frag_ld(tacc + j0, v);
                    add_bias(v, j0);
                    if (res_tma) {
                        add_res_smem(v, (j0 >> 5) & 1);
                    } else {
                        add_tile(v, r);  // r is dead now: refill it with this warp's next tile
                        if (j0 + 64 < E) issue_tile(r, p.R, p.ldr, wrow0, p.T, j0 + 64);
                    }
#pragma unroll
                    for (int h = 0; h < 2; h++)
#pragma unroll
                        for (int j = 0; j < 2; j++)
#pragma unroll
                            for (int k = 0; k < 4; k++) sum[h * 2 + j] += v[h][k * 4 + j * 2] + v[h][k * 4 + j * 2 + 1];
                    frag_st(tacc + j0, v);
                };
                for (int j0 = 0; j0 < E; j0 += 64) {
                    if (CHUNKED && j0 % BN == 0) {          // chunk j0 / BN has just been completed by the MMAs
                        mbar_wait(acc_full(j0 / BN), aphase);
                        tc_fence_after();
                    }
                    if (NW == 1) {
                        block(j0, ra);
                        block(j0 + 32, rb);
                    } else {
                        block(j0 + 32 * half, ra);
                    }
                }
                exchange(sum, 0);
                float mean[4], rstd[4], sq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 4; i++) mean[i] = sum[i] * (1.0f / E);
                for (int j0 = 32 * half; j0 < E; j0 += 32 * NW) {
                    float v[2][16];
                    frag_ld(tacc + j0, v);
#pragma unroll
                    for (int h = 0; h < 2; h++)
#pragma unroll
                        for (int j = 0; j < 2; j++)
#pragma unroll
                            for (int k = 0; k < 4; k++)
#pragma unroll
                                for (int e = 0; e < 2; e++) {
                                    const float d = v[h][k * 4 + j * 2 + e] - mean[h * 2 + j];
                                    sq[h * 2 + j] = fmaf(d, d, sq[h * 2 + j]);
                                }
                }
                exchange(sq, 1);
#pragma unroll
                for (int i = 0; i < 4; i++) rstd[i] = rsqrtf(sq[i] * (1.0f / E) + p.eps);
                for (int j0 = 32 * half; j0 < E; j0 += 32 * NW) {
                    float v[2][16];
                    frag_ld(tacc + j0, v);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float2 g2 = *reinterpret_cast<const float2 *>(sgamma + j0 + 8 * k + fc);
                        const float2 b2 = *reinterpret_cast<const float2 *>(sbeta + j0 + 8 * k + fc);
#pragma unroll
                        for (int h = 0; h < 2; h++)
#pragma unroll
                            for (int j = 0; j < 2; j++) {
                                const int i = h * 2 + j;
                                v[h][k * 4 + j * 2] = (v[h][k * 4 + j * 2] - mean[i]) * rstd[i] * g2.x + b2.x;
                                v[h][k * 4 + j * 2 + 1] = (v[h][k * 4 + j * 2 + 1] - mean[i]) * rstd[i] * g2.y + b2.y;
                            }
                    }
                    store_tile(v, p.out, p.ldo, wrow0, p.T, j0);
                    // this warp's last block of a chunk: hand the chunk back to the MMA warp (acc_empty counts every warp)
                    if (CHUNKED && (j0 + 32 * NW) % BN == 32 * half) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 2 && !leader) mbar_arrive_remote(acc_empty(j0 / BN), 0);
                            else mbar_arrive(acc_empty(j0 / BN));
                        }
                    }
                }
            } else if (EPI == E_ATTN) {
                // One attention head per item: accumulator columns [0, HD) = Q, [HD, 2 HD) = K, [2 HD, 3 HD) = V of the
                // quadrant's 32 tokens = one neighbourhood of 32 or two of 16 (the in_proj weight rows are permuted head
                // by head on the host).  ctx = softmax(Q K^T / sqrt(HD)) V is computed here and only ctx (rows, heads * HD)
                // is stored: the (rows, 3E) qkv matrix never exists.  Both contractions are 32 x 32 x HD per quadrant —
                // far below a tcgen05 tile — so they run on mma.sync m16n8k8 TF32 with the 3x hi/lo compensation, fed
                // straight from the TMEM fragments: the 16x256b load hands out C fragments (row g: columns 2t, 2t+1),
                // which ARE A fragments (row g: k = t, t + 4) and B fragments (column n = g: k = t, t + 4) of the next
                // MMA under the k-permutation t -> 2t, t + 4 -> 2t + 1, applied to both operands.  Only V, whose rows
                // become the contraction index, goes through a shared-memory tile per quadrant.
                // A warp owns MTW 16-row m-tiles of its quadrant: both (EW = 4) or one (EW = 8: two warps per quadrant).
                // The k-bias is dropped (a per-row constant of the scores cancels in the softmax) and the v-bias is added
                // to the output (the rows of P sum to one).
                constexpr int HD = BN / 3;
                static_assert(EPI != E_ATTN || (HD == 64 && NCH == 1), "attention epilogue: head_dim 64, one head per chunk");
                constexpr int MTW = 8 / EW;                      // m-tiles per warp
                const int mt0 = EW == 8 ? (ew >> 2) : 0;         // first m-tile of this warp
                const int ns = p.ns;
                const int g = fr, t2 = fc;  // fragment coordinates: row g, column pair t2 = 2 (lane % 4)
                float *sV = reinterpret_cast<float *>(smem + C::PARAM_OFF) + EW * 512 + 2 * 512 + 2 * 256 + q * 32 * kAttnVP;
                auto split = [](float x, u32 &hi, u32 &lo) {
                    hi = __float_as_uint(x) & 0xffffe000u;
                    lo = __float_as_uint(x - __uint_as_float(hi));
                };
                auto mma = [](float (&d)[4], const u32 (&a)[4], const u32 (&b)[2]) {
                    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                        "{%0,%1,%2,%3};"
                        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
                };
                // NPASS = 2 (split-bf16 GEMM): the attention contractions use the same product class — bf16 hi / lo pairs,
                // m16n8k16, three MMAs per k-step — which halves the mma.sync count.  The legacy mma.sync shares the tensor
                // pipe with the tcgen05 MMAs of the next item and does not overlap with them (measured: the kernel takes
                // main loop + ~18 cycles per HMMA), so the instruction count is what the epilogue costs.  A 16x256b
                // fragment pair (column blocks 2s, 2s + 1) IS the k16 A / B fragment, no permutation needed.
                // NPASS = 4 (fp16 single-pass GEMM): fp16 m16n8k16, ONE MMA per k-step — a third of the instructions again.
                constexpr bool BF = NPASS == 2 || NPASS == 4;
                constexpr bool H1 = NPASS == 4;
                constexpr int PASS0 = H1 ? 2 : 0;     // pass 2 = (a_hi, b_hi); passes 0, 1 = the compensation terms
                auto split2 = [](float x0, float x1, u32 &hi, u32 &lo) {   // (x0, x1) -> packed 16-bit pairs, x0 in the low half
                    if constexpr (H1) {
                        hi = f16x2(x0, x1);
                        lo = 0u;
                    } else {
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
                        const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xffff0000u);
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
                    }
                };
                auto mma16 = [](float (&d)[4], const u32 (&a)[4], const u32 (&b)[2]) {
                    if constexpr (H1)
                        asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                            "{%0,%1,%2,%3};"
                            : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
                    else
                        asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                            "{%0,%1,%2,%3};"
                            : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
                };
                auto pair_sync = [&]() {  // the two warps of a quadrant (EW = 8)
                    if (EW == 8) asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
                    else __syncwarp();
                };
                // FULL: one neighbourhood of 32 rows; otherwise two of 16 (each m-tile attends to its own 16 keys).  Compiled
                // twice so that the MMA sweeps are branch-free and the scheduler can interleave them.
                auto attend = [&](auto full_tag) {
                    constexpr bool FULL = decltype(full_tag)::value;
                    constexpr int NTW = FULL ? 4 : 2;            // key n-tiles per m-tile
                    constexpr int KH = FULL ? 2 : MTW;           // 16-row halves of K this warp reads
                    // ---- scores sc[mi][ni] : rows 16 (mt0 + mi) + {g, g + 8}, keys 8 nt + {t2, t2 + 1}
                    // (two accumulator sets, even / odd k-steps: the tensor pipe is shared with the tcgen05 MMAs of the next
                    // item, a dependent mma.sync waits ~100 cycles, so the epilogue is bound by its dependency chains)
                    float sc[MTW][NTW][4], sc2[MTW][NTW][4];
#pragma unroll
                    for (int mi = 0; mi < MTW; mi++)
#pragma unroll
                        for (int ni = 0; ni < NTW; ni++)
#pragma unroll
                            for (int e = 0; e < 4; e++) sc[mi][ni][e] = sc2[mi][ni][e] = 0.f;
#pragma unroll 1
                    for (int kb = 0; kb < HD; kb += 32) {
                        float qf[MTW][16], kf[KH][16];
#pragma unroll
                        for (int mi = 0; mi < MTW; mi++) tmem_ld_16x256b_x4(tacc + ((u32)(16 * (mt0 + mi)) << 16) + kb, qf[mi]);
#pragma unroll
                        for (int kh = 0; kh < KH; kh++)
                            tmem_ld_16x256b_x4(tacc + ((u32)(16 * (FULL ? kh : mt0 + kh)) << 16) + HD + kb, kf[kh]);
                        tmem_wait_ld();
                        if constexpr (BF) {
#pragma unroll
                            for (int k = 0; k < 4; k += 2) {       // one k16 step = column blocks k, k + 1
                                const float2 bq0 = *reinterpret_cast<const float2 *>(sbias + kb + 8 * k + t2);
                                const float2 bq1 = *reinterpret_cast<const float2 *>(sbias + kb + 8 * k + 8 + t2);
                                u32 ah[MTW][4], al[MTW][4];
#pragma unroll
                                for (int mi = 0; mi < MTW; mi++) {
                                    split2(qf[mi][4 * k + 0] + bq0.x, qf[mi][4 * k + 1] + bq0.y, ah[mi][0], al[mi][0]);
                                    split2(qf[mi][4 * k + 2] + bq0.x, qf[mi][4 * k + 3] + bq0.y, ah[mi][1], al[mi][1]);
                                    split2(qf[mi][4 * k + 4] + bq1.x, qf[mi][4 * k + 5] + bq1.y, ah[mi][2], al[mi][2]);
                                    split2(qf[mi][4 * k + 6] + bq1.x, qf[mi][4 * k + 7] + bq1.y, ah[mi][3], al[mi][3]);
                                }
                                u32 bh[KH][2][2], bl[KH][2][2];
#pragma unroll
                                for (int kh = 0; kh < KH; kh++)
#pragma unroll
                                    for (int e = 0; e < 2; e++) {
                                        split2(kf[kh][4 * k + 2 * e], kf[kh][4 * k + 2 * e + 1], bh[kh][e][0], bl[kh][e][0]);
                                        split2(kf[kh][4 * k + 4 + 2 * e], kf[kh][4 * k + 5 + 2 * e], bh[kh][e][1], bl[kh][e][1]);
                                    }
#pragma unroll
                                for (int pass = PASS0; pass < 3; pass++)
#pragma unroll
                                    for (int ni = 0; ni < NTW; ni++)
#pragma unroll
                                        for (int mi = 0; mi < MTW; mi++) {
                                            const int kh = FULL ? (ni >> 1) : mi;
                                            mma16((k & 2) ? sc2[mi][ni] : sc[mi][ni], pass == 1 ? al[mi] : ah[mi],
                                                  pass == 0 ? bl[kh][ni & 1] : bh[kh][ni & 1]);
                                        }
                            }
                        } else {
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const float2 bq = *reinterpret_cast<const float2 *>(sbias + kb + 8 * k + t2);
                            u32 ah[MTW][4], al[MTW][4];
#pragma unroll
                            for (int mi = 0; mi < MTW; mi++) {
                                split(qf[mi][4 * k + 0] + bq.x, ah[mi][0], al[mi][0]);   // (g,     k = t)
                                split(qf[mi][4 * k + 2] + bq.x, ah[mi][1], al[mi][1]);   // (g + 8, k = t)
                                split(qf[mi][4 * k + 1] + bq.y, ah[mi][2], al[mi][2]);   // (g,     k = t + 4)
                                split(qf[mi][4 * k + 3] + bq.y, ah[mi][3], al[mi][3]);   // (g + 8, k = t + 4)
                            }
                            // B fragments of n-tile (half kh, sub-tile e): rows 16 kh + 8 e + g of K
                            u32 bh[KH][2][2], bl[KH][2][2];
#pragma unroll
                            for (int kh = 0; kh < KH; kh++)
#pragma unroll
                                for (int e = 0; e < 2; e++) {
                                    split(kf[kh][4 * k + 2 * e], bh[kh][e][0], bl[kh][e][0]);
                                    split(kf[kh][4 * k + 2 * e + 1], bh[kh][e][1], bl[kh][e][1]);
                                }
                            // the three compensation terms as three sweeps over the independent accumulators (small
                            // terms first), so that dependent MMAs on one accumulator are several instructions apart
#pragma unroll
                            for (int pass = 0; pass < 3; pass++)
#pragma unroll
                                for (int ni = 0; ni < NTW; ni++)
#pragma unroll
                                    for (int mi = 0; mi < MTW; mi++) {
                                        const int kh = FULL ? (ni >> 1) : mi;
                                        mma((k & 1) ? sc2[mi][ni] : sc[mi][ni], pass == 1 ? al[mi] : ah[mi],
                                            pass == 0 ? bl[kh][ni & 1] : bh[kh][ni & 1]);
                                    }
                        }
                        }
                    }
#pragma unroll
                    for (int mi = 0; mi < MTW; mi++)
#pragma unroll
                        for (int ni = 0; ni < NTW; ni++)
#pragma unroll
                            for (int e = 0; e < 4; e++) sc[mi][ni][e] += sc2[mi][ni][e];
                    // ---- V -> the quadrant's staging tile [token][channel]; then the accumulator stage is free
                    pair_sync();  // (EW = 8) the other warp has finished reading the tile for the previous item
#pragma unroll
                    for (int mi = 0; mi < MTW; mi++)
#pragma unroll
                        for (int cb = 0; cb < HD; cb += 32) {
                            float v[16];
                            tmem_ld_16x256b_x4(tacc + ((u32)(16 * (mt0 + mi)) << 16) + 2 * HD + cb, v);
                            tmem_wait_ld();
#pragma unroll
                            for (int k = 0; k < 4; k++)
#pragma unroll
                                for (int j = 0; j < 2; j++)
                                    *reinterpret_cast<float2 *>(sV + (16 * (mt0 + mi) + 8 * j + g) * kAttnVP + cb + 8 * k + t2) =
                                        make_float2(v[4 * k + 2 * j], v[4 * k + 2 * j + 1]);
                        }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CG == 2 && !leader) mbar_arrive_remote(acc_empty(as), 0);
                        else mbar_arrive(acc_empty(as));
                    }
                    // ---- softmax over the keys of each row (a row's values live in one quad); fp32, full-precision expf.
                    // The head_dim ** -0.5 factor PyTorch applies to q is applied to the scores (one rounding later).
                    constexpr float scaling = 0.125f;
#pragma unroll
                    for (int mi = 0; mi < MTW; mi++)
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            float mx = -3.4e38f;
#pragma unroll
                            for (int ni = 0; ni < NTW; ni++)
#pragma unroll
                                for (int e = 0; e < 2; e++) {
                                    sc[mi][ni][2 * r + e] *= scaling;
                                    mx = fmaxf(mx, sc[mi][ni][2 * r + e]);
                                }
                            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                            float sum = 0.f;
#pragma unroll
                            for (int ni = 0; ni < NTW; ni++)
#pragma unroll
                                for (int e = 0; e < 2; e++) {
                                    sc[mi][ni][2 * r + e] = expf(sc[mi][ni][2 * r + e] - mx);
                                    sum += sc[mi][ni][2 * r + e];
                                }
                            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                            const float inv = 1.0f / sum;
#pragma unroll
                            for (int ni = 0; ni < NTW; ni++)
#pragma unroll
                                for (int e = 0; e < 2; e++) sc[mi][ni][2 * r + e] *= inv;
                        }
                    if (FULL) pair_sync();  // all 32 rows of V staged (two 16-row neighbourhoods read only their own rows)
                    else __syncwarp();
                    // ---- ctx = P V + bias_v : contraction over the keys; 32 output channels at a time
#pragma unroll
                    for (int nb = 0; nb < HD / 32; nb++) {   // (unrolled: the two channel halves are independent chains)
                        float o[MTW][4][4];
#pragma unroll
                        for (int mi = 0; mi < MTW; mi++)
#pragma unroll
                            for (int n4 = 0; n4 < 4; n4++) o[mi][n4][0] = o[mi][n4][1] = o[mi][n4][2] = o[mi][n4][3] = 0.f;
                        auto bfrag = [&](int ks, u32 (&bh)[4][2], u32 (&bl)[4][2]) {
#pragma unroll
                            for (int n4 = 0; n4 < 4; n4++) {
                                split(sV[(8 * ks + t2) * kAttnVP + 32 * nb + 8 * n4 + g], bh[n4][0], bl[n4][0]);
                                split(sV[(8 * ks + t2 + 1) * kAttnVP + 32 * nb + 8 * n4 + g], bh[n4][1], bl[n4][1]);
                            }
                        };
                        auto afrag = [&](const float (&pf)[4], u32 (&ah)[4], u32 (&al)[4]) {
                            split(pf[0], ah[0], al[0]);
                            split(pf[2], ah[1], al[1]);
                            split(pf[1], ah[2], al[2]);
                            split(pf[3], ah[3], al[3]);
                        };
                        if constexpr (BF) {
                            // keys 16 s .. 16 s + 15 per step: A = two score n-tiles packed, B = four V rows packed in pairs
                            auto bfrag16 = [&](int s16, u32 (&bh)[4][2], u32 (&bl)[4][2]) {
#pragma unroll
                                for (int n4 = 0; n4 < 4; n4++) {
                                    const float *vp = sV + (16 * s16 + t2) * kAttnVP + 32 * nb + 8 * n4 + g;
                                    split2(vp[0], vp[kAttnVP], bh[n4][0], bl[n4][0]);
                                    split2(vp[8 * kAttnVP], vp[9 * kAttnVP], bh[n4][1], bl[n4][1]);
                                }
                            };
                            auto afrag16 = [&](const float (&p0)[4], const float (&p1)[4], u32 (&ah)[4], u32 (&al)[4]) {
                                split2(p0[0], p0[1], ah[0], al[0]);
                                split2(p0[2], p0[3], ah[1], al[1]);
                                split2(p1[0], p1[1], ah[2], al[2]);
                                split2(p1[2], p1[3], ah[3], al[3]);
                            };
                            if constexpr (FULL) {
#pragma unroll
                                for (int s16 = 0; s16 < 2; s16++) {
                                    u32 bh[4][2], bl[4][2], ah[MTW][4], al[MTW][4];
                                    bfrag16(s16, bh, bl);
#pragma unroll
                                    for (int mi = 0; mi < MTW; mi++) afrag16(sc[mi][2 * s16], sc[mi][2 * s16 + 1], ah[mi], al[mi]);
#pragma unroll
                                    for (int pass = PASS0; pass < 3; pass++)
#pragma unroll
                                        for (int n4 = 0; n4 < 4; n4++)
#pragma unroll
                                            for (int mi = 0; mi < MTW; mi++)
                                                mma16(o[mi][n4], pass == 1 ? al[mi] : ah[mi], pass == 0 ? bl[n4] : bh[n4]);
                                }
                            } else {
#pragma unroll
                                for (int mi = 0; mi < MTW; mi++) {
                                    u32 bh[4][2], bl[4][2], ah[4], al[4];
                                    bfrag16(mt0 + mi, bh, bl);
                                    afrag16(sc[mi][0], sc[mi][1], ah, al);
#pragma unroll
                                    for (int pass = PASS0; pass < 3; pass++)
#pragma unroll
                                        for (int n4 = 0; n4 < 4; n4++)
                                            mma16(o[mi][n4], pass == 1 ? al : ah, pass == 0 ? bl[n4] : bh[n4]);
                                }
                            }
                        } else {
#pragma unroll
                        for (int ksl = 0; ksl < NTW; ksl++) {
                            if constexpr (FULL) {
                                u32 bh[4][2], bl[4][2], ah[MTW][4], al[MTW][4];
                                bfrag(ksl, bh, bl);
#pragma unroll
                                for (int mi = 0; mi < MTW; mi++) afrag(sc[mi][ksl], ah[mi], al[mi]);
#pragma unroll
                                for (int pass = 0; pass < 3; pass++)
#pragma unroll
                                    for (int n4 = 0; n4 < 4; n4++)
#pragma unroll
                                        for (int mi = 0; mi < MTW; mi++)
                                            mma(o[mi][n4], pass == 1 ? al[mi] : ah[mi], pass == 0 ? bl[n4] : bh[n4]);
                            } else {
#pragma unroll
                                for (int mi = 0; mi < MTW; mi++) {
                                    u32 bh[4][2], bl[4][2], ah[4], al[4];
                                    bfrag(2 * (mt0 + mi) + ksl, bh, bl);
                                    afrag(sc[mi][ksl], ah, al);
#pragma unroll
                                    for (int pass = 0; pass < 3; pass++)
#pragma unroll
                                        for (int n4 = 0; n4 < 4; n4++)
                                            mma(o[mi][n4], pass == 1 ? al : ah, pass == 0 ? bl[n4] : bh[n4]);
                                }
                            }
                        }
                        }
#pragma unroll
                        for (int mi = 0; mi < MTW; mi++)
#pragma unroll
                            for (int r = 0; r < 2; r++) {
                                const long long row = wrow0 + 16 * (mt0 + mi) + 8 * r + g;
                                if (row < p.T) {
#pragma unroll
                                    for (int n4 = 0; n4 < 4; n4++) {
                                        const float2 bv = *reinterpret_cast<const float2 *>(sbias + 2 * HD + 32 * nb + 8 * n4 + t2);
                                        const float oa = o[mi][n4][2 * r] + bv.x, ob = o[mi][n4][2 * r + 1] + bv.y;
                                        const size_t oo = (size_t)n_group * HD + (size_t)row * p.ldo + 32 * nb + 8 * n4 + t2;
                                        if (out16) *reinterpret_cast<u32 *>(reinterpret_cast<__half *>(p.out) + oo) = f16x2(oa, ob);
                                        else *reinterpret_cast<float2 *>(p.out + oo) = make_float2(oa, ob);
                                    }
                                }
                            }
                    }
                };
                if (ns == 32) attend(std::true_type{});
                else attend(std::false_type{});
                __syncwarp();  // sbias and (EW = 4) the staging tile are rewritten for the next item
            }

            if (EPI != E_ATTN && !CHUNKED) {  // (the attention and chunked LayerNorm epilogues release earlier)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 2 && !leader) mbar_arrive_remote(acc_empty(as), 0);
                    else mbar_arrive(acc_empty(as));
                }
            }
            if (ACC_STAGES == 2) {
                as ^= 1;
                if (as == 0) aphase ^= 1;
            } else {
                aphase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // the leader's MMAs read the peer's shared memory: nobody leaves before both are done
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, kTmemCols);
    }
}

// Launch policy of the calling thread (pdab_set_persistent_ctas / pdab_set_cta_pairs): thread-local, so two pipelines
// driven from two host threads do not see each other's settings, and baked into a CUDA graph at capture time.
// persistent CTAs: 0 = every SM of the current device; a pipelined caller that overlaps these kernels with SM-filling
// latency chains of another batch (FPS: one 192 KB-smem CTA per scene) lowers it so that no CTA of the grid waits for an SM.
thread_local int g_persistent_ctas = 0;
// 1: cta_group::2 CTA pairs (M = 256 per pair) whenever the problem has at least one full pair tile per pair; 0: single CTAs.
thread_local int g_cta_pairs = 1;

template <int NPASS, int BN, int NCH, int ALOAD, int EPI, int CG>
int launch_cg(GemmParams p, cudaStream_t s) {
    using C = Cfg<NPASS, BN, CG, EpiWarps<EPI, CG, ALOAD>::value, ALOAD>;
    auto kern = tc_gemm_kernel<NPASS, BN, NCH, ALOAD, EPI, CG>;
    // per device and cheap: set on every launch (a cached flag would skip the second GPU of a process)
    PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    p.n_items = ((p.T + BM * CG - 1) / (BM * CG)) * p.n_groups;
    long long grid = p.n_items * CG < p.persistent_ctas ? p.n_items * CG : p.persistent_ctas;
    if (CG == 2) grid &= ~1LL;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PDAB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    PDAB_LAUNCH_CHECK();
    return 0;
}

template <int NPASS, int BN, int NCH, int ALOAD, int EPI>
int launch(GemmParams &p, cudaStream_t s) {
    p.persistent_ctas = pdab::persistent_ctas();
    // CTA pairs pay off once every pair has whole 256-row tiles to chew on; tiny problems stay on single CTAs
    if (g_cta_pairs && p.persistent_ctas >= 2 && p.T >= 2 * BM * 8)
        return launch_cg<NPASS, BN, NCH, ALOAD, EPI, 2>(p, s);
    return launch_cg<NPASS, BN, NCH, ALOAD, EPI, 1>(p, s);
}

template <int NPASS, int ALOAD>
int dispatch(GemmParams &p, int epi, int bn, cudaStream_t s) {
    const int chunks = (p.Nout + bn - 1) / bn;
    auto items = [&](int nch) { p.n_groups = chunks / nch; };
    if (epi == E_ADD_LN) {
        if (ALOAD == A_GATHER || bn != 256 || p.Nout % 256 || chunks > 2) return PDAB_EUNSUPPORTED;
        constexpr int AL = ALOAD == A_GATHER ? A_ROWS : ALOAD;
        if (chunks == 1) {
            items(1);
            return launch<NPASS, 256, 1, AL, E_ADD_LN>(p, s);
        }
        items(2);
        return launch<NPASS, 256, 2, AL, E_ADD_LN>(p, s);
    }
    items(1);
    if (epi == E_ATTN) {
        if (ALOAD == A_GATHER || bn != 192 || p.Nout % 192) return PDAB_EUNSUPPORTED;
        constexpr int AL = ALOAD == A_GATHER ? A_ROWS : ALOAD;
        return launch<NPASS, 192, 1, AL, E_ATTN>(p, s);
    }
    if constexpr (ALOAD != A_GATHER)
        if (bn == 192 && epi == E_STORE) return launch<NPASS, 192, 1, ALOAD, E_STORE>(p, s);
    if (bn == 256) {
        switch (epi) {
            case E_STORE: return launch<NPASS, 256, 1, ALOAD, E_STORE>(p, s);
            case E_RELU: return launch<NPASS, 256, 1, ALOAD, E_RELU>(p, s);
            case E_ADD_MAXPOOL: return launch<NPASS, 256, 1, ALOAD, E_ADD_MAXPOOL>(p, s);
            case E_RELU_MAXPOOL: return launch<NPASS, 256, 1, ALOAD, E_RELU_MAXPOOL>(p, s);
        }
    } else if (bn == 128) {
        switch (epi) {
            case E_STORE: return launch<NPASS, 128, 1, ALOAD, E_STORE>(p, s);
            case E_RELU: return launch<NPASS, 128, 1, ALOAD, E_RELU>(p, s);
            case E_ADD_MAXPOOL: return launch<NPASS, 128, 1, ALOAD, E_ADD_MAXPOOL>(p, s);
            case E_RELU_MAXPOOL: return launch<NPASS, 128, 1, ALOAD, E_RELU_MAXPOOL>(p, s);
        }
    }
    return PDAB_EUNSUPPORTED;
}

}  // namespace

int pdab::persistent_ctas() {
    if (g_persistent_ctas > 0) return g_persistent_ctas;
    int dev = 0, sms = pdab::kNumSMs;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

extern "C" int pdab_set_persistent_ctas(int n) {
    if (n < 0 || n > 1024) return PDAB_EINVAL;
    g_persistent_ctas = n;
    return 0;
}

extern "C" int pdab_set_cta_pairs(int on) {
    g_cta_pairs = on ? 1 : 0;
    return 0;
}

// See include/pdab.h for the contract.
extern "C" int pdab_tc_linear(long long rows, int k, int nout, int npass, int bn, int epilogue, const float *a, int lda,
                              const float *w_packed, const float *bias, const float *residual, int ldr,
                              const float *gamma, const float *beta, float eps, int nsample, float *out, int ldo,
                              pdab_stream_t stream) {
    if (rows < 0 || k < 1 || nout < 1 || !a || !w_packed || !out) return PDAB_EINVAL;
    if (rows == 0) return 0;
    if ((k & 3) || (lda & 3) || (ldo & 3) || (nout & 3) || lda < k) return PDAB_EINVAL;
    if (npass < 1 || npass > 3) return PDAB_EINVAL;
    if (npass == 2 && (k & 7)) return PDAB_EINVAL;
    if (bn != 128 && bn != 256 && !(bn == 192 && (epilogue == E_ATTN || epilogue == E_STORE))) return PDAB_EINVAL;
    if (epilogue == E_ATTN && (bn != 192 || nout % 192 || (nsample != 16 && nsample != 32) || rows % nsample))
        return PDAB_EUNSUPPORTED;
    if ((epilogue == E_ADD_LN || epilogue == E_ADD_MAXPOOL) && (!residual || (ldr & 3))) return PDAB_EINVAL;
    if (epilogue == E_ADD_LN && (!gamma || !beta)) return PDAB_EINVAL;
    if ((epilogue == E_ADD_MAXPOOL || epilogue == E_RELU_MAXPOOL) && ((nsample != 16 && nsample != 32) || rows % nsample))
        return PDAB_EUNSUPPORTED;
    GemmParams p{};
    p.A = a;
    p.lda = lda;
    p.T = rows;
    p.K = k;
    p.KA = (k + bk_of(npass) - 1) / bk_of(npass);
    p.Wp = w_packed;
    p.bias = bias;
    p.Nout = nout;
    p.out = out;
    p.ldo = ldo;
    p.R = residual;
    p.ldr = ldr;
    p.gamma = gamma;
    p.beta = beta;
    p.eps = eps;
    p.ns = nsample;
    cudaStream_t s = pdab::to_stream(stream);
    return npass == 3 ? dispatch<3, A_ROWS>(p, epilogue, bn, s)
           : npass == 2 ? dispatch<2, A_ROWS>(p, epilogue, bn, s) : dispatch<1, A_ROWS>(p, epilogue, bn, s);
}

// fp16 single-pass form (NPASS = 4); see include/pdab.h.
extern "C" int pdab_tc_linear_h(long long rows, int k, int nout, int bn, int epilogue, const void *a, int lda, int a_fp16,
                                const float *w_packed, const float *bias, const void *res_hi, const void *res_lo, int ldr,
                                const float *gamma, const float *beta, float eps, int nsample, void *out, void *out_lo,
                                int ldo, int out_fmt, pdab_stream_t stream) {
    if (rows < 0 || k < 1 || nout < 1 || !a || !w_packed || !out) return PDAB_EINVAL;
    if (rows == 0) return 0;
    if ((k & 7) || (lda & 7) || (ldo & 3) || (nout & 3) || lda < k || rows > 0x7fffffffLL) return PDAB_EINVAL;
    if (out_fmt < 0 || out_fmt > 2 || (out_fmt == 2 && !out_lo)) return PDAB_EINVAL;
    if (bn != 128 && bn != 256 && !(bn == 192 && (epilogue == E_ATTN || epilogue == E_STORE))) return PDAB_EINVAL;
    if (epilogue == E_ATTN && (bn != 192 || nout % 192 || (nsample != 16 && nsample != 32) || rows % nsample || out_fmt == 2))
        return PDAB_EUNSUPPORTED;
    if ((epilogue == E_ADD_LN || epilogue == E_ADD_MAXPOOL) && (!res_hi || (ldr & 3))) return PDAB_EINVAL;
    if (epilogue == E_ADD_LN && (!gamma || !beta)) return PDAB_EINVAL;
    if (epilogue != E_ADD_LN && out_fmt == 2) return PDAB_EUNSUPPORTED;
    if ((epilogue == E_ADD_MAXPOOL || epilogue == E_RELU_MAXPOOL) &&
        ((nsample != 16 && nsample != 32 && !(nsample == 64 && epilogue == E_RELU_MAXPOOL)) || rows % nsample || out_fmt != 0))
        return PDAB_EUNSUPPORTED;
    GemmParams p{};
    p.A = reinterpret_cast<const float *>(a);
    p.lda = lda;
    p.T = rows;
    p.K = k;
    p.KA = (k + 63) / 64;
    p.Wp = w_packed;
    p.bias = bias;
    p.Nout = nout;
    p.out = reinterpret_cast<float *>(out);
    p.out_lo = out_lo;
    p.out16 = out_fmt;
    p.ldo = ldo;
    p.R = reinterpret_cast<const float *>(res_hi);
    p.R_lo = res_lo;
    p.res16 = res_lo != nullptr;
    p.ldr = ldr;
    p.gamma = gamma;
    p.beta = beta;
    p.eps = eps;
    p.ns = nsample;
    cudaStream_t s = pdab::to_stream(stream);
    if (nsample == 64 && epilogue == E_RELU_MAXPOOL)   // the two half-neighbourhood maxima meet through atomicMax
        PDAB_CUDA(cudaMemset2DAsync(out, (size_t)ldo * sizeof(float), 0, (size_t)nout * sizeof(float), (size_t)(rows / 64), s));
    if (a_fp16) {
        if (reinterpret_cast<uintptr_t>(a) & 15) return PDAB_EINVAL;
        int rc = make_a_map(&p.tmA, a, rows, k, lda);
        if (rc) return rc;
        // residual as a (hi, lo) fp16 plane pair covering whole accumulator chunks: staged through shared memory by TMA
        if (res_lo && (epilogue == E_ADD_LN || epilogue == E_ADD_MAXPOOL) && nout % bn == 0 && !(ldr & 7) &&
            !((reinterpret_cast<uintptr_t>(res_hi) | reinterpret_cast<uintptr_t>(res_lo)) & 15)) {
            rc = make_a_map(&p.tmRh, res_hi, rows, nout, ldr);
            if (!rc) rc = make_a_map(&p.tmRl, res_lo, rows, nout, ldr);
            if (rc) return rc;
            p.res_tma = 1;
        }
        return dispatch<4, A_TMA>(p, epilogue, bn, s);
    }
    return dispatch<4, A_ROWS>(p, epilogue, bn, s);
}

static int sa_gather(int b, int c, int n, int m, int nsample, int nout, int npass, const int *idx,
                     const float *features_t, const float *xyz, const float *new_xyz, const float *w_packed,
                     const float *bias, void *out, int ldo, int out_fp16, pdab_stream_t stream) {
    if (b < 0 || c < 0 || n < 1 || m < 1 || nsample < 1 || nout < 1 || !idx || !xyz || !new_xyz || !w_packed || !out)
        return PDAB_EINVAL;
    if (b == 0) return 0;
    if ((c & 3) || (c > 0 && !features_t) || (ldo & 3) || (nout & 3)) return PDAB_EINVAL;
    if (npass < 1 || npass > 4 || (BK_IS_64(npass) && (c & 7))) return PDAB_EINVAL;
    if (out_fp16 && npass != 4) return PDAB_EINVAL;
    GemmParams p{};
    p.T = (long long)b * m * nsample;
    p.K = c + 3;
    p.KA = (p.K + bk_of(npass) - 1) / bk_of(npass);
    p.idx = idx;
    p.feat_t = features_t;
    p.xyz = xyz;
    p.new_xyz = new_xyz;
    p.C = c;
    p.Nsrc = n;
    p.M = m;
    p.ns = nsample;
    p.Wp = w_packed;
    p.bias = bias;
    p.Nout = nout;
    p.out = reinterpret_cast<float *>(out);
    p.out16 = out_fp16 ? 1 : 0;
    p.ldo = ldo;
    cudaStream_t s = pdab::to_stream(stream);
    return npass == 4   ? dispatch<4, A_GATHER>(p, E_RELU, 256, s)
           : npass == 3 ? dispatch<3, A_GATHER>(p, E_RELU, 256, s)
           : npass == 2 ? dispatch<2, A_GATHER>(p, E_RELU, 256, s) : dispatch<1, A_GATHER>(p, E_RELU, 256, s);
}

extern "C" int pdab_tc_sa_gather_linear(int b, int c, int n, int m, int nsample, int nout, int npass, const int *idx,
                                        const float *features_t, const float *xyz, const float *new_xyz,
                                        const float *w_packed, const float *bias, float *out, int ldo,
                                        pdab_stream_t stream) {
    if (npass == 4) return PDAB_EINVAL;   // the fp16 mode is pdab_tc_sa_gather_linear_h
    return sa_gather(b, c, n, m, nsample, nout, npass, idx, features_t, xyz, new_xyz, w_packed, bias, out, ldo, 0, stream);
}

extern "C" int pdab_tc_sa_gather_linear_h(int b, int c, int n, int m, int nsample, int nout, const int *idx,
                                          const float *features_t, const float *xyz, const float *new_xyz,
                                          const float *w_packed, const float *bias, void *out, int ldo, int out_fp16,
                                          pdab_stream_t stream) {
    return sa_gather(b, c, n, m, nsample, nout, 4, idx, features_t, xyz, new_xyz, w_packed, bias, out, ldo, out_fp16, stream);
}
