// Shared device / host helpers of the tcgen05 kernels (tc_gemm.cu, tc_ffn.cu): mbarrier, TMA, tcgen05 / TMEM PTX wrappers,
// UMMA descriptors, fp16 packing, and the host-side tensor-map encoder.  sm_100a only.  Everything sits in an anonymous
// namespace: each translation unit gets its own copy.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

using u32 = uint32_t;
using u64 = uint64_t;

constexpr int BM = 128;                    // rows per tile (TMEM lanes)
constexpr int A_TILE_BYTES = BM * 128;     // 16 KB: 128 rows x one 128-byte swizzle row
constexpr int kTmemCols = 512;

// ------------------------------------------------------------------------------------------- PTX wrappers

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u32 bar, u32 parity) {
    u32 ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s at 1.9 GHz; a healthy wait is microseconds
    }
}
// Barriers that the peer CTA of a pair arrives on remotely are waited on with the ordinary (CTA-scope acquire) form and
// signalled with the default-semantics remote arrive below — the forms CUTLASS's ClusterBarrier uses for the same
// producer / consumer hand-offs.  (The explicit .release.cluster / .acquire.cluster forms compile to MEMBAR.ALL.GPU and
// CCTL.IVALL on every stage: measured in profiles/r01_ncu_tc_gemm_*.)  What is handed over is either shared memory the
// peer wrote behind a fence.proxy.async and that the PEER's own tensor-core datapath reads, or TMEM the peer finished
// reading behind tcgen05.wait::ld + tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mbar_wait_cluster(u32 bar, u32 parity) { mbar_wait(bar, parity); }
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(u32 bar, u32 cta) {
    asm volatile(
        "{\n.reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n}" ::"r"(bar),
        "r"(cta)
        : "memory");
}
__device__ __forceinline__ u32 cluster_ctarank() {
    u32 r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// TMA tensor load (UTMALDG): box (c0 = first column, c1 = first row) of a 2-D tensor map -> shared memory of this CTA,
// completing `bytes` on this CTA's mbarrier.  The map lives in the kernel's parameter space (__grid_constant__).
__device__ __forceinline__ void tma_load_2d(u32 dst, const CUtensorMap *map, int c0, int c1, u32 bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// the same box pulled into L2 only (no shared memory, no barrier): hides DRAM latency behind a short smem ring
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// CG = 1: one SM; CG = 2: CTA pair (both CTAs' allocating warps execute the instruction, same smem slot offset)
template <int CG>
__device__ __forceinline__ void tmem_alloc(u32 dst_smem, u32 ncols) {
    if (CG == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(u32 taddr, u32 ncols) {
    if (CG == 2)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32, issued by ONE thread.
template <int CG>
__device__ __forceinline__ void umma_tf32(u32 d_tmem, u64 adesc, u64 bdesc, u32 idesc, u32 accumulate) {
    if (CG == 2)
        asm volatile(
            "{\n.reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n.reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// Same with bf16 operands (kind::f16): K = 16 per instruction, twice the tf32 rate.
template <int CG>
__device__ __forceinline__ void umma_bf16(u32 d_tmem, u64 adesc, u64 bdesc, u32 idesc, u32 accumulate) {
    if (CG == 2)
        asm volatile(
            "{\n.reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n.reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// Arrive on an mbarrier when all MMAs issued so far by this thread have completed (implies fence::before_thread_sync).
// CG = 2: the arrival is multicast to the barrier at this offset in BOTH CTAs of the pair.
template <int CG>
__device__ __forceinline__ void umma_commit(u32 bar) {
    if (CG == 2)
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
            "h"((unsigned short)3)
            : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread `lane` of the warp gets columns [c, c+32) of TMEM lane (warp%4)*32+lane.
__device__ __forceinline__ void tmem_ld32(u32 taddr, float (&v)[32]) {
    u32 r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(u32 taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};\n"
        "tcgen05.wait::st.sync.aligned;" ::"r"(__float_as_uint(v[0])),
        "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
        "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
        "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])),
        "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])),
        "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])),
        "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31])), "r"(taddr)
        : "memory");
}

// 16 lanes x 32 columns in the MMA-C-fragment layout: register k*4 + j*2 + e of lane t holds TMEM lane
// base + 8 j + t / 4, column base + 8 k + 2 (t % 4) + e   (cute SM100_TMEM_LOAD_16dp256b4x).
__device__ __forceinline__ void tmem_ld_16x256b_x4(u32 taddr, float (&v)[16]) {
    u32 r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st_16x256b_x4(u32 taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%16], "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};" ::"r"(__float_as_uint(v[0])),
        "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
        "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
        "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, canonical K-major SWIZZLE_128B tile (rows of 128 B, 8-row groups 1024 B apart):
// start address >> 4 | LBO (unused for swizzled K-major, 1) | SBO = 1024 >> 4 | version 1 (sm_100) | layout 2 (SW128).
__device__ __forceinline__ u64 umma_desc(u32 smem_addr) {
    return (u64)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor: D fp32, A/B tf32, both K-major, N = BN, M = 128.
// FMT: 2 = TF32 (kind::tf32), 1 = BF16 (kind::f16).
template <int BN, int FMT, int CG>
__device__ __forceinline__ constexpr u32 umma_idesc() {
    return (1u << 4) | ((u32)FMT << 7) | ((u32)FMT << 10) | ((u32)(BN >> 3) << 17) | ((u32)((BM * CG) >> 4) << 24);
}
// 8 fp32 -> 8 bf16 (round to nearest even) packed in a uint4, and the bf16 of the remainders
__device__ __forceinline__ void bf16_split8(const float4 &a, const float4 &b, uint4 &hi, uint4 &lo) {
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    u32 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        // cvt.rn.bf16x2.f32 d, hi_half_src, lo_half_src
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(x[2 * i + 1]), "f"(x[2 * i]));
        const float r0 = x[2 * i] - __uint_as_float(h[i] << 16);
        const float r1 = x[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l[i]) : "f"(r1), "f"(r0));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// 8 fp32 -> 8 fp16 (round to nearest even) packed in a uint4
__device__ __forceinline__ u32 f16x2(float lo_half, float hi_half) {
    u32 r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_half), "f"(lo_half));
    return r;
}
__device__ __forceinline__ uint4 f16_pack8(const float4 &a, const float4 &b) {
    return make_uint4(f16x2(a.x, a.y), f16x2(a.z, a.w), f16x2(b.x, b.y), f16x2(b.z, b.w));
}
__device__ __forceinline__ float2 f16x2_to_float2(u32 h) {
    return __half22float2(*reinterpret_cast<const __half2 *>(&h));
}
// (a, b) -> fp16 pair `hi` and the fp16 pair of the remainders `lo`: a = hi.x + lo.x to ~2^-22 relative
__device__ __forceinline__ void f16_split2(float a, float b, u32 &hi, u32 &lo) {
    hi = f16x2(a, b);
    const float2 h = f16x2_to_float2(hi);
    lo = f16x2(a - h.x, b - h.y);
}

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ float tf32_rna(float v) {
    u32 r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

struct Pipe {
    int stage = 0;
    u32 phase = 0;
    template <int S>
    __device__ __forceinline__ void advance() {
        if (++stage == S) {
            stage = 0;
            phase ^= 1;
        }
    }
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// fp16 (rows, ld) row-major matrix -> boxes of 64 columns x 128 rows, SWIZZLE_128B (the K-major UMMA tile), zero fill
int make_a_map(CUtensorMap *map, const void *a, long long rows, int k, int ld) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return PDAB_EUNSUPPORTED;
    const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(a), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : PDAB_EINVAL;
}

}  // namespace
