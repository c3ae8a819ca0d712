// Farthest-point sampling for sm_100a.
//
// One CTA of 1024 threads per scene keeps the whole point set in shared memory
// (SoA, up to 16384 points = 192 KB) and the running min-distances in
// registers (<= 16 per thread); larger clouds use a thread-block cluster of
// 2..16 CTAs per scene that exchange one 32-byte candidate per iteration
// through distributed shared memory.  The per-iteration argmax is two REDUX
// warp reductions + one block barrier (+ one cluster barrier when clustered)
// instead of the reference's 10-level shared-memory tree
// (PB/src/sampling_gpu.cu:143-203).
//
// Exactness contract (include/pdab.h, pdab_fps): distances in the reference's
// compiled fp32 op order, and the reference's tie rule — among points tied at
// the maximum the winner is argmin (bitrev_L(k mod BS), k).  Both are folded
// into ONE 64-bit key per candidate: [ fp32 bits of dist | ~tiekey(k) ], so a
// plain unsigned max picks the reference's winner.
#include <cooperative_groups.h>

#include <cmath>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pdab {
int fps_pruned(int b, int n, int m, const float *xyz, float *temp, int *idx, int L, int CL, cudaStream_t stream);
}

namespace {

constexpr int kThreads = 1024;
constexpr int kMaxCluster = 16;
// Largest cluster (CTAs per scene) the N > 16384 path may grow to.  Default 16: as many SMs per scene as fit, shortest chain.
// A caller that pipelines batches lowers it (pdab_set_fps_max_cluster) so that the FPS chain of one batch occupies few SMs
// and runs beside the other batches' kernels instead of taking the whole GPU.
thread_local int g_max_cluster = kMaxCluster;   // launch policy of the calling thread (pdab_set_fps_max_cluster)

struct __align__(16) Candidate {
    unsigned long long key;  // [dist bits | ~tiekey]; 0 = no candidate
    float x, y, z;
    float pad;
};

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | ml;
}

// tiekey(k): top L bits = bit-reversed (k mod 2^L), low 32-L bits = k >> L.
__device__ __forceinline__ unsigned tie_key(int k, int L) {
    if (L == 0) return (unsigned)k;
    return __brev((unsigned)k & ((1u << L) - 1u)) | ((unsigned)k >> L);
}
__device__ __forceinline__ int tie_key_decode(unsigned key, int L) {
    if (L == 0) return (int)key;
    const unsigned lowmask = (1u << (32 - L)) - 1u;
    return (int)(((key & lowmask) << L) | __brev(key & ~lowmask));
}

// P: points per thread.  MATRIX: distances come from row `old` of a (N,N)
// matrix (F-FPS, PB/src/sampling_gpu.cu:294) instead of coordinates.
template <int P, bool MATRIX>
__global__ void __launch_bounds__(kThreads, 1)
fps_kernel(int n, int m, const float *__restrict__ src, float *__restrict__ temp, int *__restrict__ idxs, int L,
           int CL) {
    extern __shared__ float smem[];
    __shared__ unsigned long long red[2][32];
    __shared__ Candidate xchg[2][kMaxCluster];
    __shared__ __align__(8) unsigned long long xbar[2];  // mbarriers: candidates of all CL CTAs have landed in xchg[buf]

    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int rank = CL > 1 ? (int)cg::this_cluster().block_rank() : 0;
    const int scene = blockIdx.x / CL;
    constexpr int cap = P * kThreads;
    float *sx = smem, *sy = smem + cap, *sz = smem + 2 * cap;

    const float *xyz = MATRIX ? nullptr : src + (size_t)scene * n * 3;
    const float *mat = MATRIX ? src + (size_t)scene * n * n : nullptr;
    temp += (size_t)scene * n;
    idxs += (size_t)scene * m;

    // thread t of CTA `rank` owns points k = t + 1024 * (rank + CL * j), j < P
    float dist[P];
#pragma unroll
    for (int j = 0; j < P; j++) {
        const int k = t + kThreads * (rank + CL * j);
        if (k < n) {
            dist[j] = temp[k];
            if (!MATRIX) {
                sx[j * kThreads + t] = xyz[k * 3 + 0];
                sy[j * kThreads + t] = xyz[k * 3 + 1];
                sz[j * kThreads + t] = xyz[k * 3 + 2];
            }
        } else {
            dist[j] = -1.0f;  // never a candidate: real distances are >= 0
            if (!MATRIX) {
                sx[j * kThreads + t] = 0.f;
                sy[j * kThreads + t] = 0.f;
                sz[j * kThreads + t] = 0.f;
            }
        }
    }
    if (CL > 1) {
        if (t == 0) {
            for (int i = 0; i < 2; i++)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&xbar[i])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cg::this_cluster().sync();  // peers resident and their barriers initialised before any DSMEM store
    } else {
        __syncthreads();
    }

    int old = 0;
    float x1 = 0.f, y1 = 0.f, z1 = 0.f;
    if (!MATRIX) {
        x1 = xyz[0];
        y1 = xyz[1];
        z1 = xyz[2];
    }
    if (t == 0 && rank == 0) idxs[0] = 0;

    int buf = 0;
    for (int it = 1; it < m; it++) {
        float bestv = -1.0f;
        int bestj = -1;
#pragma unroll
        for (int j = 0; j < P; j++) {
            float d;
            if (MATRIX) {
                const int k = t + kThreads * j;
                d = k < n ? __ldg(mat + (size_t)old * n + k) : 0.f;
            } else {
                d = pdab::sqdist3(sx[j * kThreads + t], sy[j * kThreads + t], sz[j * kThreads + t], x1, y1, z1);
            }
            const float dd = fminf(d, dist[j]);
            dist[j] = dd;
            if (dd > bestv) {  // strict: lowest j (= lowest k of this thread) wins ties
                bestv = dd;
                bestj = j;
            }
        }
        unsigned long long cand = 0ull;
        if (bestj >= 0) {
            const int k = t + kThreads * (rank + CL * bestj);
            cand = ((unsigned long long)__float_as_uint(bestv) << 32) | (unsigned long long)(~tie_key(k, L));
        }
        cand = warp_max_u64(cand);
        if (lane == 0) red[buf][warp] = cand;
        __syncthreads();
        cand = warp_max_u64(red[buf][lane]);  // every warp redoes the final 32 -> 1: no second barrier

        if (CL == 1) {
            old = tie_key_decode(~(unsigned)cand, L);
            if (!MATRIX) {
                const int slot = (old >> 10) * kThreads + (old & (kThreads - 1));
                x1 = sx[slot];
                y1 = sy[slot];
                z1 = sz[slot];
            }
        } else {
            // Exchange without a cluster barrier (barrier.cluster release/acquire costs ~2 us per step): every CTA
            // pushes its 32-byte candidate into each peer's xchg[buf][rank] with st.async, which completes on the
            // RECEIVER's mbarrier; a CTA waits for CL x 32 bytes on its own barrier.  Buffer reuse is safe: a peer can
            // only send step it+2 after it has consumed step it+1, which needs this CTA's step it+1 candidate, which is
            // sent after this CTA's block barrier of step it+1, i.e. after all its threads have read step it's buffer.
            const unsigned bar_local = (unsigned)__cvta_generic_to_shared(&xbar[buf]);
            if (t == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_local), "r"(CL * 32)
                             : "memory");
            if (t < CL) {
                float cx = 0.f, cy = 0.f, cz = 0.f;
                if (cand != 0ull) {
                    const int k = tie_key_decode(~(unsigned)cand, L);
                    const int slot = ((k >> 10) / CL) * kThreads + (k & (kThreads - 1));
                    cx = sx[slot];
                    cy = sy[slot];
                    cz = sz[slot];
                }
                const unsigned dst_local = (unsigned)__cvta_generic_to_shared(&xchg[buf][rank]);
                unsigned dst, rbar;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(dst_local), "r"(t));
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(bar_local), "r"(t));
                asm volatile(
                    "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
                    "r"((unsigned)cand), "r"((unsigned)(cand >> 32)), "r"(__float_as_uint(cx)), "r"(__float_as_uint(cy)),
                    "r"(rbar)
                    : "memory");
                asm volatile(
                    "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst + 16),
                    "r"(__float_as_uint(cz)), "r"(0u), "r"(0u), "r"(0u), "r"(rbar)
                    : "memory");
            }
            {
                const unsigned parity = (unsigned)(((it - 1) >> 1) & 1);   // buffer `buf` is used every second step
                unsigned ok = 0;
                const long long t0 = clock64();
                while (!ok) {
                    asm volatile(
                        "{\n.reg .pred p;\n"
                        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                        "selp.u32 %0, 1, 0, p;\n}"
                        : "=r"(ok)
                        : "r"(bar_local), "r"(parity)
                        : "memory");
                    if (!ok && clock64() - t0 > 4000000000LL) __trap();  // never hang the GPU on a protocol bug
                }
            }
            // every warp picks the winner itself: lane i holds CTA i's key, REDUX max, lowest lane with the max
            const unsigned long long mine = lane < CL ? xchg[buf][lane].key : 0ull;
            const unsigned long long w = warp_max_u64(mine);
            const int wi = __ffs(__ballot_sync(0xffffffffu, mine == w && lane < CL)) - 1;
            old = tie_key_decode(~(unsigned)w, L);
            x1 = xchg[buf][wi].x;
            y1 = xchg[buf][wi].y;
            z1 = xchg[buf][wi].z;
        }
        if (t == 0 && rank == 0) idxs[it] = old;
        buf ^= 1;
    }

#pragma unroll
    for (int j = 0; j < P; j++) {
        const int k = t + kThreads * (rank + CL * j);
        if (k < n) temp[k] = dist[j];
    }
    if (CL > 1) cg::this_cluster().sync();  // nobody exits while a peer's last st.async may still target its shared memory
}

template <int P, bool MATRIX>
int launch(int b, int n, int m, const float *src, float *temp, int *idx, int L, int CL, cudaStream_t stream) {
    const size_t smem = MATRIX ? 0 : (size_t)3 * P * kThreads * sizeof(float);
    auto kern = fps_kernel<P, MATRIX>;
    // per device and cheap: set on every launch (a cached flag would skip the second GPU of a process)
    PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(b * CL);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PDAB_CUDA(cudaLaunchKernelEx(&cfg, kern, n, m, src, temp, idx, L, CL));
    return 0;
}

// log2 of the reference's block size, PB/src/cuda_utils.h:10-14 (same double
// arithmetic, including its truncation).
int ref_log2_block(int n) {
    const int pow_2 = (int)(std::log((double)n) / std::log(2.0));
    int L = pow_2;
    if (L > 10) L = 10;
    if (L < 0) L = 0;
    return L;
}

template <bool MATRIX>
int dispatch(int b, int n, int m, const float *src, float *temp, int *idx, cudaStream_t stream) {
    if (b < 0 || n < 1 || m < 0 || !src || !temp || !idx) return PDAB_EINVAL;
    if (b == 0 || m == 0) return 0;
    int CL = 1;
    int per = pdab::div_up(n, kThreads);
    while (per > 16 && CL < kMaxCluster) {
        CL *= 2;
        per = pdab::div_up(n, kThreads * CL);
    }
    if (per > 16 || (MATRIX && CL > 1)) return PDAB_EUNSUPPORTED;
    if (!MATRIX && n > 16384) {
        // A serial chain: more SMs per scene shorten every step (fewer points per thread) as long as the clusters of all
        // scenes are resident together; grow the cluster while the grid still fits the 148 SMs.
        // (clusters of 16 are placed one per GPC at best: stop at the portable size 8 unless the batch is tiny)
        int cl_cap = b <= 4 ? kMaxCluster : 8;
        if (cl_cap > g_max_cluster) cl_cap = g_max_cluster;
        while (CL < cl_cap && b * CL * 2 <= pdab::kNumSMs && per > 1) {
            CL *= 2;
            per = pdab::div_up(n, kThreads * CL);
        }
    }
    const int L = ref_log2_block(n);
    if (!MATRIX && n >= 1024 && n <= 16384 && m > 64) {
        // spatially pruned variant: same results, ~N ln m instead of N m point updates
        const int rc = pdab::fps_pruned(b, n, m, src, temp, idx, L, 1, stream);
        if (rc != PDAB_EUNSUPPORTED) return rc;
    }
    if (!MATRIX && n > 16384 && m > 64) {
        // clustered pruned variant: CTA r prunes its own slice of <= 16384 points, the CL local winners are exchanged over
        // DSMEM.  The smallest cluster that holds the scene, doubled once (8192-point slices: shorter local steps) when
        // the policy (pdab_set_fps_max_cluster) and the SM count allow.
        int clp = 1;
        while (clp * 16384 < n) clp *= 2;
        if (clp * 2 <= g_max_cluster && b * clp * 2 <= pdab::kNumSMs) clp *= 2;
        if (clp <= kMaxCluster) {
            const int rc = pdab::fps_pruned(b, n, m, src, temp, idx, L, clp, stream);
            if (rc != PDAB_EUNSUPPORTED) return rc;
        }
    }
    if (per <= 1) return launch<1, MATRIX>(b, n, m, src, temp, idx, L, CL, stream);
    if (per <= 2) return launch<2, MATRIX>(b, n, m, src, temp, idx, L, CL, stream);
    if (per <= 4) return launch<4, MATRIX>(b, n, m, src, temp, idx, L, CL, stream);
    if (per <= 8) return launch<8, MATRIX>(b, n, m, src, temp, idx, L, CL, stream);
    return launch<16, MATRIX>(b, n, m, src, temp, idx, L, CL, stream);
}

}  // namespace

extern "C" int pdab_set_fps_max_cluster(int n) {
    if (n < 1 || n > kMaxCluster) return PDAB_EINVAL;
    g_max_cluster = n;
    return 0;
}

extern "C" int pdab_fps(int b, int n, int m, const float *xyz, float *temp, int *idx, pdab_stream_t stream) {
    return dispatch<false>(b, n, m, xyz, temp, idx, pdab::to_stream(stream));
}

extern "C" int pdab_fps_with_dist(int b, int n, int m, const float *dist, float *temp, int *idx,
                                  pdab_stream_t stream) {
    return dispatch<true>(b, n, m, dist, temp, idx, pdab::to_stream(stream));
}
