// Pruned farthest-point sampling for sm_100a: one CTA per scene for N <= 16384, a thread-block cluster of CL CTAs per scene
// beyond that (CTA r owns the contiguous slice [r * ceil(N / CL), ...) of the scene, Morton-sorted and pruned locally; the CL
// local winners of a step are exchanged through distributed shared memory — st.async into every peer's buffer, completion on
// the receiver's mbarrier, no cluster barrier in the loop — and every CTA picks the same global winner).
//
// FPS is a serial chain of m-1 steps; the plain kernel (fps.cu) touches all N points in every
// step although a new sample only lowers the min-distance of the points in its own neighbourhood
// (about N/j of them at step j).  Here the points of a scene are Morton-sorted once into buckets
// of 32 spatially coherent points (one warp-wide register each); every step
//   1. tests each bucket's bounding box against the new sample: if the box is provably farther
//      than the bucket's largest min-distance, no point in it can change and the bucket is skipped;
//   2. updates only the surviving buckets (coordinates from shared memory, min-distances in
//      registers) and refreshes their cached (max, tie-key, slot) candidate with two REDUX;
//   3. reduces the 8 warps' cached candidates (REDUX + one named barrier) to the next sample.
// Total work drops from N*m to about N*ln(m) point updates; what remains per step is the fixed
// cost of the box tests and of the 2-level argmax.
//
// Exactness (same contract as fps.cu / include/pdab.h): every distance that IS evaluated uses the
// reference's compiled fp32 op order; skipping is conservative — a bucket is skipped only when
// fl(LB)*(1-2^-20) >= max min-distance of the bucket, where LB is the squared distance to the box;
// the fp32 evaluation errors of LB and of a point distance are each below 4*2^-24 relative, so a
// skipped point satisfies fl(d) >= its stored min-distance and fminf would have left it unchanged.
// The argmax key [dist bits | ~tiekey(k)] is the one fps.cu uses, on ORIGINAL point indices, so
// the reference's tie rule (argmin (bitrev_L(k mod BS), k) over maxima) is preserved under the
// permutation.  Preconditions: finite coordinates, temp >= 0 (the caller fills 1e10).

#include "common.cuh"

namespace {


__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | ml;
}

__device__ __forceinline__ unsigned tie_key(int k, int L) {
    if (L == 0) return (unsigned)k;
    return __brev((unsigned)k & ((1u << L) - 1u)) | ((unsigned)k >> L);
}
__device__ __forceinline__ int tie_key_decode(unsigned key, int L) {
    if (L == 0) return (int)key;
    const unsigned lowmask = (1u << (32 - L)) - 1u;
    return (int)(((key & lowmask) << L) | __brev(key & ~lowmask));
}

__device__ __forceinline__ int ordered_int(float f) {
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ordered_int_inv(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

__device__ __forceinline__ unsigned spread10(unsigned v) {  // 10 bits -> every third bit
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}


struct __align__(16) Candidate {
    unsigned long long key;  // [dist bits | ~tiekey]; 0 = no candidate
    float x, y, z;
    float pad;
};
constexpr int kMaxCluster = 16;

struct __align__(16) WarpBest {
    unsigned long long key;
    int slot;
    int pad;
};

// WARPS warps per CTA, BPW buckets per warp (BPW <= 32).  Capacity = WARPS * BPW * 32 points.
// The per-bucket update is replicated BPW times in the loop body (the min-distances live in
// registers and registers cannot be indexed dynamically), so BPW also sets the code size of the
// loop: BPW = 64 on 8 warps measured 3.4 us/step because the 60 KB body thrashed the instruction
// cache; 16 buckets on 32 warps keeps it near 14 KB.
// PPL points per lane per bucket: a bucket is 32*PPL Morton-consecutive points.
template <int WARPS, int BPW, int PPL, bool CLUSTER>
__global__ void __launch_bounds__(WARPS * 32, 1)
fps_pruned_kernel(int n_scene, int m, const float *__restrict__ xyz_all, float *__restrict__ temp_all,
                  int *__restrict__ idx_all, int L, int CL) {
    constexpr int kWarps = WARPS, kInitThreads = WARPS * 32;
    constexpr int CAP = kWarps * BPW * 32 * PPL;
    constexpr int BPL = 1;  // buckets tested per lane
    static_assert(BPW <= 16, "one tested bucket per lane, 16 switch cases");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // region A: sort keys (CAP x u64), later the coordinates (3 x CAP x f32); region B: original index per slot
    unsigned long long *skey = reinterpret_cast<unsigned long long *>(smem_raw);
    float *sx = reinterpret_cast<float *>(smem_raw), *sy = sx + CAP, *sz = sy + CAP;
    unsigned short *sorig = reinterpret_cast<unsigned short *>(smem_raw + (size_t)12 * CAP);
    __shared__ int sbox[6];
    __shared__ WarpBest red[2][kWarps];
    __shared__ Candidate xchg[2][kMaxCluster];
    __shared__ __align__(8) unsigned long long xbar[2];  // mbarriers: the candidates of all CL CTAs have landed in xchg[buf]

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned rank = 0;
    if (CLUSTER) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int scene = CLUSTER ? blockIdx.x / CL : blockIdx.x;
    // this CTA's slice of the scene: local point k <-> scene point base + k
    const int chunk = CLUSTER ? (n_scene + CL - 1) / CL : n_scene;
    const int base = (int)rank * chunk;
    const int n = max(0, min(chunk, n_scene - base));
    const float *xyz0 = xyz_all + (size_t)scene * n_scene * 3;   // the scene (point 0 seeds the chain)
    const float *xyz = xyz0 + (size_t)base * 3;
    float *temp = temp_all + (size_t)scene * n_scene + base;
    int *idxs = idx_all + (size_t)scene * m;
    if (CLUSTER) {
        if (t == 0) {
            for (int i = 0; i < 2; i++)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&xbar[i])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        // peers resident and their barriers initialised before any DSMEM store (the sort below gives plenty of slack, but
        // correctness must not depend on it)
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }

    // ---- 0. scene bounding box -------------------------------------------------------------
    if (t < 3) sbox[t] = 0x7fffffff;
    else if (t < 6) sbox[t] = (int)0x80000000;
    __syncthreads();
    {
        float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
        for (int k = t; k < n; k += kInitThreads)
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const float v = __ldg(xyz + (size_t)k * 3 + a);
                lo[a] = fminf(lo[a], v);
                hi[a] = fmaxf(hi[a], v);
            }
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const int l = __reduce_min_sync(0xffffffffu, ordered_int(lo[a]));
            const int h = __reduce_max_sync(0xffffffffu, ordered_int(hi[a]));
            if (lane == 0) {
                atomicMin(&sbox[a], l);
                atomicMax(&sbox[3 + a], h);
            }
        }
    }
    __syncthreads();
    const float bx = ordered_int_inv(sbox[0]), by = ordered_int_inv(sbox[1]), bz = ordered_int_inv(sbox[2]);
    const float ext = fmaxf(fmaxf(ordered_int_inv(sbox[3]) - bx, ordered_int_inv(sbox[4]) - by),
                            ordered_int_inv(sbox[5]) - bz);
    const float qscale = ext > 0.f ? 1023.0f / ext : 0.f;  // isotropic: buckets are compact in real space

    // ---- 1. Morton keys + bitonic sort (ascending; padding keys sort last) ---------------------
    for (int s = t; s < CAP; s += kInitThreads) {
        unsigned long long key = ~0ull;
        if (s < n) {
            const unsigned qx = min(1023u, (unsigned)((__ldg(xyz + (size_t)s * 3 + 0) - bx) * qscale));
            const unsigned qy = min(1023u, (unsigned)((__ldg(xyz + (size_t)s * 3 + 1) - by) * qscale));
            const unsigned qz = min(1023u, (unsigned)((__ldg(xyz + (size_t)s * 3 + 2) - bz) * qscale));
            const unsigned code = spread10(qx) | (spread10(qy) << 1) | (spread10(qz) << 2);
            key = ((unsigned long long)code << 32) | (unsigned)s;
        }
        skey[s] = key;
    }
    __syncthreads();
    for (int size = 2; size <= CAP; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = t; i < CAP / 2; i += kInitThreads) {
                const int a = 2 * i - (i & (stride - 1));
                const int b = a + stride;
                const bool up = (a & size) == 0;
                const unsigned long long ka = skey[a], kb = skey[b];
                if ((ka > kb) == up) {
                    skey[a] = kb;
                    skey[b] = ka;
                }
            }
            __syncthreads();
        }
    // ---- 2. slot -> original index, then coordinates in slot order (region A is reused) -------
    int korig[CAP / kInitThreads];
#pragma unroll
    for (int q = 0; q < CAP / kInitThreads; q++) korig[q] = (int)(unsigned)skey[t + q * kInitThreads];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < CAP / kInitThreads; q++) {
        const int s = t + q * kInitThreads;
        const bool valid = s < n;  // valid keys sort before the padding
        sorig[s] = valid ? (unsigned short)korig[q] : (unsigned short)0;
        sx[s] = valid ? __ldg(xyz + (size_t)korig[q] * 3 + 0) : 0.f;
        sy[s] = valid ? __ldg(xyz + (size_t)korig[q] * 3 + 1) : 0.f;
        sz[s] = valid ? __ldg(xyz + (size_t)korig[q] * 3 + 2) : 0.f;
    }
    __syncthreads();

    // ---- 3. per-lane state: min-distances of one point of each owned bucket, cached bucket data -
    // warp w owns buckets b = w + kWarps*i (interleaved: a neighbourhood's buckets spread over warps);
    // lane l holds slot 32*b + l.  Bucket i is TESTED by lane (i % 32), register (i / 32).
    float d[BPW][PPL];
    float blo[BPL][3], bhi[BPL][3], bmax[BPL];
    unsigned long long bkey[BPL];
    int bslot[BPL];
#pragma unroll
    for (int r = 0; r < BPL; r++) {
        bmax[r] = -1.f;
        bkey[r] = 0ull;
        bslot[r] = 0;
#pragma unroll
        for (int a = 0; a < 3; a++) blo[r][a] = bhi[r][a] = 0.f;
    }

    // slot of (bucket i of this warp, row q, this lane)
    auto slot_of = [&](int i, int q) { return ((warp + kWarps * i) * PPL + q) * 32 + lane; };

    auto refresh = [&](int i, const float (&dv)[PPL]) {
        // cached candidate of bucket i: largest min-distance, reference tie rule among equals
        float mv = dv[0];
#pragma unroll
        for (int q = 1; q < PPL; q++) mv = fmaxf(mv, dv[q]);
        const bool valid = mv >= 0.f;
        const unsigned hi = valid ? __float_as_uint(mv) : 0u;
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const bool top = valid && hi == mh;
        unsigned lo = 0u;
        int myq = 0;
        if (top) {
#pragma unroll
            for (int q = 0; q < PPL; q++)
                if (dv[q] == mv) {
                    const unsigned c = ~tie_key(base + (int)sorig[slot_of(i, q)], L);
                    if (c > lo) {
                        lo = c;
                        myq = q;
                    }
                }
        }
        const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
        const unsigned who = __ballot_sync(0xffffffffu, top && lo == ml);
        const int wl = who ? __ffs(who) - 1 : 0;
        const int wq = __shfl_sync(0xffffffffu, myq, wl);
        if (lane == i) {
            const bool any = who != 0u;
            bmax[0] = any ? __uint_as_float(mh) : -1.f;
            bkey[0] = any ? (((unsigned long long)mh << 32) | ml) : 0ull;
            bslot[0] = ((warp + kWarps * i) * PPL + wq) * 32 + wl;
        }
    };

#pragma unroll 1
    for (int i = 0; i < BPW; i++) {  // bounding boxes (rolled: one-off, keeps the code small)
        int l3[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
        int h3[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
        for (int q = 0; q < PPL; q++) {
            const int slot = slot_of(i, q);
            if (slot < n) {
                const float p[3] = {sx[slot], sy[slot], sz[slot]};
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    l3[a] = min(l3[a], ordered_int(p[a]));
                    h3[a] = max(h3[a], ordered_int(p[a]));
                }
            }
        }
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const int l = __reduce_min_sync(0xffffffffu, l3[a]);
            const int h = __reduce_max_sync(0xffffffffu, h3[a]);
            if (lane == i) {
                blo[0][a] = ordered_int_inv(l);
                bhi[0][a] = ordered_int_inv(h);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < BPW; i++) {
#pragma unroll
        for (int q = 0; q < PPL; q++) {
            const int slot = slot_of(i, q);
            d[i][q] = slot < n ? temp[sorig[slot]] : -1.f;
        }
        refresh(i, d[i]);
    }

    int old = 0;
    float x1 = __ldg(xyz0 + 0), y1 = __ldg(xyz0 + 1), z1 = __ldg(xyz0 + 2);
    if (t == 0 && rank == 0) idxs[0] = 0;

    int buf = 0;
    for (int it = 1; it < m; it++) {
        // -- box tests ---------------------------------------------------------------------------
        unsigned active[BPL];
#pragma unroll
        for (int r = 0; r < BPL; r++) {
            const float ex = fmaxf(fmaxf(blo[r][0] - x1, x1 - bhi[r][0]), 0.f);
            const float ey = fmaxf(fmaxf(blo[r][1] - y1, y1 - bhi[r][1]), 0.f);
            const float ez = fmaxf(fmaxf(blo[r][2] - z1, z1 - bhi[r][2]), 0.f);
            const float lb = ex * ex + ey * ey + ez * ez;
            // empty / unowned buckets carry bmax = -1 and are never active
            active[r] = __ballot_sync(0xffffffffu, !(lb * 0.99999905f >= bmax[r]));
        }
        // -- update the surviving buckets -----------------------------------------------------------
        // most warps have nothing to do in most steps; busy ones jump straight to the bodies of their
        // active buckets (indexed branch) instead of walking BPW bit tests spread over the whole loop body
        unsigned todo = active[0];
        while (todo) {
            const int i = __ffs(todo) - 1;
            todo &= todo - 1;
#define PDAB_BUCKET_CASE(I)                                                                                   \
    case I:                                                                                                   \
        if (I < BPW) {                                                                                        \
            _Pragma("unroll") for (int q = 0; q < PPL; q++) {                                                 \
                const int slot = slot_of(I, q);                                                               \
                if (d[I < BPW ? I : 0][q] >= 0.f)                                                             \
                    d[I < BPW ? I : 0][q] =                                                                   \
                        fminf(pdab::sqdist3(sx[slot], sy[slot], sz[slot], x1, y1, z1), d[I < BPW ? I : 0][q]); \
            }                                                                                                 \
            refresh(I, d[I < BPW ? I : 0]);                                                                   \
        }                                                                                                     \
        break;
            switch (i) {
                PDAB_BUCKET_CASE(0) PDAB_BUCKET_CASE(1) PDAB_BUCKET_CASE(2) PDAB_BUCKET_CASE(3)
                PDAB_BUCKET_CASE(4) PDAB_BUCKET_CASE(5) PDAB_BUCKET_CASE(6) PDAB_BUCKET_CASE(7)
                PDAB_BUCKET_CASE(8) PDAB_BUCKET_CASE(9) PDAB_BUCKET_CASE(10) PDAB_BUCKET_CASE(11)
                PDAB_BUCKET_CASE(12) PDAB_BUCKET_CASE(13) PDAB_BUCKET_CASE(14) PDAB_BUCKET_CASE(15)
                default: break;
            }
#undef PDAB_BUCKET_CASE
        }
        // -- argmax over cached candidates: lane -> warp -> CTA -------------------------------------
        unsigned long long mykey = bkey[0];
        int myslot = bslot[0];
#pragma unroll
        for (int r = 1; r < BPL; r++)
            if (bkey[r] > mykey) {
                mykey = bkey[r];
                myslot = bslot[r];
            }
        const unsigned long long wkey = warp_max_u64(mykey);
        const unsigned owner = __ballot_sync(0xffffffffu, mykey == wkey);
        const int wslot = __shfl_sync(0xffffffffu, myslot, __ffs(owner) - 1);
        if (lane == 0) {
            red[buf][warp].key = wkey;
            red[buf][warp].slot = wslot;
        }
        __syncthreads();
        const unsigned long long rkey = lane < kWarps ? red[buf][lane].key : 0ull;
        const int rslot = lane < kWarps ? red[buf][lane].slot : 0;
        const unsigned long long best = warp_max_u64(rkey);
        const unsigned src = __ballot_sync(0xffffffffu, rkey == best && lane < kWarps);
        const int slot = __shfl_sync(0xffffffffu, rslot, __ffs(src) - 1);
        if (!CLUSTER) {
            old = tie_key_decode(~(unsigned)best, L);
            x1 = sx[slot];
            y1 = sy[slot];
            z1 = sz[slot];
        } else {
            // every CTA pushes its 32-byte candidate into each peer's xchg[buf][rank] with st.async, which completes on the
            // RECEIVER's mbarrier; a CTA waits for CL x 32 bytes on its own barrier.  Buffer reuse is safe: a peer can only
            // send step it+2 after it has consumed step it+1, which needs this CTA's step it+1 candidate, which is sent
            // after this CTA's block barrier of step it+1, i.e. after all its threads have read step it's buffer.
            const unsigned bar_local = (unsigned)__cvta_generic_to_shared(&xbar[buf]);
            if (t == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_local), "r"(CL * 32)
                             : "memory");
            if (t < CL) {
                float cx = 0.f, cy = 0.f, cz = 0.f;
                if (best != 0ull) {
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
                    "r"((unsigned)best), "r"((unsigned)(best >> 32)), "r"(__float_as_uint(cx)), "r"(__float_as_uint(cy)),
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
    for (int i = 0; i < BPW; i++)
#pragma unroll
        for (int q = 0; q < PPL; q++) {
            const int slot = slot_of(i, q);
            if (slot < n) temp[sorig[slot]] = d[i][q];
        }
    // nobody exits while a peer's last st.async may still target its shared memory
    if (CLUSTER) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int WARPS, int BPW, int PPL>
int launch(int b, int n, int m, const float *xyz, float *temp, int *idx, int L, int CL, cudaStream_t stream) {
    constexpr int CAP = WARPS * BPW * 32 * PPL;
    constexpr int kInitThreads = WARPS * 32;
    const size_t smem = (size_t)12 * CAP + (size_t)2 * CAP;
    if (CL == 1) {
        auto kern = fps_pruned_kernel<WARPS, BPW, PPL, false>;
        PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<b, kInitThreads, smem, stream>>>(n, m, xyz, temp, idx, L, 1);
        PDAB_LAUNCH_CHECK();
        return 0;
    }
    auto kern = fps_pruned_kernel<WARPS, BPW, PPL, true>;
    PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(b * CL);
    cfg.blockDim = dim3(kInitThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PDAB_CUDA(cudaLaunchKernelEx(&cfg, kern, n, m, xyz, temp, idx, L, CL));
    PDAB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

namespace pdab {

// Returns PDAB_EUNSUPPORTED when the pruned kernel does not cover the size (caller falls back to fps.cu).
// CL > 1: a cluster of CL CTAs per scene, each holding ceil(n / CL) <= 16384 points.
int fps_pruned(int b, int n, int m, const float *xyz, float *temp, int *idx, int L, int CL, cudaStream_t stream) {
    if (CL < 1 || CL > kMaxCluster || n < 1) return PDAB_EUNSUPPORTED;
    const int per_cta = (n + CL - 1) / CL;
    if (per_cta > 16384) return PDAB_EUNSUPPORTED;
    // 16 warps, 64-point buckets; buckets per warp grow with the slice (the other warp / bucket shapes were measured slower,
    // profiles/r01_microbench.txt)
    const int sz = per_cta <= 2048 ? 0 : per_cta <= 4096 ? 1 : per_cta <= 8192 ? 2 : 3;
    if (sz == 0) return launch<16, 2, 2>(b, n, m, xyz, temp, idx, L, CL, stream);
    if (sz == 1) return launch<16, 4, 2>(b, n, m, xyz, temp, idx, L, CL, stream);
    if (sz == 2) return launch<16, 8, 2>(b, n, m, xyz, temp, idx, L, CL, stream);
    return launch<16, 16, 2>(b, n, m, xyz, temp, idx, L, CL, stream);
}

}  // namespace pdab
