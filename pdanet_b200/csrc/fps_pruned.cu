// Pruned, batched farthest-point sampling for sm_100a: one CTA per scene for N <= 16384, a thread-block cluster of CL CTAs per
// scene beyond that (CTA r owns the contiguous slice [r * ceil(N / CL), ...) of the scene, Morton-sorted and pruned locally; the
// CTAs' candidates are exchanged through distributed shared memory — st.async into every peer's buffer, completion on the
// receiver's mbarrier, no cluster barrier in the loop — and every CTA takes the same decisions).
//
// FPS is a serial chain of m-1 steps; the plain kernel (fps.cu) touches all N points in every step although a new sample only
// lowers the min-distance of the points in its own neighbourhood (about N/j of them at step j).  Two things are done about it.
//
// PRUNING.  The points of a scene are Morton-sorted once into buckets of P spatially coherent points, ONE BUCKET PER THREAD
// (min-distances in that thread's registers).  A new sample is tested against the bounding sphere of a warp's 32 buckets (lane r
// tests pending sample r: one test for a whole batch), then every thread tests its bucket's box: if the box is provably
// farther than the bucket's largest min-distance, no point in it can change and the bucket is skipped;
// surviving threads update their P points (coordinates from shared memory) and refresh their cached (max, tie-key, position)
// candidate with no cross-lane traffic.  Total work drops from N*m to about N*ln(m) point updates.
//
// ROUNDS (single CTA, and clusters of 2 / 4 / 8 with the list exchanged over DSMEM).  What is left per step is a fixed latency chain — tests, 2-level argmax, a CTA barrier, the
// winner's coordinates: ~1 650 cycles however little changes.  The chain is cut by taking SEVERAL samples per barrier, exactly:
// every thread whose best point reaches a threshold tau pushes it (key, coordinates, and `sec` = the largest min-distance among
// the thread's OTHER points) into a shared list; after one barrier every warp ranks the <= 32 candidates c_1 > c_2 > ... by the
// argmax key and accepts the longest prefix in which every c_i
//   (a) keeps its min-distance when c_1 .. c_{i-1} are inserted:  fl(|c_j - c_i|^2) >= d(c_i) for all j < i (the very
//       expression and operand order of the update, so fminf would return d(c_i) bit for bit), and
//   (b) stays ahead of every point that is not in the list: points of non-pushing threads are < tau <= d(c_i); points of a
//       pushing thread j other than its candidate are <= sec_j, and sec_j < d(c_i) is required for all j ranked above i.
// Insertions only lower min-distances, so under (a) and (b) c_i is the argmax after c_1 .. c_{i-1} went in — the sequence is
// the one-at-a-time sequence (c_1 is always accepted: it is the global argmax).  The accepted samples are then applied together
// (fminf commutes, so `temp` is identical too).  tau follows the last accepted value through a multiplicative gap steered to
// keep 10-28 candidates in the list; an empty or overflowing list (massive ties, all-equal clouds) falls back to the plain
// 2-level argmax for that round.  On LiDAR-like clouds a round accepts ~10 samples (tools/fps_batch_sim.c replays the rule on
// the host against the oracle: 16384 -> 4096 in 385 rounds; a band of 6-20 gave 454 rounds and 7 % more time).
//
// Exactness (same contract as fps.cu / include/pdab.h): every distance that IS evaluated uses the
// reference's compiled fp32 op order; skipping is conservative — a bucket is skipped only when
// fl(LB)*(1-2^-20) >= max min-distance of the bucket, where LB is the squared distance to the box;
// the fp32 evaluation errors of LB and of a point distance are each below 4*2^-24 relative, so a
// skipped point satisfies fl(d) >= its stored min-distance and fminf would have left it unchanged.
// Inside a round the tests use the bucket maxima of the round's start: min-distances only fall, so stale maxima only test more.
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


using pdab::pack2;
using pdab::unpack2;
using pdab::sub2;
using pdab::mul2;
using pdab::fma2;

struct __align__(16) Candidate {
    unsigned long long key;  // [dist bits | ~tiekey]; 0 = no candidate
    float x, y, z;
    float pad;
};
constexpr int kMaxCluster = 16;

// Phase timers and round statistics of the sampling loop (tools/fps_phase_probe.cu defines PDAB_FPS_TIMERS and includes this
// file; never defined in the library build).  A clock read is ordered after `dep` by a branch on it.
#ifdef PDAB_FPS_TIMERS
__device__ unsigned long long g_fps_phase[16];
#define PDAB_FPS_TICK(slot, dep)                       \
    do {                                               \
        if ((dep) != 0x7fc12345u) {                    \
            const unsigned now_ = (unsigned)clock();   \
            acc_[slot] += now_ - tick_;                \
            tick_ = now_;                              \
        }                                              \
    } while (0)
#define PDAB_FPS_COUNT(slot, v) acc_[slot] += (v)
#else
#define PDAB_FPS_TICK(slot, dep) do {} while (0)
#define PDAB_FPS_COUNT(slot, v) do {} while (0)
#endif

struct __align__(16) WarpBest {
    unsigned long long key;
    int pos;
    int pad;
};

// One entry of a round's candidate list: a thread's best point and the bound on the rest of that thread's points.
struct __align__(16) RoundCand {
    unsigned long long key;  // [dist bits | ~tiekey]
    float x, y;
    float z, sec;            // sec: largest min-distance among the thread's other points (-1: none)
    float pad0, pad1;
};
// A candidate after ranking, stored at its rank.
struct __align__(8) RankedCand {
    float x, y, z, val;      // val: its min-distance
    unsigned low;            // ~tiekey
    int blocked;             // a better candidate moves it, or hides a point that may pass it
};
constexpr int kMaxList = 32;     // one candidate per lane in the merge
#ifndef PDAB_FPS_LIST_LO      // tools/fps_phase_probe.cu sweeps the band; the library build uses the defaults
#define PDAB_FPS_LIST_LO 10
#define PDAB_FPS_LIST_HI 28
#endif
constexpr int kListLo = PDAB_FPS_LIST_LO, kListHi = PDAB_FPS_LIST_HI;   // the threshold gap is steered to keep the list length in this band

// NB buckets of P Morton-consecutive points PER THREAD, T threads (capacity P * NB * T points per CTA).
//
// Round 1 kept a 64-point bucket across the 32 lanes of a warp: every surviving bucket cost its warp a serial
// update + 2 REDUX + ballot + shuffle, a warp walked its survivors one after the other, and a step measured 0.87 us
// (1 650 cycles) at N = 16384.  With a bucket inside a thread its min-distances are P registers of ONE lane, its
// running (max, second, tie-key) needs no cross-lane traffic, every surviving bucket of a warp is updated in the same SIMT
// pass, the boxes are those of 8-16 points instead of 64 (tighter pruning), and Morton-adjacent buckets sit in adjacent
// lanes, so a new sample wakes one or two warps.  The NB buckets of a thread are T buckets apart in Morton order, i.e.
// spatially unrelated: a thread's second-best point is as good as a random point's, which keeps bound (b) of the rounds slack.
//
// Shared memory: coordinates as three planes indexed [((q / 4) * NB * T + bucket) * 4 + q % 4] (a thread fetches four points of
// its bucket with one LDS.128 per plane; a warp's 32 buckets are 512 contiguous bytes: conflict-free), the original index of
// every position as u16 behind them: 14 bytes per point, 224 KB at 16384.
// MODE 0: one CTA per scene, rounds.  MODE 1: cluster of 2 / 4 / 8 CTAs per scene, rounds over a list exchanged through
// distributed shared memory.  MODE 2: cluster of any size up to 16, one sample per step (the round-1 protocol).
template <int P, int NB, int T, int MODE>
__global__ void __launch_bounds__(T, 1)
fps_pruned_kernel(int n_scene, int m, const float *__restrict__ xyz_all, float *__restrict__ temp_all,
                  int *__restrict__ idx_all, int L, int CL) {
    constexpr int NBT = NB * T;        // buckets per CTA
    constexpr int CAP = NBT * P;
    constexpr int kWarps = T / 32;
    constexpr bool CLUSTER = MODE != 0;
    constexpr int V = P < 4 ? P : 4;   // points of one bucket that sit side by side in a plane: one LDS.(32 V) fetches them
    static_assert(P % V == 0 && (V == 1 || V == 2 || V == 4), "bucket rows");
    static_assert(kWarps <= 32, "CTA argmax: one lane per warp");
    // position of point q of bucket g in the coordinate planes: rows of V points, a warp's 32 buckets contiguous
    auto pos_of = [](int q, int g) { return ((q / V) * NBT + g) * V + (q % V); };
    static_assert(NB <= 8, "dirty mask");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // region A: sort keys (CAP x u64), later the coordinates (3 x CAP x f32); region B: original index per position
    unsigned long long *skey = reinterpret_cast<unsigned long long *>(smem_raw);
    float *sx = reinterpret_cast<float *>(smem_raw), *sy = sx + CAP, *sz = sy + CAP;
    unsigned short *sorig = reinterpret_cast<unsigned short *>(smem_raw + (size_t)12 * CAP);
    // static shared memory is scarce (3 KB beside the 224 KB of a full scene): one block carved per mode.
    //   steps (MODE 2): red[2][32] (1 KB) | xchg[2][kMaxCluster] (1 KB)
    //   rounds:         list[2][kMaxList] (2 KB) | ranked[kMaxList] (768 B); a fallback round uses the bytes of `ranked` as
    //                   red[32] (512 B) | fxchg[8] (256 B) instead — a round runs one or the other, between its two barriers
    __shared__ __align__(16) unsigned char sstat[2048 + 768];
    __shared__ int sbox[6];
    __shared__ int scount[3];                            // list lengths, rotating: a counter is zeroed a full round before its use
    __shared__ __align__(8) unsigned long long xbar[3];  // mbarriers: the candidates / lists of all CL CTAs have landed in buffer
                                                         // [buf]; [2]: the fallback candidates of a cluster round

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
            for (int i = 0; i < 3; i++)
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
    if (t < 3) scount[t] = 0;
    if (MODE == 2 && t < 64) reinterpret_cast<WarpBest *>(sstat)[t].key = 0ull;   // lanes beyond kWarps never win the CTA argmax
    __syncthreads();
    {
        float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
        for (int k = t; k < n; k += T)
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
    for (int s = t; s < CAP; s += T) {
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
            for (int i = t; i < CAP / 2; i += T) {
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
    // ---- 2. Morton rank s -> position pos_of(s % P, s / P): bucket g owns ranks [P g, P g + P); region A is reused ---------
    {
        constexpr int PER = CAP / T;
        int korig[PER];
#pragma unroll
        for (int q = 0; q < PER; q++) korig[q] = (int)(unsigned)skey[t + q * T];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int s = t + q * T;
            const int pos = pos_of(s % P, s / P);
            const bool valid = s < n;  // valid keys sort before the padding
            sorig[pos] = valid ? (unsigned short)korig[q] : (unsigned short)0;
            sx[pos] = valid ? __ldg(xyz + (size_t)korig[q] * 3 + 0) : 0.f;
            sy[pos] = valid ? __ldg(xyz + (size_t)korig[q] * 3 + 1) : 0.f;
            sz[pos] = valid ? __ldg(xyz + (size_t)korig[q] * 3 + 2) : 0.f;
        }
    }
    __syncthreads();

    // ---- 3. per-thread state: bucket j of this thread is bucket g = j * T + t -------------------------------------------
    float d[NB][P];                                   // min-distances; padding keeps -1 through every fminf
    float blo[NB][3], bhi[NB][3];                     // bounding box
    float wcen[NB][3], wrad[NB], wthr[NB];            // bounding sphere of the WARP's 32 buckets j (warp-uniform); wthr: squared
                                                      // centre distance below which a sample can reach one of them
    float bmax[NB], bsec[NB];                         // largest / second largest min-distance of the bucket; -1: none
    unsigned blow[NB];                                // ~tie_key of the bucket's candidate
    int bpos[NB];                                     // its position in the coordinate planes

    // cached candidate of bucket j: largest min-distance, reference tie rule among equals (rare: one LDS otherwise)
    auto refresh = [&](int j, const float (&dj)[P], float best) {
        unsigned eq = 0u;
        float sec = -1.f;
#pragma unroll
        for (int q = 0; q < P; q++) {
            const bool top = dj[q] == best;
            eq |= top ? (1u << q) : 0u;
            sec = fmaxf(sec, top ? -1.f : dj[q]);
        }
        bmax[j] = best;
        bsec[j] = (eq & (eq - 1)) ? best : sec;       // several points at the maximum: the runner-up equals it
        if (best >= 0.f) {
            const int g = j * T + t;
            int q = __ffs(eq) - 1;
            eq &= eq - 1;
            bpos[j] = pos_of(q, g);
            blow[j] = ~tie_key(base + (int)sorig[bpos[j]], L);
            while (eq) {
                q = __ffs(eq) - 1;
                eq &= eq - 1;
                const unsigned c = ~tie_key(base + (int)sorig[pos_of(q, g)], L);
                if (c > blow[j]) {
                    blow[j] = c;
                    bpos[j] = pos_of(q, g);
                }
            }
        }
    };
    // a sample farther than sqrt(largest min-distance) + radius from the centre of the warp's block cannot lower any
    // min-distance in it; the 1e-5 margins dwarf the fp32 rounding of both sides (each below 4 * 2^-24 relative)
    auto warp_reach = [&](int j) {
        const unsigned top = __reduce_max_sync(0xffffffffu, bmax[j] >= 0.f ? __float_as_uint(bmax[j]) + 1u : 0u);
        const float reach = __fmaf_rn(__fsqrt_ru(__uint_as_float(top - 1u)), 1.00001f, wrad[j]);
        wthr[j] = top ? reach * reach * 1.00001f : -1.f;   // no valid point in the block: never reached
    };
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const int g = j * T + t;
        float best = -1.f;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            blo[j][a] = 3.4e38f;
            bhi[j][a] = -3.4e38f;
        }
#pragma unroll
        for (int q = 0; q < P; q++) {
            const int pos = pos_of(q, g);
            const bool valid = g * P + q < n;
            d[j][q] = valid ? temp[sorig[pos]] : -1.f;
            best = fmaxf(best, d[j][q]);
            if (valid) {
                const float p[3] = {sx[pos], sy[pos], sz[pos]};
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    blo[j][a] = fminf(blo[j][a], p[a]);
                    bhi[j][a] = fmaxf(bhi[j][a], p[a]);
                }
            }
        }
        // bounding sphere of the warp's block: centre of its box, radius to its farthest point
        float r2 = 0.f;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float wl = ordered_int_inv(__reduce_min_sync(0xffffffffu, ordered_int(blo[j][a])));
            const float wh = ordered_int_inv(__reduce_max_sync(0xffffffffu, ordered_int(bhi[j][a])));
            wcen[j][a] = wl <= wh ? 0.5f * (wl + wh) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < P; q++) {
            const int pos = pos_of(q, g);
            if (g * P + q < n) r2 = fmaxf(r2, pdab::sqdist3(sx[pos], sy[pos], sz[pos], wcen[j][0], wcen[j][1], wcen[j][2]));
        }
        wrad[j] = __fsqrt_ru(__uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(r2)))) * 1.00001f;
        bpos[j] = pos_of(0, g);
        blow[j] = 0u;
        refresh(j, d[j], best);
        warp_reach(j);
    }
    // thread-level candidate over its NB buckets and the bound on the thread's other points
    unsigned long long mykey = 0ull;
    int mypos = pos_of(0, t);
    float mysec = -1.f;
    auto thread_best = [&]() {
        mykey = 0ull;
        mypos = pos_of(0, t);
        int jb = 0;
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const unsigned long long k = bmax[j] >= 0.f ? (((unsigned long long)__float_as_uint(bmax[j]) << 32) | blow[j]) : 0ull;
            if (k > mykey) {
                mykey = k;
                mypos = bpos[j];
                jb = j;
            }
        }
        mysec = -1.f;
#pragma unroll
        for (int j = 0; j < NB; j++) mysec = fmaxf(mysec, j == jb ? bsec[j] : bmax[j]);
    };
    thread_best();
    // the P points of bucket j against one sample (coordinates: V points per load and plane)
    auto update_bucket = [&](int j, float x1, float y1, float z1) {
        const int g = j * T + t;
        const unsigned long long xx = pack2(x1, x1), yy = pack2(y1, y1), zz = pack2(z1, z1);
#pragma unroll
        for (int q0 = 0; q0 < P; q0 += V) {
            float px[V], py[V], pz[V];
            const int pos = pos_of(q0, g);
            if constexpr (V == 4) {
                const float4 a = *reinterpret_cast<const float4 *>(sx + pos), b = *reinterpret_cast<const float4 *>(sy + pos),
                             c = *reinterpret_cast<const float4 *>(sz + pos);
                px[0] = a.x, px[1] = a.y, px[2] = a.z, px[3] = a.w;
                py[0] = b.x, py[1] = b.y, py[2] = b.z, py[3] = b.w;
                pz[0] = c.x, pz[1] = c.y, pz[2] = c.z, pz[3] = c.w;
            } else if constexpr (V == 2) {
                const float2 a = *reinterpret_cast<const float2 *>(sx + pos), b = *reinterpret_cast<const float2 *>(sy + pos),
                             c = *reinterpret_cast<const float2 *>(sz + pos);
                px[0] = a.x, px[1] = a.y;
                py[0] = b.x, py[1] = b.y;
                pz[0] = c.x, pz[1] = c.y;
            } else {
                px[0] = sx[pos], py[0] = sy[pos], pz[0] = sz[pos];
            }
            if constexpr (V >= 2) {
                // two points per instruction (FADD2 / FMUL2 / FFMA2): each half is the scalar op with the same rounding, in the
                // order of pdab::sqdist3
#pragma unroll
                for (int u = 0; u < V; u += 2) {
                    const unsigned long long dx = sub2(pack2(px[u], px[u + 1]), xx), dy = sub2(pack2(py[u], py[u + 1]), yy),
                                             dz = sub2(pack2(pz[u], pz[u + 1]), zz);
                    float e0, e1;
                    unpack2(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), e0, e1);
                    d[j][q0 + u] = fminf(e0, d[j][q0 + u]);
                    d[j][q0 + u + 1] = fminf(e1, d[j][q0 + u + 1]);
                }
            } else {
                d[j][q0] = fminf(pdab::sqdist3(px[0], py[0], pz[0], x1, y1, z1), d[j][q0]);
            }
        }
    };
    auto block_near = [&](int j, float x1, float y1, float z1) {   // can the sample reach the warp's block j at all
        return pdab::sqdist3(wcen[j][0], wcen[j][1], wcen[j][2], x1, y1, z1) < wthr[j];
    };
    auto box_hit = [&](int j, float x1, float y1, float z1) {
        const float ex = fmaxf(fmaxf(blo[j][0] - x1, x1 - bhi[j][0]), 0.f);
        const float ey = fmaxf(fmaxf(blo[j][1] - y1, y1 - bhi[j][1]), 0.f);
        const float ez = fmaxf(fmaxf(blo[j][2] - z1, z1 - bhi[j][2]), 0.f);
        const float lb = ex * ex + ey * ey + ez * ez;
        return !(lb * 0.99999905f >= bmax[j]);
    };
    // one sample against this thread's buckets: sphere, then box, against the bucket's largest min-distance (of the last refresh)
    auto apply = [&](float x1, float y1, float z1) -> unsigned {
        unsigned dirty = 0u;
#pragma unroll
        for (int j = 0; j < NB; j++)
            if (block_near(j, x1, y1, z1) && box_hit(j, x1, y1, z1)) {
                update_bucket(j, x1, y1, z1);
                dirty |= 1u << j;
            }
        return dirty;
    };
    auto refresh_dirty = [&](unsigned dirty) {
#pragma unroll
        for (int j = 0; j < NB; j++)
            if (dirty & (1u << j)) {
                float mx[P];   // pairwise maximum: log2(P) dependent steps instead of P
#pragma unroll
                for (int q = 0; q < P; q++) mx[q] = d[j][q];
#pragma unroll
                for (int w = P / 2; w >= 1; w >>= 1)
#pragma unroll
                    for (int q = 0; q < w; q++) mx[q] = fmaxf(mx[q], mx[q + w]);
                refresh(j, d[j], mx[0]);
            }
        if (dirty) thread_best();
        const unsigned wd = __reduce_or_sync(0xffffffffu, dirty);   // every lane of the warp calls this
#pragma unroll
        for (int j = 0; j < NB; j++)
            if (wd & (1u << j)) warp_reach(j);
    };

    if (t == 0 && rank == 0) idxs[0] = 0;
#ifdef PDAB_FPS_TIMERS
    unsigned acc_[12] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    unsigned tick_ = (unsigned)clock();
#endif

    // mbarrier helpers of the cluster modes (a CTA waits on its own barrier for the bytes its peers st.async into it)
    auto bar_expect = [&](unsigned bar, int bytes) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    };
    auto bar_wait = [&](unsigned bar, unsigned parity) {
        unsigned ok = 0;
        const long long t0 = clock64();
        while (!ok) {
            asm volatile(
                "{\n.reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n}"
                : "=r"(ok)
                : "r"(bar), "r"(parity)
                : "memory");
            if (!ok && clock64() - t0 > 4000000000LL) __trap();  // never hang the GPU on a protocol bug
        }
    };
    // 16 bytes into the same shared-memory offset of CTA `peer`, completing on that CTA's barrier
    auto send16 = [&](unsigned local_addr, unsigned local_bar, int peer, unsigned a, unsigned b, unsigned c, unsigned e) {
        unsigned dst, rbar;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(local_addr), "r"(peer));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(local_bar), "r"(peer));
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
                     "r"(a), "r"(b), "r"(c), "r"(e), "r"(rbar)
                     : "memory");
    };

    if constexpr (MODE != 2) {
        // ======================= rounds: several samples per barrier pair (header: ROUNDS) =================================
        // MODE 1: the list is global — CTA r owns slots [r S, r S + S), S = kMaxList / CL, fills them from its own threads and
        // sends them to every CTA of the cluster (zero-padded; a CTA with more than S candidates sends an overflow mark), so
        // all CTAs rank the same 32 slots and take the same decisions.  Buffer reuse: a peer sends round r + 2 only after it
        // has consumed round r + 1, which needs this CTA's round r + 1 segment, sent after this CTA's barrier of round r + 1,
        // i.e. after all its threads are through with round r's buffer.
        RoundCand(*list)[kMaxList] = reinterpret_cast<RoundCand(*)[kMaxList]>(sstat);
        RankedCand *ranked = reinterpret_cast<RankedCand *>(sstat + 2048);
        WarpBest *red = reinterpret_cast<WarpBest *>(sstat + 2048);
        Candidate *fxchg = reinterpret_cast<Candidate *>(sstat + 2048 + 512);
        const int S = MODE == 1 ? kMaxList / CL : kMaxList;
        const int seg = (int)rank * S;
        constexpr unsigned long long kOverflow = ~0ull;   // not a key: its distance bits would be a NaN
        // the samples whose insertion is pending live in lanes 0 .. npend-1 of EVERY warp
        float psx = __ldg(xyz0 + 0), psy = __ldg(xyz0 + 1), psz = __ldg(xyz0 + 2);
        int npend = 1, it = 1, buf = 0, cb = 0;
        unsigned lpar = 0u, fpar = 0u;   // parities of the list barriers (bit buf) and of the fallback barrier
        float vref = 0.f, gap = 0.25f;   // tau = vref * (1 - gap); vref = min-distance of the last accepted sample
        while (m > 1) {
            // -- insert the pending samples.  Per block j of the warp: lane r tests sample r against the block's sphere (one
            //    test for the whole batch); the samples that can reach it get the per-bucket box test; then every lane walks
            //    ITS OWN hits (shuffle with a per-lane source), so the warp makes max-over-lanes update passes.
            unsigned dirty = 0u;
#pragma unroll
            for (int j = 0; j < NB; j++) {
                unsigned wm = __ballot_sync(0xffffffffu, lane < npend && block_near(j, psx, psy, psz));
                unsigned mm = 0u;
                while (wm) {   // warp-uniform
                    const int r = __ffs(wm) - 1;
                    wm &= wm - 1u;
                    const float x1 = __shfl_sync(0xffffffffu, psx, r), y1 = __shfl_sync(0xffffffffu, psy, r),
                                z1 = __shfl_sync(0xffffffffu, psz, r);
                    mm |= (box_hit(j, x1, y1, z1) ? 1u : 0u) << r;
                }
                while (__any_sync(0xffffffffu, mm != 0u)) {
                    const int r = mm ? __ffs(mm) - 1 : 0;
                    const float x1 = __shfl_sync(0xffffffffu, psx, r), y1 = __shfl_sync(0xffffffffu, psy, r),
                                z1 = __shfl_sync(0xffffffffu, psz, r);
                    if (mm != 0u) {
                        update_bucket(j, x1, y1, z1);
                        dirty |= 1u << j;
                    }
                    mm &= mm - 1u;
                }
            }
            if (it >= m) break;   // the last batch is inserted up to its last-but-one sample (the reference's temp)
            PDAB_FPS_TICK(0, dirty);
            refresh_dirty(dirty);
            // -- candidates at or above tau -------------------------------------------------------------------------------
            const float myval = mykey ? __uint_as_float((unsigned)(mykey >> 32)) : -1.f;
            const float tau = vref > 0.f ? vref * (1.f - gap) : __int_as_float(0x7f800000);
            const bool push = myval >= tau;
            const unsigned pm = __ballot_sync(0xffffffffu, push);
            if (pm) {
                const int leader = __ffs(pm) - 1;
                int slot = 0;
                if (lane == leader) slot = atomicAdd(&scount[cb], __popc(pm));
                RoundCand c;
                c.key = mykey;
                c.x = sx[mypos];
                c.y = sy[mypos];
                c.z = sz[mypos];
                c.sec = mysec;
                c.pad0 = c.pad1 = 0.f;
                slot = __shfl_sync(0xffffffffu, slot, leader) + __popc(pm & ((1u << lane) - 1u));
                if (push && slot < S) list[buf][seg + slot] = c;
            }
            const int cnext = cb == 2 ? 0 : cb + 1;
            if (t == 0) scount[cnext] = 0;   // last read two rounds ago, next used after the coming barrier
            PDAB_FPS_TICK(1, pm);
            __syncthreads();
            const int nloc = scount[cb];
            unsigned vm;           // valid slots of the list
            bool fallback;         // empty list (tau too high) or an overflowing segment (ties en masse): plain argmax this round
            bool overflow;
            if constexpr (MODE == 1) {
                const unsigned lbar = (unsigned)__cvta_generic_to_shared(&xbar[buf]);
                if (t == 0) bar_expect(lbar, CL * S * 32);
                if (t < 2 * S) {   // lane pair e: the two halves of slot e of this CTA's segment
                    const int e = t >> 1, h = t & 1;
                    const unsigned src = (unsigned)__cvta_generic_to_shared(&list[buf][seg + e]) + 16u * h;
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (e < nloc && nloc <= S) v = *reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(&list[buf][seg + e]) + 16 * h);
                    if (nloc > S && e == 0 && h == 0) v.x = v.y = 0xffffffffu;
                    for (int peer = 0; peer < CL; peer++) send16(src, lbar, peer, v.x, v.y, v.z, v.w);
                }
                bar_wait(lbar, (lpar >> buf) & 1u);
                lpar ^= 1u << buf;
                const unsigned long long k = list[buf][lane].key;
                overflow = __any_sync(0xffffffffu, k == kOverflow);
                vm = __ballot_sync(0xffffffffu, k != 0ull && k != kOverflow);
            } else {
                overflow = nloc > kMaxList;
                vm = nloc >= 32 ? 0xffffffffu : ((1u << nloc) - 1u);
            }
            const int nc = __popc(vm);
            fallback = overflow || nc == 0;
            PDAB_FPS_TICK(2, (unsigned)nc);
            int A;
            if (fallback) {
                // -- plain 2-level argmax ----------------------------------------------------------------------------------------
                const unsigned long long wkey = warp_max_u64(mykey);
                const unsigned owner = __ballot_sync(0xffffffffu, mykey == wkey);
                const int wpos = __shfl_sync(0xffffffffu, mypos, __ffs(owner) - 1);
                if (lane == 0) {
                    red[warp].key = wkey;
                    red[warp].pos = wpos;
                }
                __syncthreads();
                const unsigned long long rkey = lane < kWarps ? red[lane].key : 0ull;
                const int rpos = lane < kWarps ? red[lane].pos : 0;
                unsigned long long best = warp_max_u64(rkey);
                const unsigned src = __ballot_sync(0xffffffffu, rkey == best);
                const int slot = __shfl_sync(0xffffffffu, rpos, __ffs(src) - 1);
                psx = sx[slot];
                psy = sy[slot];
                psz = sz[slot];
                if constexpr (MODE == 1) {
                    // the CL local winners, as in the step protocol (fxchg is written by the peers of THIS fallback round only:
                    // they get here after this CTA's list of the round, sent after its threads left the previous fallback)
                    const unsigned fbar = (unsigned)__cvta_generic_to_shared(&xbar[2]);
                    if (t == 0) bar_expect(fbar, CL * 32);
                    if (t < CL) {
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(&fxchg[rank]);
                        const bool any = best != 0ull;
                        send16(dst, fbar, t, (unsigned)best, (unsigned)(best >> 32), any ? __float_as_uint(psx) : 0u,
                               any ? __float_as_uint(psy) : 0u);
                        send16(dst + 16u, fbar, t, any ? __float_as_uint(psz) : 0u, 0u, 0u, 0u);
                    }
                    bar_wait(fbar, fpar);
                    fpar ^= 1u;
                    const unsigned long long mine = lane < CL ? fxchg[lane].key : 0ull;
                    best = warp_max_u64(mine);
                    const int wi = __ffs(__ballot_sync(0xffffffffu, mine == best && lane < CL)) - 1;
                    psx = fxchg[wi].x;
                    psy = fxchg[wi].y;
                    psz = fxchg[wi].z;
                }
                if (t == 0 && rank == 0) idxs[it] = tie_key_decode(~(unsigned)best, L);
                vref = __uint_as_float((unsigned)(best >> 32));
                gap = overflow ? fmaxf(gap * 0.25f, 1e-7f) : fminf(gap * 4.f, 0.5f);
                A = 1;
                PDAB_FPS_COUNT(overflow ? 6 : 5, 1);
            } else {
                // -- rank the candidates: warp w takes slots w, w + kWarps, ...; lane j holds slot j ---------------------------
                const bool valid = (vm >> lane) & 1u;
                const RoundCand cj = list[buf][lane];
                for (int i = warp; i < kMaxList; i += kWarps) {
                    if (!((vm >> i) & 1u)) continue;
                    const RoundCand me = list[buf][i];   // broadcast
                    const float vme = __uint_as_float((unsigned)(me.key >> 32));
                    const bool gt = valid && cj.key > me.key;
                    // (a) the insertion of c_j must leave d(me) as it is — the update's own expression, point first —
                    // (b) and c_j's thread must hold nothing at or above me
                    const float dd = pdab::sqdist3(me.x, me.y, me.z, cj.x, cj.y, cj.z);
                    const unsigned above = __ballot_sync(0xffffffffu, gt);
                    const unsigned block = __ballot_sync(0xffffffffu, gt && (dd < vme || cj.sec >= vme));
                    if (lane == 0) {
                        RankedCand o;
                        o.x = me.x;
                        o.y = me.y;
                        o.z = me.z;
                        o.val = vme;
                        o.low = (unsigned)me.key;
                        o.blocked = block != 0u;
                        ranked[__popc(above)] = o;   // keys are distinct: the ranks are a permutation of 0 .. nc-1
                    }
                }
                __syncthreads();
                // -- accept the longest prefix without a blocked candidate; lane r keeps the sample of rank r ----------------
                const RankedCand mine = ranked[lane < nc ? lane : 0];
                A = __reduce_min_sync(0xffffffffu, (lane < nc && mine.blocked) ? lane : nc);
                A = min(A, m - it);
                psx = mine.x;
                psy = mine.y;
                psz = mine.z;
                if (warp == 0 && rank == 0 && lane < A) idxs[it + lane] = tie_key_decode(~mine.low, L);
                vref = __shfl_sync(0xffffffffu, mine.val, A - 1);
                if (nc < kListLo) gap = fminf(gap * 1.5f, 0.5f);
                else if (nc > kListHi) gap = fmaxf(gap * (1.f / 1.5f), 1e-7f);
                PDAB_FPS_COUNT(7, (unsigned)nc);
            }
            PDAB_FPS_TICK(3, __float_as_uint(psz) + (unsigned)A);
            PDAB_FPS_COUNT(4, 1);
            it += A;
            npend = it == m ? A - 1 : A;
            buf ^= 1;
            cb = cnext;
        }
    } else {
        // ======================= MODE 2: one sample per step, candidates exchanged over DSMEM ==============================
        WarpBest(*red)[32] = reinterpret_cast<WarpBest(*)[32]>(sstat);
        Candidate(*xchg)[kMaxCluster] = reinterpret_cast<Candidate(*)[kMaxCluster]>(sstat + 1024);
        unsigned long long wkey = 0ull;
        int wpos = pos_of(0, t);
        auto warp_best = [&]() {
            wkey = warp_max_u64(mykey);
            const unsigned owner = __ballot_sync(0xffffffffu, mykey == wkey);
            wpos = __shfl_sync(0xffffffffu, mypos, __ffs(owner) - 1);
        };
        warp_best();
        float x1 = __ldg(xyz0 + 0), y1 = __ldg(xyz0 + 1), z1 = __ldg(xyz0 + 2);
        int buf = 0;
        for (int it = 1; it < m; it++) {
            const unsigned dirty = apply(x1, y1, z1);
            refresh_dirty(dirty);
            // -- argmax over cached candidates: thread -> warp -> CTA (the warp level only where a bucket changed) -------------
            if (__any_sync(0xffffffffu, dirty != 0u)) warp_best();
            if (lane == 0) {
                red[buf][warp].key = wkey;
                red[buf][warp].pos = wpos;
            }
            __syncthreads();
            const unsigned long long rkey = red[buf][lane].key;
            const int rpos = red[buf][lane].pos;
            const unsigned long long best = warp_max_u64(rkey);
            const unsigned src = __ballot_sync(0xffffffffu, rkey == best);
            const int slot = __shfl_sync(0xffffffffu, rpos, __ffs(src) - 1);
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
            if (t == 0 && rank == 0) idxs[it] = tie_key_decode(~(unsigned)w, L);
            x1 = xchg[buf][wi].x;
            y1 = xchg[buf][wi].y;
            z1 = xchg[buf][wi].z;
            buf ^= 1;
        }
    }

#ifdef PDAB_FPS_TIMERS
    if (t == 0 && blockIdx.x == 0)
        for (int i = 0; i < 8; i++) g_fps_phase[i] = acc_[i];
#endif
#pragma unroll
    for (int j = 0; j < NB; j++)
#pragma unroll
        for (int q = 0; q < P; q++) {
            const int g = j * T + t;
            if (g * P + q < n) temp[sorig[pos_of(q, g)]] = d[j][q];
        }
    // nobody exits while a peer's last st.async may still target its shared memory
    if (CLUSTER) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int P, int NB, int T>
int launch(int b, int n, int m, const float *xyz, float *temp, int *idx, int L, int CL, cudaStream_t stream, bool steps_only = false) {
    constexpr int CAP = P * NB * T;
    const size_t smem = (size_t)12 * CAP + (size_t)2 * CAP;
    if (CL == 1) {
        auto kern = fps_pruned_kernel<P, NB, T, 0>;
        PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<b, T, smem, stream>>>(n, m, xyz, temp, idx, L, 1);
        PDAB_LAUNCH_CHECK();
        return 0;
    }
    // rounds need the 32 list slots split evenly over the CTAs, at least 4 each
    const bool rounds = steps_only ? false : (CL == 2 || CL == 4 || CL == 8);
    auto kern = rounds ? fps_pruned_kernel<P, NB, T, 1> : fps_pruned_kernel<P, NB, T, 2>;
    PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PDAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(b * CL);
    cfg.blockDim = dim3(T);
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

#ifdef PDAB_FPS_PROBE
int g_probe_variant = 0;
#endif

// Returns PDAB_EUNSUPPORTED when the pruned kernel does not cover the size (caller falls back to fps.cu).
// CL > 1: a cluster of CL CTAs per scene, each holding ceil(n / CL) <= 16384 points.
int fps_pruned(int b, int n, int m, const float *xyz, float *temp, int *idx, int L, int CL, cudaStream_t stream) {
    if (CL < 1 || CL > kMaxCluster || n < 1) return PDAB_EUNSUPPORTED;
    const int per_cta = (n + CL - 1) / CL;
    if (per_cta > 16384) return PDAB_EUNSUPPORTED;
#ifdef PDAB_FPS_PROBE
    switch (g_probe_variant) {   // tools/fps_phase_probe.cu: shape sweep
        case 1: return launch<16, 1, 1024>(b, n, m, xyz, temp, idx, L, CL, stream);
        case 2: return launch<16, 2, 512>(b, n, m, xyz, temp, idx, L, CL, stream);
        case 3: return launch<16, 4, 256>(b, n, m, xyz, temp, idx, L, CL, stream);
        case 4: return launch<8, 4, 512>(b, n, m, xyz, temp, idx, L, CL, stream);
        case 5: return launch<8, 8, 256>(b, n, m, xyz, temp, idx, L, CL, stream);
        case 6: return launch<8, 2, 1024>(b, n, m, xyz, temp, idx, L, CL, stream);
        case 11: return launch<4, 1, 1024>(b, n, m, xyz, temp, idx, L, CL, stream);
        case 12: return launch<4, 2, 512>(b, n, m, xyz, temp, idx, L, CL, stream);
        case 13: return launch<4, 4, 256>(b, n, m, xyz, temp, idx, L, CL, stream);
        case 14: return launch<8, 1, 512>(b, n, m, xyz, temp, idx, L, CL, stream);
        case 15: return launch<8, 2, 256>(b, n, m, xyz, temp, idx, L, CL, stream);
        case 16: return launch<16, 1, 256>(b, n, m, xyz, temp, idx, L, CL, stream);
        case 17: return launch<8, 4, 128>(b, n, m, xyz, temp, idx, L, CL, stream);
        case 20: return launch<16, 2, 512>(b, n, m, xyz, temp, idx, L, CL, stream, true);   // cluster: one sample per step
        default: break;
    }
#endif
    // 512 threads; one bucket per thread while the slice allows (measured: <8,1,512> 0.298 ms against <4,2,512> 0.328 ms at
    // 16 x 4096 -> 1024; <16,2,512> and <16,1,1024> tie at 16 x 16384 -> 4096)
    if (per_cta <= 1024) return launch<2, 1, 512>(b, n, m, xyz, temp, idx, L, CL, stream);
    if (per_cta <= 2048) return launch<4, 1, 512>(b, n, m, xyz, temp, idx, L, CL, stream);
    if (per_cta <= 4096) return launch<8, 1, 512>(b, n, m, xyz, temp, idx, L, CL, stream);
    if (per_cta <= 8192) return launch<16, 1, 512>(b, n, m, xyz, temp, idx, L, CL, stream);
    return launch<16, 2, 512>(b, n, m, xyz, temp, idx, L, CL, stream);
}

}  // namespace pdab
