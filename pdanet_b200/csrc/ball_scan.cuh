// Thread-per-centre ordered ball scan shared by the fused kernels
// (pda_group.cu, sa_fused.cu).  Same search as ball_query.cu, but the hit list
// goes to shared memory, transposed (slot-major) so that the CTA can later walk
// it without bank conflicts, and unfilled slots are completed in place.
#pragma once
#include "common.cuh"

namespace pdab {

constexpr int kScanTile = 1024;  // points per staging tile (float4, 16 KB); callers allocate kScanTile + 8 (padding to 8)

// All threads of the CTA must call this (it contains CTA barriers).
//   THREADS : number of centres scanned by the CTA (threads 0..THREADS-1 own one each);
//             the CTA may be larger (blockDim.x >= THREADS): the extra threads only help stage tiles
//   STRIDE  : row pitch of sidx (>= THREADS; THREADS+1 staggers banks for slot-major readers)
//   sidx    : shared int[nsample * STRIDE]; on return sidx[s * STRIDE + t] holds the
//             s-th neighbour of thread t's centre (slot fill and empty-ball -> 0 applied,
//             i.e. exactly what the reference's pre-zeroed idx row would hold).
template <int THREADS, int STRIDE>
__device__ __forceinline__ void ball_scan_to_smem(int n, const float *__restrict__ xyz, bool active, float cx,
                                                  float cy, float cz, float r2, int nsample, float4 *tile,
                                                  int *sidx) {
    const int t = threadIdx.x;
    const int nthreads = blockDim.x;
    active = active && t < THREADS;
    int cnt = active ? 0 : nsample;
    for (int base = 0; base < n; base += kScanTile) {
        const int len = min(kScanTile, n - base);
        const int len8 = (len + 7) & ~7;
        __syncthreads();
        for (int i = t; i < len8; i += nthreads) {
            if (i < len) {
                const float *p = xyz + (size_t)(base + i) * 3;
                tile[i] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
            } else {
                tile[i] = make_float4(1e30f, 1e30f, 1e30f, 0.f);  // padding: squared distance overflows to +inf, never a hit
            }
        }
        __syncthreads();
        if (!__all_sync(0xffffffffu, cnt >= nsample)) {
            // eight independent distance tests in flight, one branch for the (rare) hit
            for (int i = 0; i < len8; i += 8) {
                float d2[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const float4 p = tile[i + u];
                    d2[u] = sqdist3(cx, cy, cz, p.x, p.y, p.z);
                }
                const float mn = fminf(fminf(fminf(d2[0], d2[1]), fminf(d2[2], d2[3])),
                                       fminf(fminf(d2[4], d2[5]), fminf(d2[6], d2[7])));
                if (mn < r2 && cnt < nsample) {
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        if (d2[u] < r2 && cnt < nsample) {
                            sidx[cnt * STRIDE + t] = base + i + u;
                            cnt++;
                        }
                    }
                }
            }
        }
        if (__syncthreads_and(cnt >= nsample)) break;
    }
    if (active) {
        const int first = cnt > 0 ? sidx[t] : 0;  // empty ball: the pre-zeroed row groups point 0
        for (int l = cnt; l < nsample; l++) sidx[l * STRIDE + t] = first;
    } else if (t < THREADS) {
        for (int l = 0; l < nsample; l++) sidx[l * STRIDE + t] = 0;
    }
    __syncthreads();
}

// Two radii in ONE pass over the cloud (the two scales of an SA layer share centres and points), eight independent
// distance tests in flight per thread and a single branch for the rare hit.  Same results as two calls of
// ball_scan_to_smem: hits are taken in index order, each list stops at its own nsample.
//   tile : shared float4[kScanTile + 8]; sidx_a / sidx_b : shared int[nsample_x * STRIDE]
template <int THREADS, int STRIDE>
__device__ __forceinline__ void ball_scan2_to_smem(int n, const float *__restrict__ xyz, bool active, float cx, float cy,
                                                   float cz, float r2a, int ns_a, float r2b, int ns_b, float4 *tile,
                                                   int *__restrict__ sidx_a, int *__restrict__ sidx_b) {
    const int t = threadIdx.x;
    const int nthreads = blockDim.x;
    active = active && t < THREADS;
    int ca = active ? 0 : ns_a, cb = active ? 0 : ns_b;
    const float r2max = fmaxf(r2a, r2b);
    for (int base = 0; base < n; base += kScanTile) {
        const int len = min(kScanTile, n - base);
        const int len8 = (len + 7) & ~7;
        __syncthreads();
        for (int i = t; i < len8; i += nthreads) {
            if (i < len) {
                const float *p = xyz + (size_t)(base + i) * 3;
                tile[i] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
            } else {
                tile[i] = make_float4(1e30f, 1e30f, 1e30f, 0.f);  // padding: squared distance overflows to +inf, never a hit
            }
        }
        __syncthreads();
        if (!__all_sync(0xffffffffu, ca >= ns_a && cb >= ns_b)) {
            for (int i = 0; i < len8; i += 8) {
                float d2[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const float4 p = tile[i + u];
                    d2[u] = sqdist3(cx, cy, cz, p.x, p.y, p.z);
                }
                const float mn = fminf(fminf(fminf(d2[0], d2[1]), fminf(d2[2], d2[3])),
                                       fminf(fminf(d2[4], d2[5]), fminf(d2[6], d2[7])));
                if (mn < r2max && (ca < ns_a || cb < ns_b)) {
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        if (d2[u] < r2a && ca < ns_a) {
                            sidx_a[ca * STRIDE + t] = base + i + u;
                            ca++;
                        }
                        if (d2[u] < r2b && cb < ns_b) {
                            sidx_b[cb * STRIDE + t] = base + i + u;
                            cb++;
                        }
                    }
                }
            }
        }
        if (__syncthreads_and(ca >= ns_a && cb >= ns_b)) break;
    }
    if (active) {
        const int fa = ca > 0 ? sidx_a[t] : 0;  // empty ball: the pre-zeroed row groups point 0
        for (int l = ca; l < ns_a; l++) sidx_a[l * STRIDE + t] = fa;
        const int fb = cb > 0 ? sidx_b[t] : 0;
        for (int l = cb; l < ns_b; l++) sidx_b[l * STRIDE + t] = fb;
    } else if (t < THREADS) {
        for (int l = 0; l < ns_a; l++) sidx_a[l * STRIDE + t] = 0;
        for (int l = 0; l < ns_b; l++) sidx_b[l * STRIDE + t] = 0;
    }
    __syncthreads();
}

// CTA-wide scan for 32 centres (lane = centre): the cloud is cut into WARPS contiguous slices, warp w scans slice w for
// all 32 centres (every thread of the CTA tests points, instead of one warp in WARPS), and the per-slice hit lists are
// concatenated in slice order — the same list a single in-order scan produces, since hits are taken in index order and
// the list stops at nsample.  All threads of the CTA (WARPS * 32) must call it.
//   tile  : shared float4[WARPS * kSplitTile]   slist : shared int[WARPS * nsample * 33]   scnt : shared int[WARPS * 32]
//   sidx  : shared int[nsample * STRIDE] — result, same contract as ball_scan_to_smem (slot fill, empty ball -> 0)
constexpr int kSplitTile = 128;

template <int WARPS, int STRIDE>
__device__ __forceinline__ void ball_scan_split_to_smem(int n, const float *__restrict__ xyz, bool active, float cx,
                                                        float cy, float cz, float r2, int nsample, float4 *tile,
                                                        int *slist, int *scnt, int *sidx) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int seg = ((n + WARPS - 1) / WARPS + 7) & ~7;
    const int beg = min(n, w * seg), end = min(n, beg + seg);
    float4 *mt = tile + w * kSplitTile;
    int *ml = slist + w * nsample * 33;
    int cnt = active ? 0 : nsample;
    for (int base = beg; base < end; base += kSplitTile) {
        if (__all_sync(0xffffffffu, cnt >= nsample)) break;
        const int len = min(kSplitTile, end - base);
        const int len8 = (len + 7) & ~7;
        __syncwarp();
        for (int i = lane; i < len8; i += 32) {
            if (i < len) {
                const float *q = xyz + (size_t)(base + i) * 3;
                mt[i] = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), 0.f);
            } else {
                mt[i] = make_float4(1e30f, 1e30f, 1e30f, 0.f);  // padding: squared distance overflows to +inf, never a hit
            }
        }
        __syncwarp();
        for (int i = 0; i < len8; i += 8) {
            float d2[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const float4 q = mt[i + u];
                d2[u] = sqdist3(cx, cy, cz, q.x, q.y, q.z);
            }
            const float mn = fminf(fminf(fminf(d2[0], d2[1]), fminf(d2[2], d2[3])),
                                   fminf(fminf(d2[4], d2[5]), fminf(d2[6], d2[7])));
            if (mn < r2 && cnt < nsample) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (d2[u] < r2 && cnt < nsample) {
                        ml[cnt * 33 + lane] = base + i + u;
                        cnt++;
                    }
                }
            }
        }
    }
    scnt[w * 32 + lane] = active ? cnt : 0;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int q = 0; q < WARPS; q++) {
        const int c = scnt[q * 32 + lane];
        before += q < w ? c : 0;
        total += c;
    }
    if (active)
        for (int l = 0; l < cnt && before + l < nsample; l++) sidx[(before + l) * STRIDE + lane] = ml[l * 33 + lane];
    __syncthreads();
    if (w == 0) {
        total = min(total, nsample);
        const int first = (active && total > 0) ? sidx[lane] : 0;  // empty ball: the pre-zeroed row groups point 0
        for (int l = active ? total : 0; l < nsample; l++) sidx[l * STRIDE + lane] = first;
    }
    __syncthreads();
}

// ---- hashed cell list -------------------------------------------------------------------------------------------
// A brute-force ordered scan costs M * N distance tests per scene whatever the radius.  With the points of a scene
// bucketed by cell (edge kCellSlack * radius, spatial hash into kGridBuckets buckets per scene) a centre only has to
// test the points of the 27 cells around it; the ball query's contract (the first nsample hits IN INDEX ORDER) is
// kept by inserting every hit into a bounded ascending list — the nsample smallest indices among all hits are exactly
// what the in-order scan collects.  Hash collisions only add candidates (every candidate is distance-tested with the
// reference's expression), a bucket visited twice only re-offers indices the list already holds or has already
// dropped, so the lists, and everything computed from them, are bit-identical to the scan's.
// Coverage: cell(x) = floor(fl(x * inv_edge)) is monotone in x and edge exceeds the radius by 0.1 %, far more than the
// rounding of the product for |x| < 2^20 cells, so two points closer than the radius are at most one cell apart per axis.
constexpr int kGridBuckets = 1 << 16;
constexpr float kCellSlack = 1.001f;

__device__ __forceinline__ int grid_cell(float x, float inv_edge) { return (int)floorf(x * inv_edge); }
__device__ __forceinline__ unsigned grid_hash(int ix, int iy, int iz) {
    return ((unsigned)ix * 73856093u ^ (unsigned)iy * 19349663u ^ (unsigned)iz * 83492791u) & (kGridBuckets - 1);
}

// Workspace per scene: start[kGridBuckets + 1] | cursor[kGridBuckets] ints, then float4 sorted[n] = (x, y, z, index bits).
__host__ __device__ inline size_t grid_scene_ints() { return 2 * (size_t)kGridBuckets + 4; }
__host__ __device__ inline size_t grid_scene_bytes(int n) { return grid_scene_ints() * sizeof(int) + (size_t)n * sizeof(float4); }

// One CTA per scene builds the scene's cell list: bucket counts (global atomics), exclusive scan, scatter of
// (x, y, z, index) into bucket order.  ~20 us for 16 x 16384 points; the order inside a bucket is arbitrary (the query
// keeps hits sorted by index).
constexpr int kBuildThreads = 1024;
static __global__ void __launch_bounds__(kBuildThreads)
grid_build_kernel(int n, float inv_edge, const float *__restrict__ xyz_all, unsigned char *__restrict__ ws) {
    constexpr int NB = kGridBuckets;
    __shared__ int warp_sums[kBuildThreads / 32];
    const int scene = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float *xyz = xyz_all + (size_t)scene * n * 3;
    unsigned char *wscene = ws + (size_t)scene * grid_scene_bytes(n);
    int *start = reinterpret_cast<int *>(wscene);
    int *cursor = start + NB + 4;
    float4 *sorted = reinterpret_cast<float4 *>(wscene + grid_scene_ints() * sizeof(int));
    for (int i = t; i < NB; i += kBuildThreads) cursor[i] = 0;
    __syncthreads();
    for (int k = t; k < n; k += kBuildThreads) {
        const float x = __ldg(xyz + (size_t)k * 3), y = __ldg(xyz + (size_t)k * 3 + 1), z = __ldg(xyz + (size_t)k * 3 + 2);
        atomicAdd(cursor + grid_hash(grid_cell(x, inv_edge), grid_cell(y, inv_edge),
                                           grid_cell(z, inv_edge)), 1);
    }
    __syncthreads();
    // exclusive scan of the NB counts.  Warp w owns the contiguous range [w * NB / 32, +NB / 32) and walks it 32 buckets at a time
    // (coalesced 128-byte accesses; PER consecutive buckets per THREAD, the first version, made every access of the three table
    // sweeps touch 32 different lines: 150 us per launch, issue slots 3.7 % busy): shuffle scan of the 32 counts + a running carry,
    // then the warps' totals are scanned and the offsets added in a second coalesced sweep.
    constexpr int WB = NB / (kBuildThreads / 32);           // buckets per warp
    int carry = 0;
    for (int i = 0; i < WB; i += 32) {
        const int b = warp * WB + i + lane;
        const int c = cursor[b];
        int incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        start[b] = carry + incl - c;                         // exclusive, relative to the warp's range
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) warp_sums[warp] = carry;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sums[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += v;
        }
        warp_sums[lane] = w;  // inclusive
    }
    __syncthreads();
    const int woff = warp > 0 ? warp_sums[warp - 1] : 0;
    for (int i = 0; i < WB; i += 32) {
        const int b = warp * WB + i + lane;
        const int v = start[b] + woff;
        start[b] = v;
        cursor[b] = v;
    }
    const int run = warp_sums[kBuildThreads / 32 - 1];
    if (t == kBuildThreads - 1) start[NB] = run;
    __syncthreads();
#pragma unroll 4   // four atomics with their return values in flight per thread
    for (int k = t; k < n; k += kBuildThreads) {
        const float x = __ldg(xyz + (size_t)k * 3), y = __ldg(xyz + (size_t)k * 3 + 1), z = __ldg(xyz + (size_t)k * 3 + 2);
        const int pos = atomicAdd(cursor + grid_hash(grid_cell(x, inv_edge), grid_cell(y, inv_edge),
                                                           grid_cell(z, inv_edge)), 1);
        sorted[pos] = make_float4(x, y, z, __int_as_float(k));
    }
}

// Insert `idx` into the ascending list column `t` (pitch STRIDE) holding `cnt` <= cap entries; keeps the cap smallest.
template <int STRIDE>
__device__ __forceinline__ void sorted_insert(int *list, int t, int &cnt, int cap, int idx) {
    int pos = cnt;
    while (pos > 0 && list[(pos - 1) * STRIDE + t] > idx) pos--;
    if (pos >= cap || (pos > 0 && list[(pos - 1) * STRIDE + t] == idx)) return;   // beyond the cap, or already there
    const int last = cnt < cap ? cnt : cap - 1;
    for (int j = last; j > pos; j--) list[j * STRIDE + t] = list[(j - 1) * STRIDE + t];
    list[pos * STRIDE + t] = idx;
    if (cnt < cap) cnt++;
}

// Two-radius query against the cell list: same results as ball_scan2_to_smem.  All threads of the CTA must call it.
template <int THREADS, int STRIDE>
__device__ __forceinline__ void grid_scan2_to_smem(const int *__restrict__ start, const float4 *__restrict__ sorted,
                                                   float inv_edge, bool active, float cx, float cy, float cz, float r2a,
                                                   int ns_a, float r2b, int ns_b, int *__restrict__ sidx_a,
                                                   int *__restrict__ sidx_b) {
    const int t = threadIdx.x;
    active = active && t < THREADS;
    int ca = 0, cb = 0;
    if (active) {
        const int ix = grid_cell(cx, inv_edge), iy = grid_cell(cy, inv_edge), iz = grid_cell(cz, inv_edge);
        for (int dz = -1; dz <= 1; dz++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    const unsigned bkt = grid_hash(ix + dx, iy + dy, iz + dz);
                    const int beg = __ldg(start + bkt), end = __ldg(start + bkt + 1);
                    for (int i = beg; i < end; i++) {
                        const float4 q = __ldg(sorted + i);
                        const float d2 = sqdist3(cx, cy, cz, q.x, q.y, q.z);
                        if (d2 < r2b) sorted_insert<STRIDE>(sidx_b, t, cb, ns_b, __float_as_int(q.w));
                        if (d2 < r2a) sorted_insert<STRIDE>(sidx_a, t, ca, ns_a, __float_as_int(q.w));
                    }
                }
        const int fa = ca > 0 ? sidx_a[t] : 0;  // empty ball: the pre-zeroed row groups point 0
        for (int l = ca; l < ns_a; l++) sidx_a[l * STRIDE + t] = fa;
        const int fb = cb > 0 ? sidx_b[t] : 0;
        for (int l = cb; l < ns_b; l++) sidx_b[l * STRIDE + t] = fb;
    } else if (t < THREADS) {
        for (int l = 0; l < ns_a; l++) sidx_a[l * STRIDE + t] = 0;
        for (int l = 0; l < ns_b; l++) sidx_b[l * STRIDE + t] = 0;
    }
    __syncthreads();
}

// Single-radius query against the cell list: same list as ball_scan_to_smem's in-order scan.  Returns the hit count (0 = empty
// ball; the caller decides what an empty ball means).  No barrier inside: every thread works on its own list column.
template <int STRIDE>
__device__ __forceinline__ int grid_scan_column(const int *__restrict__ start, const float4 *__restrict__ sorted, float inv_edge,
                                                float cx, float cy, float cz, float r2, int nsample, int *__restrict__ list,
                                                int t) {
    int cnt = 0;
    const int ix = grid_cell(cx, inv_edge), iy = grid_cell(cy, inv_edge), iz = grid_cell(cz, inv_edge);
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                const unsigned bkt = grid_hash(ix + dx, iy + dy, iz + dz);
                const int beg = __ldg(start + bkt), end = __ldg(start + bkt + 1);
                for (int i = beg; i < end; i++) {
                    const float4 q = __ldg(sorted + i);
                    if (sqdist3(cx, cy, cz, q.x, q.y, q.z) < r2) sorted_insert<STRIDE>(list, t, cnt, nsample, __float_as_int(q.w));
                }
            }
    return cnt;
}

}  // namespace pdab
