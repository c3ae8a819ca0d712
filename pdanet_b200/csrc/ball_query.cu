// Ball query for sm_100a.
//
// One thread per query centre; the scene's points stream through shared memory
// in float4 tiles (one coalesced global read per CTA instead of one uncached
// scan per thread as in PB/src/ball_query_gpu.cu:23-44), every lane reads the
// same point by broadcast (LDS.128), hits are written in index order.  A warp
// stops scanning once all of its centres are full; a CTA stops staging tiles
// once all of its warps have.
//
// Exactness: the fp32 test is sqdist3(centre, point) < radius*radius with the
// reference's compiled op order (common.cuh); hit order = point index order;
// unfilled slots repeat the first hit; an empty ball leaves its row untouched.
#include "ball_scan.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kTile = 1024;  // points per shared-memory tile (16 KB)

template <bool DILATED>
__global__ void __launch_bounds__(kThreads)
ball_query_kernel(int n, int m, float r2_hi, float r2_lo, int nsample, const float *__restrict__ new_xyz,
                  const float *__restrict__ xyz, int *__restrict__ idx, const int *__restrict__ todo) {
    __shared__ float4 tile[kTile];
    const int scene = blockIdx.y;
    const int j = blockIdx.x * kThreads + threadIdx.x;
    xyz += (size_t)scene * n * 3;
    // todo (optional, (B, M)): only the flagged centres are searched (the dense balls the cell-list kernel handed over)
    const bool active = j < m && (todo == nullptr || todo[(size_t)scene * m + j] != 0);
    if (todo != nullptr && !__syncthreads_or(active)) return;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    int *row = nullptr;
    if (active) {
        const float *c = new_xyz + ((size_t)scene * m + j) * 3;
        cx = c[0];
        cy = c[1];
        cz = c[2];
        row = idx + ((size_t)scene * m + j) * nsample;
    }
    int cnt = active ? 0 : nsample;
    int first = -1;

    for (int base = 0; base < n; base += kTile) {
        const int len = min(kTile, n - base);
        __syncthreads();  // previous tile fully consumed
        const int len8 = (len + 7) & ~7;
        for (int i = threadIdx.x; i < len8; i += kThreads) {
            if (i < len) {
                const float *p = xyz + (size_t)(base + i) * 3;
                tile[i] = make_float4(p[0], p[1], p[2], 0.f);
            } else {
                tile[i] = make_float4(1e30f, 1e30f, 1e30f, 0.f);  // padding: distance overflows to +inf, never a hit
            }
        }
        __syncthreads();
        if (!__all_sync(0xffffffffu, cnt >= nsample)) {
            // eight independent distance tests in flight, one branch for the (rare) hit
            for (int i0 = 0; i0 < len8; i0 += 8) {
                float d2v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const float4 p = tile[i0 + u];
                    d2v[u] = pdab::sqdist3(cx, cy, cz, p.x, p.y, p.z);
                }
                const float mn = fminf(fminf(fminf(d2v[0], d2v[1]), fminf(d2v[2], d2v[3])),
                                       fminf(fminf(d2v[4], d2v[5]), fminf(d2v[6], d2v[7])));
                if ((DILATED ? !(mn <= r2_hi) : !(mn < r2_hi)) || cnt >= nsample) continue;  // dilated also emits d2 == 0
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const float d2 = d2v[u];
                    const int i = i0 + u;
                    if (DILATED) {
                        // PB/src/ball_query_gpu.cu:92-111: two independent tests; a point
                        // can be emitted by both.
                        if (d2 == 0.f && cnt < nsample) {
                            if (cnt == 0) first = base + i;
                            row[cnt++] = base + i;
                        }
                        if (d2 >= r2_lo && d2 < r2_hi && cnt < nsample) {
                            if (cnt == 0) first = base + i;
                            row[cnt++] = base + i;
                        }
                    } else {
                        if (d2 < r2_hi && cnt < nsample) {
                            if (cnt == 0) first = base + i;
                            row[cnt++] = base + i;
                        }
                    }
                }
            }
        }
        if (__syncthreads_and(cnt >= nsample)) break;
    }
    if (active && first >= 0)
        for (int l = cnt; l < nsample; l++) row[l] = first;
}

template <bool DILATED>
int launch(int b, int n, int m, float r_hi, float r_lo, int nsample, const float *new_xyz, const float *xyz, int *idx,
           cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0 || nsample < 1 || !new_xyz || !xyz || !idx) return PDAB_EINVAL;
    if (b == 0 || m == 0 || n == 0) return 0;
    if (b > 65535) return PDAB_EUNSUPPORTED;
    dim3 grid(pdab::div_up(m, kThreads), b);
    // radius*radius in fp32, as PB/src/ball_query_gpu.cu:23,85-86
    ball_query_kernel<DILATED><<<grid, kThreads, 0, stream>>>(n, m, r_hi * r_hi, r_lo * r_lo, nsample, new_xyz, xyz, idx, nullptr);
    PDAB_LAUNCH_CHECK();
    return 0;
}

// Same result from the scene's hashed cell list (ball_scan.cuh): a centre tests the points of the 27 cells around it and keeps
// the nsample smallest indices among the hits = the first nsample hits of the in-order scan.  Thread per centre; the lists
// live in shared memory (slot-major, one column per thread) and leave through one coalesced sweep.
// Dense balls (large radius, squeezed cloud) are cheaper by the in-order scan, which stops at nsample hits: scanning costs about
// N * nsample / hits tests, the cell list cand * (1 + insertion).  A centre whose 27 cells hold more than 2 nsample + 32 points
// counts its hits first; with at least 2 nsample of them (the scan then stops after about half the cloud at most) it is flagged
// instead of searched here, and the tiled scan kernel then runs for the flagged centres only (CTAs without one return at once).
constexpr int kGridStride = kThreads + 1;
__global__ void __launch_bounds__(kThreads)
ball_query_grid_kernel(int n, int m, float r2, float inv_edge, int nsample, int dense_above, const float *__restrict__ new_xyz,
                       const unsigned char *__restrict__ ws, int *__restrict__ todo, int *__restrict__ idx) {
    extern __shared__ int slist[];                       // nsample x kGridStride
    const int scene = blockIdx.y, t = threadIdx.x;
    const int j0 = blockIdx.x * kThreads, j = j0 + t;
    const unsigned char *wscene = ws + (size_t)scene * pdab::grid_scene_bytes(n);
    const int *start = reinterpret_cast<const int *>(wscene);
    const float4 *sorted = reinterpret_cast<const float4 *>(wscene + pdab::grid_scene_ints() * sizeof(int));
    const bool active = j < m;
    bool dense = false;
    int cnt = 0;
    if (active) {
        const float *c = new_xyz + ((size_t)scene * m + j) * 3;
        const float cx = c[0], cy = c[1], cz = c[2];
        const int ix = pdab::grid_cell(cx, inv_edge), iy = pdab::grid_cell(cy, inv_edge), iz = pdab::grid_cell(cz, inv_edge);
        int cand = 0;
        for (int dz = -1; dz <= 1; dz++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    const unsigned bkt = pdab::grid_hash(ix + dx, iy + dy, iz + dz);
                    cand += __ldg(start + bkt + 1) - __ldg(start + bkt);
                }
        if (cand > dense_above) {   // many candidates: count the hits first — the scan only pays if it can stop early
            int hits = 0;
            for (int dz = -1; dz <= 1; dz++)
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        const unsigned bkt = pdab::grid_hash(ix + dx, iy + dy, iz + dz);
                        const int beg = __ldg(start + bkt), end = __ldg(start + bkt + 1);
                        for (int i = beg; i < end; i++) {
                            const float4 q = __ldg(sorted + i);
                            hits += pdab::sqdist3(cx, cy, cz, q.x, q.y, q.z) < r2 ? 1 : 0;
                        }
                    }
            dense = hits >= 2 * nsample;   // (a bucket reached through a hash collision is counted twice: only a heuristic)
        }
        todo[(size_t)scene * m + j] = dense ? 1 : 0;
        if (!dense) {
            cnt = pdab::grid_scan_column<kGridStride>(start, sorted, inv_edge, cx, cy, cz, r2, nsample, slist, t);
            const int first = cnt > 0 ? slist[t] : 0;
            for (int l = cnt; l < nsample; l++) slist[l * kGridStride + t] = first;
        }
        // rows this kernel does not write: dense ones (the scan kernel's), empty balls (stay zero, PB/pointnet2_utils.py:246)
        if (dense || cnt == 0) slist[t] = -1;
    }
    __syncthreads();
    const int nctr = min(kThreads, m - j0);
    int *out = idx + ((size_t)scene * m + j0) * nsample;
    for (int e = t; e < nctr * nsample; e += kThreads) {
        const int jl = e / nsample, l = e - jl * nsample;
        if (slist[jl] >= 0) out[e] = slist[l * kGridStride + jl];
    }
}

}  // namespace

extern "C" size_t pdab_ball_query_grid_workspace_bytes(int b, int n, int m) {
    if (b < 1 || n < 1 || m < 1) return 0;
    return (size_t)b * pdab::grid_scene_bytes(n) + (((size_t)b * m * sizeof(int) + 15) & ~(size_t)15);
}

extern "C" int pdab_ball_query_grid(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                                    int *idx, void *workspace, pdab_stream_t stream) {
    if (b < 0 || n < 0 || m < 0 || nsample < 1 || !new_xyz || !xyz || !idx || !workspace || !(radius > 0.f)) return PDAB_EINVAL;
    if (b == 0 || m == 0 || n == 0) return 0;
    if (b > 65535 || nsample > 256) return PDAB_EUNSUPPORTED;
    cudaStream_t s = pdab::to_stream(stream);
    const float inv_edge = 1.0f / (pdab::kCellSlack * radius);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    int *todo = reinterpret_cast<int *>(ws + (size_t)b * pdab::grid_scene_bytes(n));
    pdab::grid_build_kernel<<<b, pdab::kBuildThreads, 0, s>>>(n, inv_edge, xyz, ws);
    PDAB_LAUNCH_CHECK();
    const size_t smem = sizeof(int) * (size_t)nsample * kGridStride;
    // per device and cheap: set on every launch (a cached flag would skip the second GPU of a process)
    PDAB_CUDA(cudaFuncSetAttribute(ball_query_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int dense_above = 2 * nsample + 32;
    dim3 grid(pdab::div_up(m, kThreads), b);
    ball_query_grid_kernel<<<grid, kThreads, smem, s>>>(n, m, radius * radius, inv_edge, nsample, dense_above, new_xyz, ws, todo,
                                                        idx);
    PDAB_LAUNCH_CHECK();
    ball_query_kernel<false><<<grid, kThreads, 0, s>>>(n, m, radius * radius, 0.f, nsample, new_xyz, xyz, idx, todo);
    PDAB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pdab_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                               int *idx, pdab_stream_t stream) {
    return launch<false>(b, n, m, radius, 0.f, nsample, new_xyz, xyz, idx, pdab::to_stream(stream));
}

extern "C" int pdab_ball_query_dilated(int b, int n, int m, float max_radius, float min_radius, int nsample,
                                       const float *new_xyz, const float *xyz, int *idx, pdab_stream_t stream) {
    return launch<true>(b, n, m, max_radius, min_radius, nsample, new_xyz, xyz, idx, pdab::to_stream(stream));
}
