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
#include "common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kTile = 1024;  // points per shared-memory tile (16 KB)

template <bool DILATED>
__global__ void __launch_bounds__(kThreads)
ball_query_kernel(int n, int m, float r2_hi, float r2_lo, int nsample, const float *__restrict__ new_xyz,
                  const float *__restrict__ xyz, int *__restrict__ idx) {
    __shared__ float4 tile[kTile];
    const int scene = blockIdx.y;
    const int j = blockIdx.x * kThreads + threadIdx.x;
    xyz += (size_t)scene * n * 3;
    const bool active = j < m;
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
    ball_query_kernel<DILATED><<<grid, kThreads, 0, stream>>>(n, m, r_hi * r_hi, r_lo * r_lo, nsample, new_xyz, xyz, idx);
    PDAB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int pdab_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                               int *idx, pdab_stream_t stream) {
    return launch<false>(b, n, m, radius, 0.f, nsample, new_xyz, xyz, idx, pdab::to_stream(stream));
}

extern "C" int pdab_ball_query_dilated(int b, int n, int m, float max_radius, float min_radius, int nsample,
                                       const float *new_xyz, const float *xyz, int *idx, pdab_stream_t stream) {
    return launch<true>(b, n, m, max_radius, min_radius, nsample, new_xyz, xyz, idx, pdab::to_stream(stream));
}
