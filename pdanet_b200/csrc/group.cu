// gather_points / group_points (+ their gradients) for sm_100a.
//
// The reference launches one thread per OUTPUT ELEMENT PER CHANNEL and re-reads
// the index for every channel (PB/src/group_points_gpu.cu:59-71,
// PB/src/sampling_gpu.cu:15-23).  Here a thread owns one (centre, sample) slot,
// reads its index once and walks a strip of channels: index traffic drops by
// the strip length, stores stay fully coalesced along (centre, sample), and the
// 4-byte gathers of one channel row hit L1/L2 (a row is 4*N bytes).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kStrip = 8;  // channels per thread

// out[b, c, e] = points[b, c, idx[b, e]],  e in [0, E)   (E = npoints or npoints*nsample)
__global__ void __launch_bounds__(kThreads)
index_gather_kernel(int c, int n, int E, const float *__restrict__ points, const int *__restrict__ idx,
                    float *__restrict__ out) {
    const int scene = blockIdx.z;
    const int e = blockIdx.x * kThreads + threadIdx.x;
    if (e >= E) return;
    const int c0 = blockIdx.y * kStrip;
    const int k = __ldg(idx + (size_t)scene * E + e);
    const float *src = points + ((size_t)scene * c + c0) * n + k;
    float *dst = out + ((size_t)scene * c + c0) * E + e;
    const int cn = min(kStrip, c - c0);
    float v[kStrip];
#pragma unroll
    for (int q = 0; q < kStrip; q++)
        if (q < cn) v[q] = __ldg(src + (size_t)q * n);
#pragma unroll
    for (int q = 0; q < kStrip; q++)
        if (q < cn) __stcs(dst + (size_t)q * E, v[q]);
}

// grad_points[b, c, idx[b, e]] += grad_out[b, c, e]
__global__ void __launch_bounds__(kThreads)
index_scatter_add_kernel(int c, int n, int E, const float *__restrict__ grad_out, const int *__restrict__ idx,
                         float *__restrict__ grad_points) {
    const int scene = blockIdx.z;
    const int e = blockIdx.x * kThreads + threadIdx.x;
    if (e >= E) return;
    const int c0 = blockIdx.y * kStrip;
    const int k = __ldg(idx + (size_t)scene * E + e);
    const float *src = grad_out + ((size_t)scene * c + c0) * E + e;
    float *dst = grad_points + ((size_t)scene * c + c0) * n + k;
    const int cn = min(kStrip, c - c0);
#pragma unroll
    for (int q = 0; q < kStrip; q++)
        if (q < cn) atomicAdd(dst + (size_t)q * n, __ldg(src + (size_t)q * E));
}

// Deterministic gradient of a gather / group / interpolate: the reference scatters with float atomicAdd
// (PB/src/group_points_gpu.cu:30, sampling_gpu.cu:62, interpolate_gpu.cu:139), whose summation order changes from run to
// run.  Here the (scene, position) slots are sorted by the point they read (a stable sort made by the caller), so every
// source point owns one contiguous segment and a thread adds its segment up in a fixed order: bit-identical gradients
// on every run, no atomics.  grad_points[b, c, i] = sum over k in [seg[b, i], seg[b, i + 1]) of
// grad_out[b, c, order[b, k] / div] * (weight ? weight[b, order[b, k]] : 1).
__global__ void __launch_bounds__(kThreads)
segment_sum_kernel(int c, int n, int E, int div, const float *__restrict__ grad_out, const float *__restrict__ weight,
                   const int *__restrict__ order, const int *__restrict__ seg, float *__restrict__ grad_points) {
    const int scene = blockIdx.z;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const int c0 = blockIdx.y * kStrip;
    const int cn = min(kStrip, c - c0);
    const int Eo = E / div;
    const int k0 = __ldg(seg + (size_t)scene * (n + 1) + i), k1 = __ldg(seg + (size_t)scene * (n + 1) + i + 1);
    const float *src = grad_out + ((size_t)scene * c + c0) * Eo;
    float acc[kStrip];
#pragma unroll
    for (int q = 0; q < kStrip; q++) acc[q] = 0.f;
    for (int k = k0; k < k1; k++) {
        const int e = __ldg(order + (size_t)scene * E + k);
        const float w = weight ? __ldg(weight + (size_t)scene * E + e) : 1.f;
        const int col = e / div;
#pragma unroll
        for (int q = 0; q < kStrip; q++)
            if (q < cn) acc[q] = fmaf(__ldg(src + (size_t)q * Eo + col), w, acc[q]);
    }
    float *dst = grad_points + ((size_t)scene * c + c0) * n + i;
#pragma unroll
    for (int q = 0; q < kStrip; q++)
        if (q < cn) dst[(size_t)q * n] = acc[q];
}

int check(int b, int c, int n, int E, const void *a, const void *i, const void *o) {
    if (b < 0 || c < 0 || n < 0 || E < 0 || !a || !i || !o) return PDAB_EINVAL;
    if (b > 65535 || pdab::div_up(c, kStrip) > 65535) return PDAB_EUNSUPPORTED;
    return 0;
}

int gather_like(int b, int c, int n, int E, const float *points, const int *idx, float *out, cudaStream_t s) {
    if (int rc = check(b, c, n, E, points, idx, out)) return rc;
    if (b == 0 || c == 0 || E == 0) return 0;
    dim3 grid(pdab::div_up(E, kThreads), pdab::div_up(c, kStrip), b);
    index_gather_kernel<<<grid, kThreads, 0, s>>>(c, n, E, points, idx, out);
    PDAB_LAUNCH_CHECK();
    return 0;
}

int scatter_like(int b, int c, int n, int E, const float *grad_out, const int *idx, float *grad_points,
                 cudaStream_t s) {
    if (int rc = check(b, c, n, E, grad_out, idx, grad_points)) return rc;
    if (b == 0 || c == 0 || E == 0) return 0;
    dim3 grid(pdab::div_up(E, kThreads), pdab::div_up(c, kStrip), b);
    index_scatter_add_kernel<<<grid, kThreads, 0, s>>>(c, n, E, grad_out, idx, grad_points);
    PDAB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int pdab_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out,
                                  pdab_stream_t stream) {
    return gather_like(b, c, n, npoints, points, idx, out, pdab::to_stream(stream));
}

extern "C" int pdab_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                                       float *grad_points, pdab_stream_t stream) {
    return scatter_like(b, c, n, npoints, grad_out, idx, grad_points, pdab::to_stream(stream));
}

extern "C" int pdab_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx,
                                 float *out, pdab_stream_t stream) {
    return gather_like(b, c, n, npoints * nsample, points, idx, out, pdab::to_stream(stream));
}

extern "C" int pdab_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                      const int *idx, float *grad_points, pdab_stream_t stream) {
    return scatter_like(b, c, n, npoints * nsample, grad_out, idx, grad_points, pdab::to_stream(stream));
}

extern "C" int pdab_segment_sum_grad(int b, int c, int n, int e, int div, const float *grad_out, const float *weight,
                                     const int *order, const int *seg_start, float *grad_points, pdab_stream_t stream) {
    if (b < 0 || c < 0 || n < 0 || e < 0 || div < 1 || (e % div) || !grad_out || !order || !seg_start || !grad_points)
        return PDAB_EINVAL;
    if (b > 65535 || pdab::div_up(c, kStrip) > 65535) return PDAB_EUNSUPPORTED;
    if (b == 0 || c == 0 || n == 0) return 0;
    dim3 grid(pdab::div_up(n, kThreads), pdab::div_up(c, kStrip), b);
    segment_sum_kernel<<<grid, kThreads, 0, pdab::to_stream(stream)>>>(c, n, e, div, grad_out, weight, order, seg_start,
                                                                         grad_points);
    PDAB_LAUNCH_CHECK();
    return 0;
}
