// TEST INFRASTRUCTURE — not product code.
//
// C-ABI binding over the *reference's own* CUDA launchers, so that tests and
// bench baselines can run the unmodified reference kernels (compiled from
// /root/reference by oracle/build_ref.py into oracle/_ref/) on raw device
// pointers.  The reference's host wrappers (PB/src/sampling.cpp,
// ball_query.cpp, group_points.cpp) only unwrap at::Tensor pointers and
// include the long-removed <THC/THC.h>, so they cannot be built against
// torch 2.11; this file replaces exactly that unwrapping and nothing else.
//
// Launcher prototypes are the reference's (declared, not copied, here):
//   PB/src/sampling_gpu.h:12-39, PB/src/ball_query_gpu.h:12-19,
//   PB/src/group_points_gpu.h:10-20   (PB = pcdet/ops/pointnet2/pointnet2_batch)
// The reference launches on the legacy default stream; every entry point
// below therefore ends with a device synchronize so callers on other streams
// observe completed results.
#include <cuda_runtime.h>

void farthest_point_sampling_kernel_launcher(int b, int n, int m, const float *dataset, float *temp, int *idxs);
void furthest_point_sampling_with_dist_kernel_launcher(int b, int n, int m, const float *dataset, float *temp, int *idxs);
void gather_points_kernel_launcher_fast(int b, int c, int n, int npoints, const float *points, const int *idx, float *out);
void gather_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, const float *grad_out, const int *idx, float *grad_points);
// NB: the reference header names these two parameters (xyz, new_xyz) but the
// definition and every caller pass (new_xyz, xyz) — PB/src/ball_query_gpu.cu:47-48,
// PB/src/ball_query.cpp:39-41.
void ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx);
void ball_query_dilated_kernel_launcher_fast(int b, int n, int m, float max_radius, float min_radius, int nsample, const float *new_xyz, const float *xyz, int *idx);
void group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out);
void group_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx, float *grad_points);

// PB/src/interpolate_gpu.h:17-27, pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:341 (no header in the reference)
void three_nn_kernel_launcher_fast(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx);
void three_interpolate_kernel_launcher_fast(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out);
void three_interpolate_grad_kernel_launcher_fast(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points);
void points_in_boxes_launcher(int batch_size, int boxes_num, int pts_num, const float *boxes, const float *pts, int *box_idx_of_points);

static int done() {
    cudaError_t e = cudaDeviceSynchronize();
    return e == cudaSuccess ? 0 : (int)e;
}

extern "C" {

int ref_fps(int b, int n, int m, const float *xyz, float *temp, int *idx) {
    cudaDeviceSynchronize();
    farthest_point_sampling_kernel_launcher(b, n, m, xyz, temp, idx);
    return done();
}
int ref_fps_with_dist(int b, int n, int m, const float *dist, float *temp, int *idx) {
    cudaDeviceSynchronize();
    furthest_point_sampling_with_dist_kernel_launcher(b, n, m, dist, temp, idx);
    return done();
}
int ref_gather(int b, int c, int n, int np, const float *points, const int *idx, float *out) {
    cudaDeviceSynchronize();
    gather_points_kernel_launcher_fast(b, c, n, np, points, idx, out);
    return done();
}
int ref_gather_grad(int b, int c, int n, int np, const float *grad_out, const int *idx, float *grad_points) {
    cudaDeviceSynchronize();
    gather_points_grad_kernel_launcher_fast(b, c, n, np, grad_out, idx, grad_points);
    return done();
}
int ref_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx) {
    cudaDeviceSynchronize();
    ball_query_kernel_launcher_fast(b, n, m, radius, nsample, new_xyz, xyz, idx);
    return done();
}
int ref_ball_query_dilated(int b, int n, int m, float max_radius, float min_radius, int nsample,
                           const float *new_xyz, const float *xyz, int *idx) {
    cudaDeviceSynchronize();
    ball_query_dilated_kernel_launcher_fast(b, n, m, max_radius, min_radius, nsample, new_xyz, xyz, idx);
    return done();
}
int ref_group(int b, int c, int n, int np, int ns, const float *points, const int *idx, float *out) {
    cudaDeviceSynchronize();
    group_points_kernel_launcher_fast(b, c, n, np, ns, points, idx, out);
    return done();
}
int ref_group_grad(int b, int c, int n, int np, int ns, const float *grad_out, const int *idx, float *grad_points) {
    cudaDeviceSynchronize();
    group_points_grad_kernel_launcher_fast(b, c, n, np, ns, grad_out, idx, grad_points);
    return done();
}

int ref_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx) {
    cudaDeviceSynchronize();
    three_nn_kernel_launcher_fast(b, n, m, unknown, known, dist2, idx);
    return done();
}
int ref_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out) {
    cudaDeviceSynchronize();
    three_interpolate_kernel_launcher_fast(b, c, m, n, points, idx, weight, out);
    return done();
}
int ref_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight,
                               float *grad_points) {
    cudaDeviceSynchronize();
    three_interpolate_grad_kernel_launcher_fast(b, c, n, m, grad_out, idx, weight, grad_points);
    return done();
}
int ref_points_in_boxes(int batch, int nboxes, int npts, const float *boxes, const float *pts, int *out) {
    cudaDeviceSynchronize();
    points_in_boxes_launcher(batch, nboxes, npts, boxes, pts, out);
    return done();
}

}  // extern "C"
