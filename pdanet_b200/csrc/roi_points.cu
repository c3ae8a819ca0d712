// points_in_boxes_gpu: index of the first box that contains each point, -1 otherwise.
// replaces: points_in_boxes_kernel / check_pt_in_box3d / lidar_to_local_coords,
//           pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:16-38,313-338 — the native op of PDA-SSD's
//           training-side target assignment (pcdet/models/dense_heads/IASSD_head.py:169,196,214).
//
// The reference evaluates cos / sin of the heading and three double-precision half extents for every (point, box)
// pair.  Here the boxes of a scene are prepared once per CTA into shared memory — centre, cosf(-rz), sinf(-rz) and the
// three thresholds as doubles — and every point scans them in order with an early exit; the per-pair work is two
// FMAs and three compares.  Arithmetic is the reference's as compiled for sm_100a (SASS of the rebuilt reference):
//   local_x = fma(sx, cosa, -rn(sy * sina)),  local_y = fma(sy, cosa, rn(sx * sina)),  sina = sinf(-rz), cosa = cosf(-rz)
//   inside  = !(|z - cz| > dz * 0.5)  &&  |local_x| < dx * 0.5 + 1e-5f  &&  |local_y| < dy * 0.5 + 1e-5f     (doubles)
// so the result is identical, including points on a face.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kBoxTile = 512;

struct BoxPrep {
    float cx, cy, cz, cosa, sina;
    float pad;
    double hx, hy, hz;
};

__global__ void __launch_bounds__(kThreads)
points_in_boxes_kernel(int boxes_num, int pts_num, const float *__restrict__ boxes, const float *__restrict__ pts,
                       int *__restrict__ box_idx_of_points) {
    __shared__ BoxPrep sbox[kBoxTile];
    const int b = blockIdx.y;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    boxes += (size_t)b * boxes_num * 7;
    float x = 0.f, y = 0.f, z = 0.f;
    const bool active = i < pts_num;
    if (active) {
        const float *p = pts + ((size_t)b * pts_num + i) * 3;
        x = p[0];
        y = p[1];
        z = p[2];
    }
    int found = -1;
    for (int base = 0; base < boxes_num; base += kBoxTile) {
        const int len = min(kBoxTile, boxes_num - base);
        __syncthreads();
        for (int k = threadIdx.x; k < len; k += kThreads) {
            const float *bx = boxes + (size_t)(base + k) * 7;
            BoxPrep q;
            q.cx = bx[0];
            q.cy = bx[1];
            q.cz = bx[2];
            const float rz = bx[6];
            q.cosa = cosf(-rz);
            q.sina = sinf(-rz);
            q.pad = 0.f;
            q.hx = fma((double)bx[3], 0.5, (double)1e-5f);
            q.hy = fma((double)bx[4], 0.5, (double)1e-5f);
            q.hz = (double)bx[5] * 0.5;
            sbox[k] = q;
        }
        __syncthreads();
        if (active && found < 0) {
            for (int k = 0; k < len; k++) {
                const BoxPrep &q = sbox[k];
                if ((double)fabsf(__fsub_rn(z, q.cz)) > q.hz) continue;
                const float sx = __fsub_rn(x, q.cx), sy = __fsub_rn(y, q.cy);
                const float lx = __fmaf_rn(sx, q.cosa, -__fmul_rn(sy, q.sina));
                const float ly = __fmaf_rn(sy, q.cosa, __fmul_rn(sx, q.sina));
                if ((double)fabsf(lx) < q.hx && (double)fabsf(ly) < q.hy) {
                    found = base + k;
                    break;
                }
            }
        }
    }
    if (active && found >= 0) box_idx_of_points[(size_t)b * pts_num + i] = found;  // caller pre-fills -1
}

}  // namespace

extern "C" int pdab_points_in_boxes(int batch_size, int boxes_num, int pts_num, const float *boxes, const float *pts,
                                    int *box_idx_of_points, pdab_stream_t stream) {
    if (batch_size < 0 || boxes_num < 0 || pts_num < 0 || !pts || !box_idx_of_points || (boxes_num > 0 && !boxes))
        return PDAB_EINVAL;
    if (batch_size == 0 || pts_num == 0 || boxes_num == 0) return 0;
    if (batch_size > 65535) return PDAB_EUNSUPPORTED;
    dim3 grid(pdab::div_up(pts_num, kThreads), batch_size);
    points_in_boxes_kernel<<<grid, kThreads, 0, pdab::to_stream(stream)>>>(boxes_num, pts_num, boxes, pts,
                                                                           box_idx_of_points);
    PDAB_LAUNCH_CHECK();
    return 0;
}
