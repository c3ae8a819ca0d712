// Rotated-BEV IoU and bitmask NMS for sm_100a, entirely on the device.
//
// The reference does, PER SCENE: cudaMalloc -> 64x64 mask tiles for the whole
// n x n square -> blocking D2H copy -> cudaFree -> greedy scan on the host
// (IOU/src/iou3d_nms.cpp:90-136).  Here one launch builds only the tiles that
// the greedy scan reads (diagonal and right of it) for every scene of a batch,
// and a second launch runs the greedy scan on the device (tile-serial inside,
// word-parallel across the row), so the hot path has no allocation, no
// host round trip and no host loop.
//
// Exactness: keep lists must equal the reference's, which hinges on
// iou > thresh for borderline pairs.  The geometry below evaluates the same
// real-valued formulas in the same fp32 expression shapes as
// IOU/src/iou3d_nms_kernel.cu:35-234 (edge-intersection + contained-corner
// polygon, centroid-angle ordering, fan area), uses the same libdevice
// cosf/sinf/atan2f, and is compiled with the same contraction setting; the GPU
// tests compare IoU matrices bit-for-bit with the reference kernel.  The
// polygon ordering uses a stable insertion sort on precomputed angles, which
// yields the identical permutation as the reference's bubble sort with a
// strict '>' comparator that recomputes atan2 per comparison.
#include <cstring>

#include "common.cuh"

namespace {

constexpr int kTileBoxes = 64;  // one mask word

struct P2 {
    float x, y;
};

__device__ __forceinline__ float cross_o(const P2 &p1, const P2 &p2, const P2 &p0) {
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

__device__ __forceinline__ bool boxes_may_touch(const P2 &p1, const P2 &p2, const P2 &q1, const P2 &q2) {
    return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
           fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}

// point inside the rotated rectangle `box` (margin 1e-2), IOU/src/iou3d_nms_kernel.cu:51-62
__device__ __forceinline__ bool inside_box(const float *box, const P2 &p) {
    const float MARGIN = 1e-2f;
    const float center_x = box[0], center_y = box[1];
    const float angle_cos = cosf(-box[6]), angle_sin = sinf(-box[6]);
    const float rot_x = (p.x - center_x) * angle_cos + (p.y - center_y) * (-angle_sin);
    const float rot_y = (p.x - center_x) * angle_sin + (p.y - center_y) * angle_cos;
    return fabsf(rot_x) < box[3] / 2 + MARGIN && fabsf(rot_y) < box[4] / 2 + MARGIN;
}

// proper intersection of segments p0p1 and q0q1, IOU/src/iou3d_nms_kernel.cu:64-93
__device__ __forceinline__ bool segment_cross(const P2 &p1, const P2 &p0, const P2 &q1, const P2 &q0, P2 &ans) {
    const float EPS = 1e-8f;
    if (!boxes_may_touch(p0, p1, q0, q1)) return false;
    const float s1 = cross_o(q0, p1, p0);
    const float s2 = cross_o(p1, q1, p0);
    const float s3 = cross_o(p0, q1, q0);
    const float s4 = cross_o(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
    const float s5 = cross_o(q1, p1, p0);
    if (fabsf(s5 - s1) > EPS) {
        ans.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float D = a0 * b1 - a1 * b0;
        ans.x = (b0 * c1 - b1 * c0) / D;
        ans.y = (a1 * c0 - a0 * c1) / D;
    }
    return true;
}

__device__ __forceinline__ void spin(const P2 &center, const float angle_cos, const float angle_sin, P2 &p) {
    const float new_x = (p.x - center.x) * angle_cos + (p.y - center.y) * (-angle_sin) + center.x;
    const float new_y = (p.x - center.x) * angle_sin + (p.y - center.y) * angle_cos + center.y;
    p.x = new_x;
    p.y = new_y;
}

__device__ __noinline__ float overlap_area(const float *box_a, const float *box_b) {
    const float a_angle = box_a[6], b_angle = box_b[6];
    const float a_dx_half = box_a[3] / 2, b_dx_half = box_b[3] / 2, a_dy_half = box_a[4] / 2, b_dy_half = box_b[4] / 2;
    const float a_x1 = box_a[0] - a_dx_half, a_y1 = box_a[1] - a_dy_half;
    const float a_x2 = box_a[0] + a_dx_half, a_y2 = box_a[1] + a_dy_half;
    const float b_x1 = box_b[0] - b_dx_half, b_y1 = box_b[1] - b_dy_half;
    const float b_x2 = box_b[0] + b_dx_half, b_y2 = box_b[1] + b_dy_half;
    const P2 center_a = {box_a[0], box_a[1]}, center_b = {box_b[0], box_b[1]};

    P2 ca[5] = {{a_x1, a_y1}, {a_x2, a_y1}, {a_x2, a_y2}, {a_x1, a_y2}, {0.f, 0.f}};
    P2 cb[5] = {{b_x1, b_y1}, {b_x2, b_y1}, {b_x2, b_y2}, {b_x1, b_y2}, {0.f, 0.f}};
    const float a_angle_cos = cosf(a_angle), a_angle_sin = sinf(a_angle);
    const float b_angle_cos = cosf(b_angle), b_angle_sin = sinf(b_angle);
    for (int k = 0; k < 4; k++) {
        spin(center_a, a_angle_cos, a_angle_sin, ca[k]);
        spin(center_b, b_angle_cos, b_angle_sin, cb[k]);
    }
    ca[4] = ca[0];
    cb[4] = cb[0];

    P2 poly[16];
    P2 centre = {0.f, 0.f};
    int cnt = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            if (segment_cross(ca[i + 1], ca[i], cb[j + 1], cb[j], poly[cnt])) {
                centre.x = centre.x + poly[cnt].x;
                centre.y = centre.y + poly[cnt].y;
                cnt++;
            }
    for (int k = 0; k < 4; k++) {
        if (inside_box(box_a, cb[k])) {
            centre.x = centre.x + cb[k].x;
            centre.y = centre.y + cb[k].y;
            poly[cnt++] = cb[k];
        }
        if (inside_box(box_b, ca[k])) {
            centre.x = centre.x + ca[k].x;
            centre.y = centre.y + ca[k].y;
            poly[cnt++] = ca[k];
        }
    }
    if (cnt < 3) return 0.f;  // no fan triangles: the reference's area loop adds only zero-area terms
    centre.x /= cnt;
    centre.y /= cnt;

    float ang[16];
    for (int i = 0; i < cnt; i++) ang[i] = atan2f(poly[i].y - centre.y, poly[i].x - centre.x);
    for (int i = 1; i < cnt; i++) {  // stable ascending insertion sort
        const P2 p = poly[i];
        const float a = ang[i];
        int j = i - 1;
        while (j >= 0 && ang[j] > a) {
            poly[j + 1] = poly[j];
            ang[j + 1] = ang[j];
            j--;
        }
        poly[j + 1] = p;
        ang[j + 1] = a;
    }

    float area = 0.f;
    for (int k = 0; k < cnt - 1; k++) {
        const P2 u = {poly[k].x - poly[0].x, poly[k].y - poly[0].y};
        const P2 v = {poly[k + 1].x - poly[0].x, poly[k + 1].y - poly[0].y};
        area += u.x * v.y - u.y * v.x;
    }
    return fabsf(area) / 2.0f;
}

// Far-apart boxes (almost every pair of a scene): when the axis-aligned squares around the circumscribed circles are more than
// 5 cm apart, no two edges' bounding boxes touch (boxes_may_touch is an exact comparison: every segment_cross returns false) and
// every corner is farther from the other centre than the half diagonal plus inside_box's 1e-2 margin, so the reference finds
// no polygon point and its overlap is +0 (IOU/src/iou3d_nms_kernel.cu:95-165, cnt = 0).  ~15 instructions instead of ~300
// (16 rejected edge pairs + 8 corner tests with their sincos).  NaN / inf boxes fail the comparison and take the full path.
// Kept OUT of overlap_area: that function's expressions match the reference's contraction choices bit for bit as compiled,
// and extra code inside it moved them (1-ulp IoU differences on overlapping pairs).
__device__ __forceinline__ bool far_apart(const float *box_a, const float *box_b) {
    const float ra = 0.5f * sqrtf(box_a[3] * box_a[3] + box_a[4] * box_a[4]);
    const float rb = 0.5f * sqrtf(box_b[3] * box_b[3] + box_b[4] * box_b[4]);
    const float reach = (ra + rb) * 1.0001f + 0.05f;
    return fabsf(box_a[0] - box_b[0]) > reach || fabsf(box_a[1] - box_b[1]) > reach;
}

__device__ __forceinline__ float iou_rotated(const float *box_a, const float *box_b) {
    const float sa = box_a[3] * box_a[4];
    const float sb = box_b[3] * box_b[4];
    const float s_overlap = far_apart(box_a, box_b) ? 0.f : overlap_area(box_a, box_b);
    return s_overlap / fmaxf(sa + sb - s_overlap, 1e-8f);
}

// IOU/src/iou3d_nms_kernel.cu:313-325
__device__ __forceinline__ float iou_axis_aligned(const float *a, const float *b) {
    const float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    const float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    const float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
    const float interS = width * height;
    const float Sa = a[3] * a[4];
    const float Sb = b[3] * b[4];
    return interS / fmaxf(Sa + Sb - interS, 1e-8f);
}

// ---- mask: grid (col tile, row tile, scene), only col >= row does work -------------------------
// mask layout per scene: (stride rows) x (col_blocks words), col_blocks = ceil(stride/64).
// A CTA = 64 rows x kSplit column slices (256 threads): thread (row, slice) tests its row against 64 / kSplit columns, the
// slices' bit fields meet in shared memory.  (One thread per row — the reference's shape — left 2 warps per SM doing 64
// rotated-IoU evaluations back to back: the n = 1024 mask took longer than the reference's whole call.)
constexpr int kSplit = 4;
constexpr int kMaskThreads = kTileBoxes * kSplit;

__device__ __forceinline__ int scene_count(const int *counts, int scene, int fixed_n, int stride) {
    // a count outside [0, stride] would walk into the next scene's slab: clamp
    return counts ? min(max(counts[scene], 0), stride) : fixed_n;
}

template <bool NORMAL>
__global__ void __launch_bounds__(kMaskThreads)
nms_mask_kernel(const float *__restrict__ boxes, const int *__restrict__ counts, int fixed_n, int stride,
                float thresh, unsigned long long *__restrict__ mask) {
    const int col = blockIdx.x, rowt = blockIdx.y, scene = blockIdx.z;
    if (col < rowt) return;  // never read by the greedy scan (IOU/src/iou3d_nms.cpp:128)
    const int n = scene_count(counts, scene, fixed_n, stride);
    if (rowt * kTileBoxes >= n || col * kTileBoxes >= n) return;
    boxes += (size_t)scene * stride * 7;
    const int col_blocks = pdab::div_up(stride, kTileBoxes);
    mask += (size_t)scene * stride * col_blocks;

    const int row_size = min(n - rowt * kTileBoxes, kTileBoxes);
    const int col_size = min(n - col * kTileBoxes, kTileBoxes);
    __shared__ float tile[kTileBoxes * 7];
    __shared__ unsigned long long part[kSplit][kTileBoxes];
    for (int i = threadIdx.x; i < col_size * 7; i += kMaskThreads) tile[i] = boxes[(size_t)col * kTileBoxes * 7 + i];
    __syncthreads();
    const int r = threadIdx.x & (kTileBoxes - 1), slice = threadIdx.x / kTileBoxes;
    unsigned long long bits = 0ull;
    if (r < row_size) {
        const int cur = rowt * kTileBoxes + r;
        float me[7];
#pragma unroll
        for (int q = 0; q < 7; q++) me[q] = boxes[(size_t)cur * 7 + q];
        constexpr int W = kTileBoxes / kSplit;
        const int lo = max(slice * W, (rowt == col) ? r + 1 : 0), hi = min((slice + 1) * W, col_size);
        for (int i = lo; i < hi; i++) {
            const float v = NORMAL ? iou_axis_aligned(me, tile + i * 7) : iou_rotated(me, tile + i * 7);
            if (v > thresh) bits |= 1ull << i;
        }
    }
    part[slice][r] = bits;
    __syncthreads();
    if (slice == 0 && r < row_size) {
#pragma unroll
        for (int q = 1; q < kSplit; q++) bits |= part[q][r];
        mask[(size_t)(rowt * kTileBoxes + r) * col_blocks + col] = bits;
    }
}

// ---- greedy scan: one CTA per scene -----------------------------------------------------------
// Tile by tile: the 64 mask rows of the tile (all words right of the diagonal) are staged in shared memory by the whole
// CTA — coalesced, and one tile AHEAD of the serial part — thread 0 walks the tile's 64 boxes against the diagonal word,
// then every thread ORs the kept rows into the removed-bits words it owns, from shared memory.  (The first version fetched
// every kept row's word from global memory inside the OR loop: up to 64 dependent L2 round trips per tile.)
constexpr int kScanThreads = 256;
constexpr int kScanMaxWords = 64;   // words per staged row: stride <= 4096 takes this path, larger strides the direct one

__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const unsigned long long *__restrict__ mask, const int *__restrict__ counts, int fixed_n, int stride,
                long long *__restrict__ keep, int *__restrict__ num_keep) {
    const int scene = blockIdx.x;
    const int n = scene_count(counts, scene, fixed_n, stride);
    const int col_blocks = pdab::div_up(stride, kTileBoxes);
    mask += (size_t)scene * stride * col_blocks;
    keep += (size_t)scene * stride;
    const int t = threadIdx.x;
    const int ntiles = pdab::div_up(n, kTileBoxes);

    extern __shared__ unsigned long long rows[];        // [2][64][ntiles]: the tile's mask rows, double buffered
    __shared__ unsigned long long remv[256];            // removed bits, one word per tile (stride <= 16384)
    __shared__ unsigned long long kept_bits;
    __shared__ int total;
    const bool staged = ntiles <= kScanMaxWords;
    for (int w = t; w < ntiles; w += kScanThreads) remv[w] = 0ull;
    if (t == 0) total = 0;

    auto stage = [&](int tile, int buf) {   // words [tile, ntiles) of rows tile*64 .. +63 (lower-triangle words are never read)
        const int size = min(n - tile * kTileBoxes, kTileBoxes);
        const int nw = ntiles - tile;
        unsigned long long *dst = rows + (size_t)buf * kTileBoxes * ntiles;
        for (int i = t; i < size * nw; i += kScanThreads) {
            const int q = i / nw, w = tile + (i - q * nw);
            dst[q * ntiles + w] = mask[(size_t)(tile * kTileBoxes + q) * col_blocks + w];
        }
    };
    if (staged && ntiles > 0) stage(0, 0);
    __syncthreads();
    for (int tile = 0; tile < ntiles; tile++) {
        const int size = min(n - tile * kTileBoxes, kTileBoxes);
        const int buf = tile & 1;
        const unsigned long long *cur_rows = rows + (size_t)buf * kTileBoxes * ntiles;
        if (staged && tile + 1 < ntiles) stage(tile + 1, buf ^ 1);      // overlaps the serial walk below
        if (t == 0) {
            unsigned long long cur = remv[tile], kb = 0ull;
            int tot = total;
            for (int q = 0; q < size; q++) {
                if (!((cur >> q) & 1ull)) {
                    kb |= 1ull << q;
                    keep[tot++] = tile * kTileBoxes + q;
                    cur |= staged ? cur_rows[q * ntiles + tile]
                                  : mask[(size_t)(tile * kTileBoxes + q) * col_blocks + tile];
                }
            }
            kept_bits = kb;
            total = tot;
        }
        __syncthreads();
        const unsigned long long kb = kept_bits;
        // OR the kept rows into the words right of this tile
        for (int w = tile + 1 + t; w < ntiles; w += kScanThreads) {
            unsigned long long acc = 0ull, bits = kb;
            while (bits) {
                const int q = __ffsll((long long)bits) - 1;
                bits &= bits - 1;
                acc |= staged ? cur_rows[q * ntiles + w] : mask[(size_t)(tile * kTileBoxes + q) * col_blocks + w];
            }
            remv[w] |= acc;
        }
        __syncthreads();
    }
    if (t == 0) num_keep[scene] = total;
}

__global__ void pairwise_kernel(int na, const float *__restrict__ a, int nb, const float *__restrict__ b,
                                float *__restrict__ out, int want_iou) {
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na || j >= nb) return;
    out[(size_t)i * nb + j] = want_iou ? iou_rotated(a + i * 7, b + j * 7)
                                       : (far_apart(a + i * 7, b + j * 7) ? 0.f : overlap_area(a + i * 7, b + j * 7));
}

int run_nms(const float *boxes, const int *counts, int nscenes, int n_or_stride, float thresh, long long *keep,
            int *num_keep, void *workspace, bool normal, cudaStream_t stream) {
    if (nscenes < 0 || n_or_stride < 0 || !num_keep) return PDAB_EINVAL;
    if (nscenes == 0) return 0;
    if (n_or_stride == 0) {
        PDAB_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int) * nscenes, stream));
        return 0;
    }
    if (!boxes || !keep || !workspace) return PDAB_EINVAL;
    if (n_or_stride > 16384 || nscenes > 65535) return PDAB_EUNSUPPORTED;
    const int tiles = pdab::div_up(n_or_stride, kTileBoxes);
    dim3 grid(tiles, tiles, nscenes);
    auto *mask = static_cast<unsigned long long *>(workspace);
    if (normal)
        nms_mask_kernel<true><<<grid, kMaskThreads, 0, stream>>>(boxes, counts, n_or_stride, n_or_stride, thresh, mask);
    else
        nms_mask_kernel<false><<<grid, kMaskThreads, 0, stream>>>(boxes, counts, n_or_stride, n_or_stride, thresh, mask);
    PDAB_LAUNCH_CHECK();
    // staged rows: 2 buffers x 64 rows x tiles words (tiles <= 64: at most 64 KB)
    const size_t smem = tiles <= kScanMaxWords ? sizeof(unsigned long long) * 2 * kTileBoxes * tiles : 0;
    if (smem > 48 * 1024)
        PDAB_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_scan_kernel<<<nscenes, kScanThreads, smem, stream>>>(mask, counts, n_or_stride, n_or_stride, keep, num_keep);
    PDAB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" size_t pdab_nms_workspace_bytes(int n) {
    if (n <= 0) return 0;
    return (size_t)n * pdab::div_up(n, kTileBoxes) * sizeof(unsigned long long);
}

extern "C" int pdab_nms_device(const float *boxes, int n, float thresh, int64_t *keep, int *num_keep, void *workspace,
                               pdab_stream_t stream) {
    return run_nms(boxes, nullptr, 1, n, thresh, reinterpret_cast<long long *>(keep), num_keep, workspace, false,
                   pdab::to_stream(stream));
}

extern "C" int pdab_nms_batched(const float *boxes, const int *counts, int nscenes, int stride, float thresh,
                                int64_t *keep, int *num_keep, void *workspace, pdab_stream_t stream) {
    if (!counts) return PDAB_EINVAL;
    return run_nms(boxes, counts, nscenes, stride, thresh, reinterpret_cast<long long *>(keep), num_keep, workspace,
                   false, pdab::to_stream(stream));
}

// Host-facing form (the reference pybind signature): device scratch and a pinned staging buffer are cached per host
// thread and device (grow-only; the reference does cudaMalloc + blocking cudaMemcpy + cudaFree per call), and the keep
// list comes back in ONE device-to-host copy of [num | keep] followed by one stream synchronisation.
namespace {
struct HostNmsCache {
    int device = -1;
    char *dev = nullptr;
    size_t dev_bytes = 0;
    char *pinned = nullptr;
    size_t pinned_bytes = 0;
    ~HostNmsCache() {
        if (dev) cudaFree(dev);
        if (pinned) cudaFreeHost(pinned);
    }
};
thread_local HostNmsCache g_host_nms;
}  // namespace

extern "C" int pdab_nms_host(const float *boxes, int n, float thresh, int64_t *keep_host, int normal,
                             pdab_stream_t stream) {
    if (n < 0 || (n > 0 && (!boxes || !keep_host))) return PDAB_EINVAL;
    if (n == 0) return 0;
    if (n > 16384) return PDAB_EUNSUPPORTED;
    cudaStream_t s = pdab::to_stream(stream);
    const size_t mask_bytes = pdab_nms_workspace_bytes(n);
    const size_t out_off = (mask_bytes + 255) / 256 * 256;
    const size_t out_bytes = sizeof(long long) * ((size_t)n + 1);      // [num | keep[0..n)]
    const size_t total = out_off + out_bytes;
    HostNmsCache &c = g_host_nms;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return -(int)e - 1000;
    if (c.device != dev || c.dev_bytes < total) {
        if (c.dev) cudaFree(c.dev);
        c.dev = nullptr;
        c.dev_bytes = 0;
        e = cudaMalloc(&c.dev, total * 2);                               // grow geometrically: few re-allocations
        if (e != cudaSuccess) return -(int)e - 1000;
        c.dev_bytes = total * 2;
        c.device = dev;
    }
    if (c.pinned_bytes < out_bytes) {
        if (c.pinned) cudaFreeHost(c.pinned);
        c.pinned = nullptr;
        c.pinned_bytes = 0;
        e = cudaMallocHost(&c.pinned, out_bytes * 2);
        if (e != cudaSuccess) return -(int)e - 1000;
        c.pinned_bytes = out_bytes * 2;
    }
    long long *out_dev = reinterpret_cast<long long *>(c.dev + out_off);
    int rc = run_nms(boxes, nullptr, 1, n, thresh, out_dev + 1, reinterpret_cast<int *>(out_dev), c.dev, normal != 0, s);
    if (rc > 0) return -rc - 1000;
    if (rc < 0) return rc;
    e = cudaMemcpyAsync(c.pinned, out_dev, out_bytes, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return -(int)e - 1000;
    const int num = *reinterpret_cast<const int *>(c.pinned);
    if (num < 0 || num > n) return PDAB_EINVAL;
    memcpy(keep_host, c.pinned + sizeof(long long), sizeof(long long) * (size_t)num);
    return num;
}

static int pairwise(int na, const float *a, int nb, const float *b, float *out, int want_iou, cudaStream_t s) {
    if (na < 0 || nb < 0) return PDAB_EINVAL;
    if (na == 0 || nb == 0) return 0;
    if (!a || !b || !out) return PDAB_EINVAL;
    dim3 block(16, 16);
    dim3 grid(pdab::div_up(nb, 16), pdab::div_up(na, 16));
    if (grid.y > 65535) return PDAB_EUNSUPPORTED;
    pairwise_kernel<<<grid, block, 0, s>>>(na, a, nb, b, out, want_iou);
    PDAB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pdab_boxes_overlap_bev(int na, const float *boxes_a, int nb, const float *boxes_b, float *out,
                                      pdab_stream_t stream) {
    return pairwise(na, boxes_a, nb, boxes_b, out, 0, pdab::to_stream(stream));
}

extern "C" int pdab_boxes_iou_bev(int na, const float *boxes_a, int nb, const float *boxes_b, float *out,
                                  pdab_stream_t stream) {
    return pairwise(na, boxes_a, nb, boxes_b, out, 1, pdab::to_stream(stream));
}
