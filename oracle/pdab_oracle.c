/*
 * TEST INFRASTRUCTURE — CPU oracle for the PDA-SSD point-backbone hot path.
 *
 * A plain-C restatement of the reference's algorithms (Geo3DSmart/PDANet;
 * PB = pcdet/ops/pointnet2/pointnet2_batch, IOU = pcdet/ops/iou3d_nms).  It is
 * the checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg; the product (pdanet_b200/) never imports, links or calls it.
 *
 * Pinning status:
 *   - orc_boxes_iou_bev / orc_box_overlap are pinned bit-exactly HERE (CPU)
 *     against the reference's own boxes_iou_bev_cpu (IOU/src/iou3d_cpu.cpp,
 *     built unmodified into oracle/_ref/iou3d_nms_cuda.so) — tests/test_oracle_pin.py.
 *   - FPS / ball query / gather / group / NMS keep lists are pinned against the
 *     reference's CUDA kernels run on a B200 (oracle/_ref/*.so, golden fixtures
 *     in tests/golden/ made by tests/golden/make_golden.py).  The reference
 *     ships no golden vectors or tests of its own.
 *
 * Floating point: distances use fmaf() in the order nvcc 12.9 compiles the
 * reference expression for sm_100a (checked in SASS of oracle/_ref objects):
 *     t = rn(dy*dy); t = fma(dx,dx,t); d = fma(dz,dz,t)
 * Build with -ffp-contract=off so the compiler adds no contraction of its own.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ helpers */

/* PB/src/cuda_utils.h:10-14 — largest power of two <= min(work_size, 1024). */
int orc_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

/* squared distance in the reference's compiled op order (see header). */
static inline float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

/* --------------------------------------------------------------------- FPS */

/* Literal emulation of one reference FPS block (PB/src/sampling_gpu.cu:113-208,
 * tree reduction :143-203, __update :86-91).  `row(old,k)` supplies the
 * distance of point k to the last pick. */
typedef float (*dist_fn)(const void *ctx, int n, int old, int k);

static void fps_scene(const void *ctx, dist_fn dist, int n, int m, float *temp, int *idxs) {
    if (m <= 0) return;
    const int bs = orc_opt_n_threads(n);
    float *dists = (float *)malloc(sizeof(float) * bs);
    int *dists_i = (int *)malloc(sizeof(int) * bs);
    int old = 0;
    idxs[0] = old;
    for (int j = 1; j < m; j++) {
        for (int tid = 0; tid < bs; tid++) {
            int besti = 0;
            float best = -1.0f;
            for (int k = tid; k < n; k += bs) {
                const float d = dist(ctx, n, old, k);
                const float d2 = fminf(d, temp[k]);
                temp[k] = d2;
                besti = d2 > best ? k : besti;
                best = d2 > best ? d2 : best;
            }
            dists[tid] = best;
            dists_i[tid] = besti;
        }
        for (int s = bs / 2; s >= 1; s >>= 1) {
            for (int tid = 0; tid < s; tid++) {
                const float v1 = dists[tid], v2 = dists[tid + s];
                const int i1 = dists_i[tid], i2 = dists_i[tid + s];
                dists[tid] = v1 > v2 ? v1 : v2; /* max(v1, v2) */
                dists_i[tid] = v2 > v1 ? i2 : i1;
            }
        }
        old = dists_i[0];
        idxs[j] = old;
    }
    free(dists);
    free(dists_i);
}

static float dist_xyz(const void *ctx, int n, int old, int k) {
    (void)n;
    const float *p = (const float *)ctx;
    /* reference: (x2-x1)^2 ... with x2 = point k, x1 = last pick */
    return sqdist3(p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2], p[old * 3 + 0], p[old * 3 + 1], p[old * 3 + 2]);
}

static float dist_matrix(const void *ctx, int n, int old, int k) {
    const float *d = (const float *)ctx;
    return d[(size_t)old * n + k]; /* PB/src/sampling_gpu.cu:294 */
}

/* D-FPS.  xyz (B,N,3), temp (B,N) in/out (caller fills 1e10: PB/pointnet2_utils.py:26), idx (B,m). */
void orc_fps(int b, int n, int m, const float *xyz, float *temp, int *idx) {
    for (int i = 0; i < b; i++)
        fps_scene(xyz + (size_t)i * n * 3, dist_xyz, n, m, temp + (size_t)i * n, idx + (size_t)i * m);
}

/* F-FPS from a precomputed (B,N,N) distance matrix (PB/src/sampling_gpu.cu:256-416). */
void orc_fps_with_dist(int b, int n, int m, const float *dist, float *temp, int *idx) {
    for (int i = 0; i < b; i++)
        fps_scene(dist + (size_t)i * n * n, dist_matrix, n, m, temp + (size_t)i * n, idx + (size_t)i * m);
}

/* ------------------------------------------------------------ gather / group */

/* PB/src/sampling_gpu.cu:8-23 */
void orc_gather(int b, int c, int n, int np, const float *points, const int *idx, float *out) {
    for (int bi = 0; bi < b; bi++)
        for (int ci = 0; ci < c; ci++)
            for (int j = 0; j < np; j++)
                out[((size_t)bi * c + ci) * np + j] = points[((size_t)bi * c + ci) * n + idx[(size_t)bi * np + j]];
}

/* PB/src/sampling_gpu.cu:46-62 (atomicAdd scatter; summed here in index order) */
void orc_gather_grad(int b, int c, int n, int np, const float *grad_out, const int *idx, float *grad_points) {
    for (int bi = 0; bi < b; bi++)
        for (int ci = 0; ci < c; ci++)
            for (int j = 0; j < np; j++)
                grad_points[((size_t)bi * c + ci) * n + idx[(size_t)bi * np + j]] += grad_out[((size_t)bi * c + ci) * np + j];
}

/* PB/src/group_points_gpu.cu:53-71 */
void orc_group(int b, int c, int n, int np, int ns, const float *points, const int *idx, float *out) {
    for (int bi = 0; bi < b; bi++)
        for (int ci = 0; ci < c; ci++)
            for (int j = 0; j < np; j++)
                for (int s = 0; s < ns; s++)
                    out[(((size_t)bi * c + ci) * np + j) * ns + s] =
                        points[((size_t)bi * c + ci) * n + idx[((size_t)bi * np + j) * ns + s]];
}

/* PB/src/group_points_gpu.cu:14-31 */
void orc_group_grad(int b, int c, int n, int np, int ns, const float *grad_out, const int *idx, float *grad_points) {
    for (int bi = 0; bi < b; bi++)
        for (int ci = 0; ci < c; ci++)
            for (int j = 0; j < np; j++)
                for (int s = 0; s < ns; s++)
                    grad_points[((size_t)bi * c + ci) * n + idx[((size_t)bi * np + j) * ns + s]] +=
                        grad_out[(((size_t)bi * c + ci) * np + j) * ns + s];
}

/* -------------------------------------------------------------- ball query */

/* PB/src/ball_query_gpu.cu:9-45.  idx (B,M,nsample) must be pre-zeroed by the
 * caller (PB/pointnet2_utils.py:246): an empty ball keeps its zero row. */
void orc_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx) {
    const float radius2 = radius * radius;
    for (int bi = 0; bi < b; bi++) {
        const float *P = xyz + (size_t)bi * n * 3;
        for (int j = 0; j < m; j++) {
            const float *c = new_xyz + ((size_t)bi * m + j) * 3;
            int *o = idx + ((size_t)bi * m + j) * nsample;
            int cnt = 0;
            for (int k = 0; k < n; k++) {
                const float d2 = sqdist3(c[0], c[1], c[2], P[k * 3], P[k * 3 + 1], P[k * 3 + 2]);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; l++) o[l] = k;
                    o[cnt] = k;
                    if (++cnt >= nsample) break;
                }
            }
        }
    }
}

/* PB/src/ball_query_gpu.cu:70-113.  A point at distance exactly 0 that also
 * lies in [min_r^2, max_r^2) is emitted twice, as in the reference. */
void orc_ball_query_dilated(int b, int n, int m, float max_radius, float min_radius, int nsample,
                            const float *new_xyz, const float *xyz, int *idx) {
    const float radius1 = max_radius * max_radius;
    const float radius2 = min_radius * min_radius;
    for (int bi = 0; bi < b; bi++) {
        const float *P = xyz + (size_t)bi * n * 3;
        for (int j = 0; j < m; j++) {
            const float *c = new_xyz + ((size_t)bi * m + j) * 3;
            int *o = idx + ((size_t)bi * m + j) * nsample;
            int cnt = 0;
            for (int k = 0; k < n; k++) {
                const float d2 = sqdist3(c[0], c[1], c[2], P[k * 3], P[k * 3 + 1], P[k * 3 + 2]);
                if (d2 == 0) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; l++) o[l] = k;
                    o[cnt] = k;
                    if (++cnt >= nsample) break;
                }
                if (d2 >= radius2 && d2 < radius1) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; l++) o[l] = k;
                    o[cnt] = k;
                    if (++cnt >= nsample) break;
                }
            }
        }
    }
}

/* --------------------------------------------------- rotated BEV IoU + NMS */
/* Geometry follows IOU/src/iou3d_cpu.cpp:59-229 (host twin of
 * IOU/src/iou3d_nms_kernel.cu:35-234): no FMA contraction, libm float
 * trigonometry, float arithmetic throughout (the reference's double
 * constructor arguments and the /2.0 are exact in either precision). */

typedef struct { float x, y; } pt2;

static inline float cross2(pt2 a, pt2 b) { return a.x * b.y - a.y * b.x; }
static inline float cross3(pt2 p1, pt2 p2, pt2 p0) {
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
static inline float fmin_(float a, float b) { return a > b ? b : a; }
static inline float fmax_(float a, float b) { return a > b ? a : b; }

static int rect_cross(pt2 p1, pt2 p2, pt2 q1, pt2 q2) {
    return fmin_(p1.x, p2.x) <= fmax_(q1.x, q2.x) && fmin_(q1.x, q2.x) <= fmax_(p1.x, p2.x) &&
           fmin_(p1.y, p2.y) <= fmax_(q1.y, q2.y) && fmin_(q1.y, q2.y) <= fmax_(p1.y, p2.y);
}

static int in_box2d(const float *box, pt2 p) {
    const float MARGIN = 1e-2f;
    const float cx = box[0], cy = box[1];
    const float ac = cosf(-box[6]), as = sinf(-box[6]);
    const float rx = (p.x - cx) * ac + (p.y - cy) * (-as);
    const float ry = (p.x - cx) * as + (p.y - cy) * ac;
    return fabsf(rx) < box[3] / 2 + MARGIN && fabsf(ry) < box[4] / 2 + MARGIN;
}

static int seg_intersection(pt2 p1, pt2 p0, pt2 q1, pt2 q0, pt2 *ans) {
    const float EPS = 1e-8f;
    if (!rect_cross(p0, p1, q0, q1)) return 0;
    const float s1 = cross3(q0, p1, p0);
    const float s2 = cross3(p1, q1, p0);
    const float s3 = cross3(p0, q1, q0);
    const float s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    const float s5 = cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > EPS) {
        ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float D = a0 * b1 - a1 * b0;
        ans->x = (b0 * c1 - b1 * c0) / D;
        ans->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}

static pt2 rot_about(pt2 c, float ac, float as, pt2 p) {
    pt2 r;
    r.x = (p.x - c.x) * ac + (p.y - c.y) * (-as) + c.x;
    r.y = (p.x - c.x) * as + (p.y - c.y) * ac + c.y;
    return r;
}

float orc_box_overlap(const float *A, const float *B) {
    const float a_ang = A[6], b_ang = B[6];
    const float ahx = A[3] / 2, bhx = B[3] / 2, ahy = A[4] / 2, bhy = B[4] / 2;
    const float ax1 = A[0] - ahx, ay1 = A[1] - ahy, ax2 = A[0] + ahx, ay2 = A[1] + ahy;
    const float bx1 = B[0] - bhx, by1 = B[1] - bhy, bx2 = B[0] + bhx, by2 = B[1] + bhy;
    const pt2 ca = {A[0], A[1]}, cb = {B[0], B[1]};
    pt2 qa[5] = {{ax1, ay1}, {ax2, ay1}, {ax2, ay2}, {ax1, ay2}, {0, 0}};
    pt2 qb[5] = {{bx1, by1}, {bx2, by1}, {bx2, by2}, {bx1, by2}, {0, 0}};
    const float acs = cosf(a_ang), asn = sinf(a_ang), bcs = cosf(b_ang), bsn = sinf(b_ang);
    for (int k = 0; k < 4; k++) {
        qa[k] = rot_about(ca, acs, asn, qa[k]);
        qb[k] = rot_about(cb, bcs, bsn, qb[k]);
    }
    qa[4] = qa[0];
    qb[4] = qb[0];

    pt2 poly[16], ctr = {0.f, 0.f};
    int cnt = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            if (seg_intersection(qa[i + 1], qa[i], qb[j + 1], qb[j], &poly[cnt])) {
                ctr.x = ctr.x + poly[cnt].x;
                ctr.y = ctr.y + poly[cnt].y;
                cnt++;
            }
    for (int k = 0; k < 4; k++) {
        if (in_box2d(A, qb[k])) { ctr.x += qb[k].x; ctr.y += qb[k].y; poly[cnt++] = qb[k]; }
        if (in_box2d(B, qa[k])) { ctr.x += qa[k].x; ctr.y += qa[k].y; poly[cnt++] = qa[k]; }
    }
    ctr.x /= cnt; /* cnt == 0 -> inf/nan, never used (no area terms) */
    ctr.y /= cnt;
    for (int j = 0; j < cnt - 1; j++)
        for (int i = 0; i < cnt - j - 1; i++)
            if (atan2f(poly[i].y - ctr.y, poly[i].x - ctr.x) > atan2f(poly[i + 1].y - ctr.y, poly[i + 1].x - ctr.x)) {
                pt2 t = poly[i]; poly[i] = poly[i + 1]; poly[i + 1] = t;
            }
    float area = 0;
    for (int k = 0; k < cnt - 1; k++) {
        pt2 u = {poly[k].x - poly[0].x, poly[k].y - poly[0].y};
        pt2 v = {poly[k + 1].x - poly[0].x, poly[k + 1].y - poly[0].y};
        area += cross2(u, v);
    }
    return fabsf(area) / 2.0f;
}

float orc_iou_bev(const float *A, const float *B) {
    const float sa = A[3] * A[4], sb = B[3] * B[4];
    const float so = orc_box_overlap(A, B);
    return so / fmaxf(sa + sb - so, 1e-8f);
}

/* IOU/src/iou3d_nms_kernel.cu:313-325 */
float orc_iou_normal(const float *a, const float *b) {
    const float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    const float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    const float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
    const float inter = width * height;
    const float Sa = a[3] * a[4], Sb = b[3] * b[4];
    return inter / fmaxf(Sa + Sb - inter, 1e-8f);
}

void orc_boxes_overlap_bev(int na, const float *a, int nb, const float *b, float *out) {
    for (int i = 0; i < na; i++)
        for (int j = 0; j < nb; j++) out[(size_t)i * nb + j] = orc_box_overlap(a + i * 7, b + j * 7);
}

void orc_boxes_iou_bev(int na, const float *a, int nb, const float *b, float *out) {
    for (int i = 0; i < na; i++)
        for (int j = 0; j < nb; j++) out[(size_t)i * nb + j] = orc_iou_bev(a + i * 7, b + j * 7);
}

/* Bitmask NMS: mask as IOU/src/iou3d_nms_kernel.cu:267-311 (bit j of row i set
 * iff j > i within the diagonal tile / any j in tiles right of it, and
 * iou > thresh), greedy reduce as IOU/src/iou3d_nms.cpp:115-132.  boxes are
 * already sorted by descending score (IOU/iou3d_nms_utils.py:92-96).
 * keep receives positions into the sorted list; returns num_to_keep. */
static int nms_generic(const float *boxes, int n, float thresh, int64_t *keep, int normal) {
    const int cb = (n + 63) / 64;
    uint64_t *mask = (uint64_t *)calloc((size_t)n * cb + 1, sizeof(uint64_t));
    uint64_t *remv = (uint64_t *)calloc((size_t)cb + 1, sizeof(uint64_t));
    for (int i = 0; i < n; i++) {
        const int rt = i / 64;
        for (int ct = rt; ct < cb; ct++) { /* tiles left of the diagonal are computed but never read */
            const int cs = n - ct * 64 < 64 ? n - ct * 64 : 64;
            uint64_t t = 0;
            for (int q = (ct == rt ? (i % 64) + 1 : 0); q < cs; q++) {
                const float *bj = boxes + (size_t)(ct * 64 + q) * 7;
                const float v = normal ? orc_iou_normal(boxes + (size_t)i * 7, bj) : orc_iou_bev(boxes + (size_t)i * 7, bj);
                if (v > thresh) t |= 1ULL << q;
            }
            mask[(size_t)i * cb + ct] = t;
        }
    }
    int num = 0;
    for (int i = 0; i < n; i++) {
        const int nb = i / 64, ib = i % 64;
        if (!(remv[nb] & (1ULL << ib))) {
            keep[num++] = i;
            for (int j = nb; j < cb; j++) remv[j] |= mask[(size_t)i * cb + j];
        }
    }
    free(mask);
    free(remv);
    return num;
}

int orc_nms(const float *boxes, int n, float thresh, int64_t *keep) { return nms_generic(boxes, n, thresh, keep, 0); }
int orc_nms_normal(const float *boxes, int n, float thresh, int64_t *keep) { return nms_generic(boxes, n, thresh, keep, 1); }

/* ------------------------------------------------- class-aware top-k sampling */
/* PB/pointnet2_modules.py:761-770: score = sigmoid(max_c logits); topk(npoint)
 * sorted by descending score.  torch.topk leaves tie order unspecified; the
 * canonical order here (and in the CUDA kernel) is (max-logit desc, index asc),
 * which is one valid topk order of the sigmoid scores since sigmoid is monotone.
 * cls (B,N,C) -> idx (B,npoint) int32. */
typedef struct { float v; int i; } vi;
static int cmp_vi(const void *a, const void *b) {
    const vi *x = (const vi *)a, *y = (const vi *)b;
    if (x->v > y->v) return -1;
    if (x->v < y->v) return 1;
    return x->i < y->i ? -1 : (x->i > y->i ? 1 : 0);
}
void orc_topk_ctr(int b, int n, int c, int npoint, const float *cls, int *idx) {
    vi *buf = (vi *)malloc(sizeof(vi) * n);
    for (int bi = 0; bi < b; bi++) {
        for (int k = 0; k < n; k++) {
            const float *r = cls + ((size_t)bi * n + k) * c;
            float mx = r[0];
            for (int q = 1; q < c; q++) mx = r[q] > mx ? r[q] : mx;
            buf[k].v = mx;
            buf[k].i = k;
        }
        qsort(buf, n, sizeof(vi), cmp_vi);
        for (int j = 0; j < npoint; j++) idx[(size_t)bi * npoint + j] = buf[j].i;
    }
    free(buf);
}

/* ------------------------------------------------------------- PDA grouper */
/* PB/pointnet2_utils.py:586-607 given idx from ball query: out (B, 7+C, M, ns),
 * channels [xyz(3, NOT centred), density(1), direction(3), features(C)].
 * density = exp(-|g-c|^2 / (2 r^2)) / (2.5 r), |g-c| via sqrt then squared as
 * torch.norm(...)**2 does; direction = (g-c)/r.  Python-double scalars are
 * rounded to fp32 before use, as torch does for tensor-scalar ops. */
void orc_pda_group(int b, int c, int n, int m, int ns, float radius, const float *xyz, const float *new_xyz,
                   const float *features, const int *idx, float *out) {
    const float two_r2 = (float)(2.0 * (double)radius * (double)radius);
    const float norm = (float)(2.5 * (double)radius);
    const int ch = 7 + c;
    for (int bi = 0; bi < b; bi++)
        for (int j = 0; j < m; j++) {
            const float *ctr = new_xyz + ((size_t)bi * m + j) * 3;
            for (int s = 0; s < ns; s++) {
                const int k = idx[((size_t)bi * m + j) * ns + s];
                const float *g = xyz + ((size_t)bi * n + k) * 3;
                const float dx = g[0] - ctr[0], dy = g[1] - ctr[1], dz = g[2] - ctr[2];
                const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
                const float dens = expf(-(dist * dist) / two_r2) / norm;
                float *o = out + ((size_t)bi * ch * m + j) * ns + s;
                const size_t cs = (size_t)m * ns;
                o[0 * cs] = g[0]; o[1 * cs] = g[1]; o[2 * cs] = g[2];
                o[3 * cs] = dens;
                o[4 * cs] = dx / radius; o[5 * cs] = dy / radius; o[6 * cs] = dz / radius;
                for (int q = 0; q < c; q++) o[(7 + q) * cs] = features[((size_t)bi * c + q) * n + k];
            }
        }
}

/* ------------------------------------------- fused plain-SA scale (fp32 CPU) */
/* QueryAndGroup (PB/pointnet2_utils.py:689-704) + shared MLP with eval-mode
 * BatchNorm folded into (W, bias) + ReLU per layer + max over nsample
 * (PB/pointnet2_modules.py:1655-1672).  idx from orc_ball_query.
 * Input channel order: [xyz - centre (3), features (C)].
 * W[l] is (cout_l, cin_l) row-major, bias[l] is (cout_l).  out (B, cout_last, M). */
void orc_sa_mlp_maxpool(int b, int c, int n, int m, int ns, const float *xyz, const float *new_xyz,
                        const float *features, const int *idx, int nlayers, const int *dims,
                        const float *const *W, const float *const *bias, float *out) {
    int maxd = 0;
    for (int l = 0; l <= nlayers; l++) maxd = dims[l] > maxd ? dims[l] : maxd;
    float *a0 = (float *)malloc(sizeof(float) * maxd), *a1 = (float *)malloc(sizeof(float) * maxd);
    const int cout = dims[nlayers];
    for (int bi = 0; bi < b; bi++)
        for (int j = 0; j < m; j++) {
            const float *ctr = new_xyz + ((size_t)bi * m + j) * 3;
            for (int s = 0; s < ns; s++) {
                const int k = idx[((size_t)bi * m + j) * ns + s];
                const float *g = xyz + ((size_t)bi * n + k) * 3;
                a0[0] = g[0] - ctr[0]; a0[1] = g[1] - ctr[1]; a0[2] = g[2] - ctr[2];
                for (int q = 0; q < c; q++) a0[3 + q] = features[((size_t)bi * c + q) * n + k];
                float *in = a0, *o = a1;
                for (int l = 0; l < nlayers; l++) {
                    const int ci = dims[l], co = dims[l + 1];
                    for (int r = 0; r < co; r++) {
                        float acc = 0.f;
                        const float *w = W[l] + (size_t)r * ci;
                        for (int q = 0; q < ci; q++) acc = fmaf(w[q], in[q], acc);
                        acc += bias[l][r];
                        o[r] = acc > 0.f ? acc : 0.f;
                    }
                    float *t = in; in = o; o = t;
                }
                for (int r = 0; r < cout; r++) {
                    float *dst = out + ((size_t)bi * cout + r) * m + j;
                    if (s == 0 || in[r] > *dst) *dst = in[r];
                }
            }
        }
    free(a0);
    free(a1);
}

/* -------------------------------------------------------------- feature propagation */

/* PB/src/interpolate_gpu.cu:16-59.  Three nearest `known` points of every `unknown` point: scan k ascending, strict '<'
 * insertion (lowest index wins ties); the reference keeps the bests as doubles starting at 1e40 and compares the fp32
 * distance against them, so an infinite / NaN distance never enters and missing entries read (float)1e40 = +inf, idx 0.
 * Distance in the compiled op order (SASS of the rebuilt reference): rn(dy*dy), fma(dx,dx,.), fma(dz,dz,.). */
void orc_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx) {
    for (int bi = 0; bi < b; bi++)
        for (int i = 0; i < n; i++) {
            const float *u = unknown + ((size_t)bi * n + i) * 3;
            const float *K = known + (size_t)bi * m * 3;
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; k++) {
                const float d = sqdist3(u[0], u[1], u[2], K[k * 3], K[k * 3 + 1], K[k * 3 + 2]);
                if (d < best1) {
                    best3 = best2; i3 = i2;
                    best2 = best1; i2 = i1;
                    best1 = d; i1 = k;
                } else if (d < best2) {
                    best3 = best2; i3 = i2;
                    best2 = d; i2 = k;
                } else if (d < best3) {
                    best3 = d; i3 = k;
                }
            }
            float *o = dist2 + ((size_t)bi * n + i) * 3;
            int *oi = idx + ((size_t)bi * n + i) * 3;
            o[0] = (float)best1; o[1] = (float)best2; o[2] = (float)best3;
            oi[0] = i1; oi[1] = i2; oi[2] = i3;
        }
}

/* PB/src/interpolate_gpu.cu:84-101: out = w0*p0 + w1*p1 + w2*p2, contracted by nvcc as
 * fma(w2, p2, fma(w0, p0, rn(w1*p1))) (SASS of the rebuilt reference). */
void orc_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight,
                           float *out) {
    for (int bi = 0; bi < b; bi++)
        for (int ci = 0; ci < c; ci++) {
            const float *p = points + ((size_t)bi * c + ci) * m;
            for (int i = 0; i < n; i++) {
                const int *id = idx + ((size_t)bi * n + i) * 3;
                const float *w = weight + ((size_t)bi * n + i) * 3;
                float t = w[1] * p[id[1]];
                t = fmaf(w[0], p[id[0]], t);
                t = fmaf(w[2], p[id[2]], t);
                out[((size_t)bi * c + ci) * n + i] = t;
            }
        }
}

/* PB/src/interpolate_gpu.cu:127-147 (the reference scatters with float atomics; summed here in index order). */
void orc_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight,
                                float *grad_points) {
    for (int bi = 0; bi < b; bi++)
        for (int ci = 0; ci < c; ci++) {
            float *gp = grad_points + ((size_t)bi * c + ci) * m;
            for (int i = 0; i < n; i++) {
                const int *id = idx + ((size_t)bi * n + i) * 3;
                const float *w = weight + ((size_t)bi * n + i) * 3;
                const float g = grad_out[((size_t)bi * c + ci) * n + i];
                for (int k = 0; k < 3; k++) gp[id[k]] += g * w[k];
            }
        }
}

/* -------------------------------------------------------------- points in boxes */

/* pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:16-38,313-338.  First box (lowest k) containing the point;
 * out must be pre-filled with -1 by the caller (roiaware_pool3d_utils.py:38).  Arithmetic as compiled for sm_100a:
 * local_x = fma(sx, cosa, -rn(sy*sina)), local_y = fma(sy, cosa, rn(sx*sina)) with cosa = cosf(-rz), sina = sinf(-rz);
 * the three extent tests run in double.  (libm cosf/sinf may differ from libdevice in the last ulp: a point within
 * ~1e-7 of a face can flip; the GPU kernel is pinned against the reference kernel itself in tests/test_gpu_ops.py.) */
void orc_points_in_boxes(int batch, int nboxes, int npts, const float *boxes, const float *pts, int *out) {
    for (int bi = 0; bi < batch; bi++)
        for (int i = 0; i < npts; i++) {
            const float *p = pts + ((size_t)bi * npts + i) * 3;
            for (int k = 0; k < nboxes; k++) {
                const float *bx = boxes + ((size_t)bi * nboxes + k) * 7;
                if ((double)fabsf(p[2] - bx[2]) > (double)bx[5] / 2.0) continue;
                const float sx = p[0] - bx[0], sy = p[1] - bx[1];
                const float cosa = cosf(-bx[6]), sina = sinf(-bx[6]);
                const float lx = fmaf(sx, cosa, -(sy * sina));
                const float ly = fmaf(sy, cosa, sx * sina);
                const float margin = 1e-5f;
                if ((double)fabsf(lx) < (double)bx[3] / 2.0 + (double)margin &&
                    (double)fabsf(ly) < (double)bx[4] / 2.0 + (double)margin) {
                    out[(size_t)bi * npts + i] = k;
                    break;
                }
            }
        }
}
