// Batched post-processing around the NMS for sm_100a: two kernels replace the ~40 elementwise / reduce /
// sort / gather launches of Detector3DTemplate.post_processing + class_agnostic_nms for a whole batch
// (pcdet/models/detectors/detector3d_template.py:196-285, pcdet/models/model_utils/model_nms_utils.py:6-25), with no host
// synchronisation and fixed output shapes (so the whole forward can sit in a CUDA graph).
//
//   post_front  (one CTA per scene): per centre  p_c = sigmoid(logit_c), score = max_c p_c (first maximum wins, as torch.max),
//               label = argmax + 1, valid = score >= SCORE_THRESH; then the centres are ordered by (valid ? score : -inf)
//               descending, ties by ascending index — what `key.sort(descending=True, stable=True)` returns — with a bitonic
//               network on 64-bit keys in shared memory; writes the order, the boxes gathered into that order (the NMS
//               input) and counts = min(#valid, NMS_PRE_MAXSIZE).
//   [pdab_nms_batched]
//   post_select (one CTA per scene): the first min(num_keep, P) kept positions -> padded (B, P, .) outputs, zeros behind.
//
// Exactness: the sigmoid is 1 / (1 + expf(-x)) in fp32 with IEEE division — the expression ATen's CUDA kernel evaluates — so
// scores, the threshold test and every tie are bit-identical to the torch statement of the same steps (tests compare them).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxM = 4096;   // centres per scene: 32 KB of keys

__device__ __forceinline__ unsigned ordered_bits(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending in f
}

__global__ void __launch_bounds__(kThreads)
post_front_kernel(int m, int nc, int ldc, int ldb, int normalized, float thresh, int pre_max, const float *__restrict__ cls,
                  const float *__restrict__ boxes, int *__restrict__ order, float *__restrict__ sorted_boxes,
                  int *__restrict__ counts, float *__restrict__ scores, float *__restrict__ raw_max,
                  long long *__restrict__ labels) {
    extern __shared__ unsigned long long skeys[];   // next_pow2(m)
    __shared__ int s_valid;
    const int scene = blockIdx.x, t = threadIdx.x;
    cls += (size_t)scene * m * ldc;
    boxes += (size_t)scene * m * ldb;
    int P2 = 1;
    while (P2 < m) P2 <<= 1;
    if (t == 0) s_valid = 0;
    __syncthreads();
    int nvalid = 0;
    for (int i = t; i < P2; i += kThreads) {
        unsigned long long key = 0ull;          // padding sorts last
        if (i < m) {
            const float *r = cls + (size_t)i * ldc;
            float best = 0.f, braw = 0.f;
            int lab = 0;
            for (int c = 0; c < nc; c++) {
                const float x = __ldg(r + c);
                const float p = normalized ? x : 1.0f / (1.0f + expf(-x));
                if (c == 0 || p > best) {        // strict '>': the first maximum wins
                    best = p;
                    lab = c;
                }
                if (c == 0 || x > braw) braw = x;
            }
            scores[(size_t)scene * m + i] = best;
            raw_max[(size_t)scene * m + i] = braw;
            labels[(size_t)scene * m + i] = lab + 1;
            const bool valid = best >= thresh;
            nvalid += valid ? 1 : 0;
            const float k = valid ? best : __int_as_float(0xff800000);   // -inf
            key = ((unsigned long long)ordered_bits(k) << 32) | (unsigned long long)(~(unsigned)i);
        }
        skeys[i] = key;
    }
    if (nvalid) atomicAdd(&s_valid, nvalid);
    __syncthreads();
    for (int size = 2; size <= P2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = t; i < (P2 >> 1); i += kThreads) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long a = skeys[lo], b = skeys[hi];
                if ((a < b) == desc) {
                    skeys[lo] = b;
                    skeys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int j = t; j < m; j += kThreads) {
        const int i = (int)(~(unsigned)skeys[j]);
        order[(size_t)scene * m + j] = i;
        const float *b = boxes + (size_t)i * ldb;
        float *o = sorted_boxes + ((size_t)scene * m + j) * 7;
#pragma unroll
        for (int q = 0; q < 7; q++) o[q] = __ldg(b + q);
    }
    if (t == 0) counts[scene] = min(s_valid, pre_max);
}

__global__ void __launch_bounds__(kThreads)
post_select_kernel(int m, int p, int nb, int ldb, const long long *__restrict__ keep, const int *__restrict__ num_keep,
                   const int *__restrict__ order, const float *__restrict__ boxes, const float *__restrict__ scores,
                   const long long *__restrict__ labels, float *__restrict__ out_boxes, float *__restrict__ out_scores,
                   long long *__restrict__ out_labels, int *__restrict__ out_num) {
    const int scene = blockIdx.x;
    const int num = min(max(num_keep[scene], 0), p);
    if (threadIdx.x == 0) out_num[scene] = num;
    for (int j = threadIdx.x; j < p; j += kThreads) {
        float *ob = out_boxes + ((size_t)scene * p + j) * nb;
        if (j < num) {
            const int pos = (int)keep[(size_t)scene * m + j];
            const int i = order[(size_t)scene * m + pos];
            const float *b = boxes + ((size_t)scene * m + i) * ldb;
            for (int q = 0; q < nb; q++) ob[q] = __ldg(b + q);
            out_scores[(size_t)scene * p + j] = scores[(size_t)scene * m + i];
            out_labels[(size_t)scene * p + j] = labels[(size_t)scene * m + i];
        } else {
            for (int q = 0; q < nb; q++) ob[q] = 0.f;
            out_scores[(size_t)scene * p + j] = 0.f;
            out_labels[(size_t)scene * p + j] = 0;
        }
    }
}

}  // namespace

extern "C" int pdab_post_front(int b, int m, int nc, int ldc, int ldb, int normalized, float score_thresh, int pre_max,
                               const float *cls, const float *boxes, int *order, float *sorted_boxes, int *counts,
                               float *scores, float *raw_max, int64_t *labels, pdab_stream_t stream) {
    if (b < 0 || m < 1 || nc < 1 || ldc < nc || ldb < 7 || !cls || !boxes || !order || !sorted_boxes || !counts || !scores ||
        !raw_max || !labels)
        return PDAB_EINVAL;
    if (b == 0) return 0;
    if (m > kMaxM || b > 65535) return PDAB_EUNSUPPORTED;
    int P2 = 1;
    while (P2 < m) P2 <<= 1;
    post_front_kernel<<<b, kThreads, (size_t)P2 * sizeof(unsigned long long), pdab::to_stream(stream)>>>(
        m, nc, ldc, ldb, normalized, score_thresh, pre_max, cls, boxes, order, sorted_boxes, counts, scores, raw_max,
        reinterpret_cast<long long *>(labels));
    PDAB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pdab_post_select(int b, int m, int p, int nb, int ldb, const int64_t *keep, const int *num_keep,
                                const int *order, const float *boxes, const float *scores, const int64_t *labels,
                                float *out_boxes, float *out_scores, int64_t *out_labels, int *out_num, pdab_stream_t stream) {
    if (b < 0 || m < 1 || p < 1 || nb < 1 || ldb < nb || !keep || !num_keep || !order || !boxes || !scores || !labels ||
        !out_boxes || !out_scores || !out_labels || !out_num)
        return PDAB_EINVAL;
    if (b == 0) return 0;
    if (b > 65535) return PDAB_EUNSUPPORTED;
    post_select_kernel<<<b, kThreads, 0, pdab::to_stream(stream)>>>(
        m, p, nb, ldb, reinterpret_cast<const long long *>(keep), num_keep, order, boxes, scores,
        reinterpret_cast<const long long *>(labels), out_boxes, out_scores, reinterpret_cast<long long *>(out_labels), out_num);
    PDAB_LAUNCH_CHECK();
    return 0;
}
