// Class-aware ("ctr_aware") top-k sampling for sm_100a.
//
// Replaces max-over-classes -> sigmoid -> torch.topk -> .int()
// (PB/pointnet2_modules.py:761-770) with one CTA per scene:
//   1. every point gets a 64-bit key [ordered fp32 bits of max_c logit | ~index]
//      (sigmoid is monotone, so selecting on the raw maximum logit selects the
//      same set; the low word makes keys unique and breaks ties towards the
//      lower index);
//   2. an 8-bit-per-pass MSB radix select over the keys (warp-aggregated
//      shared-memory histograms) finds the npoint-th largest key;
//   3. the npoint survivors are compacted into shared memory and ordered by a
//      bitonic network, because the reference's consumers rely on the
//      descending-score order of topk.
// Keys are recomputed from global memory on every pass (the class scores of a
// scene are at most N*C*4 bytes and stay in L1/L2), so N is bounded only by the
// survivor buffer: npoint <= 8192.
#include "common.cuh"

namespace {

constexpr int kThreads = 1024;
constexpr int kMaxSelect = 8192;  // 64 KB of keys

__device__ __forceinline__ unsigned ordered_bits(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending in f
}

__device__ __forceinline__ unsigned long long make_key(const float *__restrict__ cls, int k, int c) {
    const float *r = cls + (size_t)k * c;
    float mx = __ldg(r);
    for (int q = 1; q < c; q++) {
        const float v = __ldg(r + q);
        mx = v > mx ? v : mx;  // same NaN behaviour as a '>' scan
    }
    return ((unsigned long long)ordered_bits(mx) << 32) | (unsigned long long)(~(unsigned)k);
}

__global__ void __launch_bounds__(kThreads)
topk_ctr_kernel(int n, int c, int npoint, const float *__restrict__ cls, int *__restrict__ idx) {
    extern __shared__ unsigned long long skeys[];  // next_pow2(npoint)
    __shared__ unsigned hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_remaining, s_fill;

    const int scene = blockIdx.x;
    const int t = threadIdx.x;
    cls += (size_t)scene * n * c;
    idx += (size_t)scene * npoint;

    // ---- radix select: find T = npoint-th largest key --------------------------------------
    unsigned long long prefix = 0ull, prefix_mask = 0ull;
    int remaining = npoint;  // rank (1-based, from the top) still to resolve inside the prefix bucket
    for (int shift = 56; shift >= 0; shift -= 8) {
        if (t < 256) hist[t] = 0u;
        __syncthreads();
        for (int k = t; k < n; k += kThreads) {
            const unsigned long long key = make_key(cls, k, c);
            if ((key & prefix_mask) == prefix) atomicAdd(&hist[(unsigned)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (t == 0) {
            int rem = remaining;
            int d = 255;
            for (; d > 0; d--) {
                const int h = (int)hist[d];
                if (rem <= h) break;
                rem -= h;
            }
            s_prefix = prefix | ((unsigned long long)d << shift);
            s_remaining = rem;
        }
        __syncthreads();
        prefix = s_prefix;
        remaining = s_remaining;
        prefix_mask |= 0xffull << shift;
        __syncthreads();
    }
    const unsigned long long threshold = prefix;  // keys are unique: exactly npoint keys are >= threshold

    // ---- compact survivors, pad to a power of two, bitonic sort descending ---------------------
    int P2 = 1;
    while (P2 < npoint) P2 <<= 1;
    if (t == 0) s_fill = 0;
    for (int i = t; i < P2; i += kThreads) skeys[i] = 0ull;  // 0 sorts last; real keys have a non-zero low word
    __syncthreads();
    for (int k = t; k < n; k += kThreads) {
        const unsigned long long key = make_key(cls, k, c);
        if (key >= threshold) skeys[atomicAdd(&s_fill, 1)] = key;
    }
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
    for (int i = t; i < npoint; i += kThreads) idx[i] = (int)(~(unsigned)skeys[i]);
}

}  // namespace

extern "C" int pdab_topk_ctr(int b, int n, int c, int npoint, const float *cls, int *idx, pdab_stream_t stream) {
    if (b < 0 || n < 1 || c < 1 || npoint < 0 || npoint > n || !cls || !idx) return PDAB_EINVAL;
    if (b == 0 || npoint == 0) return 0;
    if (npoint > kMaxSelect) return PDAB_EUNSUPPORTED;
    int P2 = 1;
    while (P2 < npoint) P2 <<= 1;
    const size_t smem = (size_t)P2 * sizeof(unsigned long long);
    // per device and cheap: set on every launch (a cached flag would skip the second GPU of a process)
    PDAB_CUDA(cudaFuncSetAttribute(topk_ctr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(kMaxSelect * sizeof(unsigned long long))));
    topk_ctr_kernel<<<b, kThreads, smem, pdab::to_stream(stream)>>>(n, c, npoint, cls, idx);
    PDAB_LAUNCH_CHECK();
    return 0;
}
