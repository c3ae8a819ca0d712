// Thread-per-centre ordered ball scan shared by the fused kernels
// (pda_group.cu, sa_fused.cu).  Same search as ball_query.cu, but the hit list
// goes to shared memory, transposed (slot-major) so that the CTA can later walk
// it without bank conflicts, and unfilled slots are completed in place.
#pragma once
#include "common.cuh"

namespace pdab {

constexpr int kScanTile = 1024;  // points per staging tile (float4, 16 KB)

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
        __syncthreads();
        for (int i = t; i < len; i += nthreads) {
            const float *p = xyz + (size_t)(base + i) * 3;
            tile[i] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
        }
        __syncthreads();
        if (!__all_sync(0xffffffffu, cnt >= nsample)) {
#pragma unroll 4
            for (int i = 0; i < len; i++) {
                const float4 p = tile[i];
                const float d2 = sqdist3(cx, cy, cz, p.x, p.y, p.z);
                if (d2 < r2 && cnt < nsample) {
                    sidx[cnt * STRIDE + t] = base + i;
                    cnt++;
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

}  // namespace pdab
