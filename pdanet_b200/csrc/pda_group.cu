// PDA grouper for sm_100a: ball query + grouping + distribution-aware encoding
// (Gaussian density and direction vectors) in one kernel.
//
// The reference does this in Python (PB/pointnet2_utils.py:567-614): a ball-query
// launch, two group_points launches and ~10 elementwise / permute / cat kernels
// that make three extra passes over the grouped tensor.  Here a CTA owns 128
// query centres: each thread scans the cloud for its centre (hit list kept in
// shared memory), then the CTA streams the (7+C, 128, nsample) output slab
// channel by channel with fully coalesced stores.  The output tensor is written
// exactly once; the index tensor is optional.
//
// Output channel order [xyz(3, not centred), density, direction(3), features(C)]
// (PB/pointnet2_utils.py:607).  This stays on CUDA cores by design.
#include "ball_scan.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kStride = kThreads + 1;
constexpr int kChunk = 8;  // output elements per thread held in registers

__global__ void __launch_bounds__(kThreads)
pda_group_kernel(int c, int n, int m, float radius, float r2, float two_r2, float dens_norm, int nsample,
                 const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                 const float *__restrict__ features, float *__restrict__ out, int *__restrict__ idx_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *tile = reinterpret_cast<float4 *>(smem_raw);
    float *sctr = reinterpret_cast<float *>(tile + pdab::kScanTile + 8);  // 3 * kThreads
    int *sidx = reinterpret_cast<int *>(sctr + 3 * kThreads);          // nsample * kStride

    const int scene = blockIdx.y;
    const int t = threadIdx.x;
    const int j0 = blockIdx.x * kThreads;
    const int j = j0 + t;
    const bool active = j < m;
    xyz += (size_t)scene * n * 3;
    features += (size_t)scene * c * n;

    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (active) {
        const float *ctr = new_xyz + ((size_t)scene * m + j) * 3;
        cx = ctr[0];
        cy = ctr[1];
        cz = ctr[2];
    }
    sctr[t * 3 + 0] = cx;
    sctr[t * 3 + 1] = cy;
    sctr[t * 3 + 2] = cz;
    pdab::ball_scan_to_smem<kThreads, kStride>(n, xyz, active, cx, cy, cz, r2, nsample, tile, sidx);

    const int nctr = min(kThreads, m - j0);
    const int E = nctr * nsample;  // elements of this CTA's slab per channel
    const size_t chan_stride = (size_t)m * nsample;
    float *obase = out + (size_t)scene * (7 + c) * chan_stride + (size_t)j0 * nsample;
    int *ibase = idx_out ? idx_out + ((size_t)scene * m + j0) * nsample : nullptr;

    for (int e0 = 0; e0 < E; e0 += kThreads * kChunk) {
        int k[kChunk];
        int e[kChunk];
#pragma unroll
        for (int q = 0; q < kChunk; q++) {
            e[q] = e0 + q * kThreads + t;
            if (e[q] < E) {
                const int jl = e[q] / nsample, s = e[q] - jl * nsample;
                k[q] = sidx[s * kStride + jl];
                if (ibase) ibase[e[q]] = k[q];
                const float gx = __ldg(xyz + (size_t)k[q] * 3 + 0);
                const float gy = __ldg(xyz + (size_t)k[q] * 3 + 1);
                const float gz = __ldg(xyz + (size_t)k[q] * 3 + 2);
                const float dx = gx - sctr[jl * 3 + 0], dy = gy - sctr[jl * 3 + 1], dz = gz - sctr[jl * 3 + 2];
                // torch.norm(...) then **2 (PB/pointnet2_utils.py:592-593)
                const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
                const float dens = expf(-(dist * dist) / two_r2) / dens_norm;
                float *o = obase + e[q];
                __stcs(o + 0 * chan_stride, gx);
                __stcs(o + 1 * chan_stride, gy);
                __stcs(o + 2 * chan_stride, gz);
                __stcs(o + 3 * chan_stride, dens);
                __stcs(o + 4 * chan_stride, dx / radius);
                __stcs(o + 5 * chan_stride, dy / radius);
                __stcs(o + 6 * chan_stride, dz / radius);
            } else {
                k[q] = -1;
            }
        }
        for (int ch = 0; ch < c; ch++) {
            const float *frow = features + (size_t)ch * n;
            float v[kChunk];
#pragma unroll
            for (int q = 0; q < kChunk; q++)
                if (k[q] >= 0) v[q] = __ldg(frow + k[q]);
            float *o = obase + (size_t)(7 + ch) * chan_stride;
#pragma unroll
            for (int q = 0; q < kChunk; q++)
                if (k[q] >= 0) __stcs(o + e[q], v[q]);
        }
    }
}

// ---- token-major variant --------------------------------------------------------------------------
// Same search and encoding, but the output is laid out for the consumers that follow in the PDA block
// (per-token MLPs and a transformer over each neighbourhood): out (B, M, nsample, pitch) with one
// contiguous row per (centre, neighbour) token:
//   [0..2] neighbour xyz (not centred), [3] Gaussian density, [4..6] direction, [7] 0, [8..8+C) features
// and the features are read from a POINT-major copy (B, N, C), so a token is one coalesced read of its
// neighbour's feature row and one coalesced, 16-byte aligned write (pitch % 4 == 0).  A warp moves one
// token at a time; the CTA's whole output slab is one contiguous range.  HBM traffic is the output
// (4 * pitch bytes per token) plus L2-resident gathers.
constexpr int kTokThreads = 512;  // 128 scan threads + helpers: the sweep phase wants many loads in flight

__global__ void __launch_bounds__(kTokThreads)
pda_group_tokens_kernel(int c, int n, int m, int pitch, float radius, float r2, float two_r2, float dens_norm,
                        int nsample, const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                        const float *__restrict__ features_t, float *__restrict__ out, int *__restrict__ idx_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *tile = reinterpret_cast<float4 *>(smem_raw);
    float *sctr = reinterpret_cast<float *>(tile + pdab::kScanTile + 8);
    int *sidx = reinterpret_cast<int *>(sctr + 3 * kThreads);

    const int scene = blockIdx.y;
    const int t = threadIdx.x;
    const int j0 = blockIdx.x * kThreads;
    const int j = j0 + t;
    const bool active = t < kThreads && j < m;
    xyz += (size_t)scene * n * 3;
    features_t += (size_t)scene * n * c;

    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (active) {
        const float *ctr = new_xyz + ((size_t)scene * m + j) * 3;
        cx = ctr[0];
        cy = ctr[1];
        cz = ctr[2];
    }
    if (t < kThreads) {
        sctr[t * 3 + 0] = cx;
        sctr[t * 3 + 1] = cy;
        sctr[t * 3 + 2] = cz;
    }
    pdab::ball_scan_to_smem<kThreads, kStride>(n, xyz, active, cx, cy, cz, r2, nsample, tile, sidx);

    const int nctr = min(kThreads, m - j0);
    const int ntok = nctr * nsample;
    float *obase = out + ((size_t)scene * m + j0) * nsample * pitch;
    int *ibase = idx_out ? idx_out + ((size_t)scene * m + j0) * nsample : nullptr;
    const int p4 = pitch >> 2;
    const int c4 = c >> 2;
    float4 *o4 = reinterpret_cast<float4 *>(obase);

    // (a) geometry: one thread per token writes the two leading float4 of its row
    for (int tok = t; tok < ntok; tok += kTokThreads) {
        const int jl = tok / nsample, s = tok - jl * nsample;
        const int k = sidx[s * kStride + jl];
        const float gx = __ldg(xyz + (size_t)k * 3 + 0), gy = __ldg(xyz + (size_t)k * 3 + 1),
                    gz = __ldg(xyz + (size_t)k * 3 + 2);
        const float dx = gx - sctr[jl * 3 + 0], dy = gy - sctr[jl * 3 + 1], dz = gz - sctr[jl * 3 + 2];
        const float dist = sqrtf(dx * dx + dy * dy + dz * dz);  // torch.norm(...)**2, PB/pointnet2_utils.py:592-593
        __stcs(o4 + (size_t)tok * p4, make_float4(gx, gy, gz, expf(-(dist * dist) / two_r2) / dens_norm));
        __stcs(o4 + (size_t)tok * p4 + 1, make_float4(dx / radius, dy / radius, dz / radius, 0.f));
        if (ibase) ibase[tok] = k;
    }
    // (b) features: flat sweep over (token, float4-of-row); lanes of a warp read consecutive 16-byte pieces of
    // one neighbour's feature row and write consecutive pieces of the output row; kChunk gathers in flight / thread
    const int f4 = p4 - 2;  // feature float4 per row incl. padding beyond c
    const int total = ntok * f4;
    for (int e0 = 0; e0 < total; e0 += kTokThreads * kChunk) {
        float4 v[kChunk];
#pragma unroll
        for (int u = 0; u < kChunk; u++) {
            const int e = e0 + u * kTokThreads + t;
            if (e < total) {
                const int tok = e / f4, q = e - tok * f4;
                const int jl = tok / nsample, s = tok - jl * nsample;
                const int k = sidx[s * kStride + jl];
                v[u] = q < c4 ? __ldg(reinterpret_cast<const float4 *>(features_t + (size_t)k * c) + q)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < kChunk; u++) {
            const int e = e0 + u * kTokThreads + t;
            if (e < total) {
                const int tok = e / f4, q = e - tok * f4;
                __stcs(o4 + (size_t)tok * p4 + 2 + q, v[u]);
            }
        }
    }
}

}  // namespace

extern "C" int pdab_pda_group_tokens(int b, int c, int n, int m, float radius, int nsample, int pitch,
                                     const float *xyz, const float *new_xyz, const float *features_t, float *out,
                                     int *idx_out, pdab_stream_t stream) {
    if (b < 0 || c < 0 || n < 1 || m < 0 || nsample < 1 || !xyz || !new_xyz || !out || (c > 0 && !features_t))
        return PDAB_EINVAL;
    if (pitch < 8 + c || (pitch & 3) || (c & 3)) return PDAB_EINVAL;
    if (b == 0 || m == 0) return 0;
    if (nsample > 128 || b > 65535) return PDAB_EUNSUPPORTED;
    const size_t smem = sizeof(float4) * (pdab::kScanTile + 8) + sizeof(float) * 3 * kThreads +
                        sizeof(int) * (size_t)nsample * kStride;
    PDAB_CUDA(cudaFuncSetAttribute(pda_group_tokens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float two_r2 = (float)(2.0 * (double)radius * (double)radius);
    const float dens_norm = (float)(2.5 * (double)radius);
    dim3 grid(pdab::div_up(m, kThreads), b);
    pda_group_tokens_kernel<<<grid, kTokThreads, smem, pdab::to_stream(stream)>>>(
        c, n, m, pitch, radius, radius * radius, two_r2, dens_norm, nsample, xyz, new_xyz, features_t, out, idx_out);
    PDAB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pdab_pda_group(int b, int c, int n, int m, float radius, int nsample, const float *xyz,
                              const float *new_xyz, const float *features, float *out, int *idx_out,
                              pdab_stream_t stream) {
    if (b < 0 || c < 0 || n < 1 || m < 0 || nsample < 1 || !xyz || !new_xyz || !out || (c > 0 && !features))
        return PDAB_EINVAL;
    if (b == 0 || m == 0) return 0;
    if (nsample > 128 || b > 65535) return PDAB_EUNSUPPORTED;
    const size_t smem = sizeof(float4) * (pdab::kScanTile + 8) + sizeof(float) * 3 * kThreads +
                        sizeof(int) * (size_t)nsample * kStride;
    // per device and cheap: set on every launch (a cached flag would skip the second GPU of a process)
    PDAB_CUDA(cudaFuncSetAttribute(pda_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // Python-double scalars are rounded to fp32 once, as torch does for tensor-scalar ops
    // (PB/pointnet2_utils.py:593).
    const float two_r2 = (float)(2.0 * (double)radius * (double)radius);
    const float dens_norm = (float)(2.5 * (double)radius);
    dim3 grid(pdab::div_up(m, kThreads), b);
    pda_group_kernel<<<grid, kThreads, smem, pdab::to_stream(stream)>>>(c, n, m, radius, radius * radius, two_r2,
                                                                         dens_norm, nsample, xyz, new_xyz, features,
                                                                         out, idx_out);
    PDAB_LAUNCH_CHECK();
    return 0;
}
