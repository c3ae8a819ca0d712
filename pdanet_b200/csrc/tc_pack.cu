// Weight packing for the tcgen05 kernels (tc_gemm.cu): W (nout, k) row-major -> the exact shared-memory image of
// every (column chunk, k-atom) tile, so that a pipeline stage's weights are one contiguous cp.async.bulk.
//
// Packed layout, in floats:  [chunk = n / bn][atom = k / 32][part: hi, (lo)][row r = n % bn][16-byte slot][4]
// where the 16-byte slot of logical k-chunk c (c = (k % 32) / 4) in row r is  c ^ (r & 7)  — the canonical K-major
// SWIZZLE_128B pattern the UMMA shared-memory descriptor expects (rows of 128 B, 8-row groups 1024 B apart).
// npass = 3: hi = top 19 bits of w (exact in TF32), lo = w - hi (exact in fp32).  npass = 1: cvt.rna.tf32 of w.
// npass = 2: bf16 pairs, k-atoms of 64: hi = bf16(w), lo = bf16(w - hi) (round to nearest even).
// npass = 4: fp16 pairs, k-atoms of 64, ONE plane: fp16(w) (round to nearest even) — the single-pass mode.
// Runs once per module (weights are static in eval mode); not on the hot path.
#include "common.cuh"

namespace {

__global__ void pack_kernel(int nout, int k, int npass, int bn, int xyz_last, const float *__restrict__ w,
                            float *__restrict__ packed, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    // decode the destination index
    const int parts = (npass == 2 || npass == 3) ? 2 : 1;
    const int bke = (npass == 2 || npass == 4) ? 64 : 32;      // elements per k-atom (one 128-byte row: 32 tf32 or 64 x 16 bit)
    const int e = (int)(i & 3);                // 32-bit word inside the 16-byte slot
    const int slot = (int)((i >> 2) & 7);
    long long rest = i >> 5;
    const int r = (int)(rest % bn);
    rest /= bn;
    const int part = (int)(rest % parts);
    rest /= parts;
    const int atoms = (k + bke - 1) / bke;
    const int atom = (int)(rest % atoms);
    const int chunk = (int)(rest / atoms);
    const int c = slot ^ (r & 7);
    const int n = chunk * bn + r;
    // A operand order [features (k - xyz_last), xyz (xyz_last)]  <-  reference order [xyz, features]
    auto weight = [&](int kk) {
        if (n >= nout || kk >= k) return 0.f;
        const int src = xyz_last > 0 ? (kk < k - xyz_last ? kk + xyz_last : kk - (k - xyz_last)) : kk;
        return w[(size_t)n * k + src];
    };
    if (npass == 4) {  // two fp16 per word
        const int kk = atom * 64 + c * 8 + e * 2;
        uint32_t h;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(weight(kk + 1)), "f"(weight(kk)));
        packed[i] = __uint_as_float(h);
        return;
    }
    if (npass == 2) {  // two bf16 per word: (hi, lo) split with round-to-nearest-even, lo = bf16(w - hi)
        const int kk = atom * 64 + c * 8 + e * 2;
        const float v0 = weight(kk), v1 = weight(kk + 1);
        uint32_t h, l;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(v1), "f"(v0));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(v1 - __uint_as_float(h & 0xffff0000u)),
            "f"(v0 - __uint_as_float(h << 16)));
        packed[i] = __uint_as_float(part == 0 ? h : l);
        return;
    }
    const float v = weight(atom * 32 + c * 4 + e);  // column of the (possibly permuted) A operand
    float o;
    if (npass == 3) {
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        o = part == 0 ? hi : v - hi;
    } else {
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
        o = __uint_as_float(t);
    }
    packed[i] = o;
}

}  // namespace

extern "C" size_t pdab_tc_packed_floats(int nout, int k, int npass, int bn) {
    if (nout < 1 || k < 1 || bn < 1) return 0;
    const size_t bke = (npass == 2 || npass == 4) ? 64 : 32;
    const size_t chunks = (size_t)(nout + bn - 1) / bn, atoms = ((size_t)k + bke - 1) / bke;
    return chunks * atoms * ((npass == 2 || npass == 3) ? 2 : 1) * (size_t)bn * 32;
}

extern "C" int pdab_tc_pack_weights(int nout, int k, int npass, int bn, int xyz_last, const float *w, float *packed,
                                    pdab_stream_t stream) {
    if (nout < 1 || k < 1 || !w || !packed || xyz_last < 0 || xyz_last > k) return PDAB_EINVAL;
    if (npass < 1 || npass > 4 || (bn != 128 && bn != 192 && bn != 256)) return PDAB_EINVAL;
    const long long total = (long long)pdab_tc_packed_floats(nout, k, npass, bn);
    const long long blocks = (total + 255) / 256;
    pack_kernel<<<(unsigned)blocks, 256, 0, pdab::to_stream(stream)>>>(nout, k, npass, bn, xyz_last, w, packed, total);
    PDAB_LAUNCH_CHECK();
    return 0;
}
