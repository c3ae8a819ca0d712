// Shared helpers for libpdab.so kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pdab.h"

#define PDAB_LAUNCH_CHECK()                          \
    do {                                             \
        cudaError_t e__ = cudaGetLastError();        \
        if (e__ != cudaSuccess) return (int)e__;     \
    } while (0)

#define PDAB_CUDA(call)                              \
    do {                                             \
        cudaError_t e__ = (call);                    \
        if (e__ != cudaSuccess) return (int)e__;     \
    } while (0)

namespace pdab {

constexpr int kNumSMs = 148;  // B200

__host__ __device__ static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline cudaStream_t to_stream(pdab_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Grid cap of the persistent tensor-core kernels for the calling thread: pdab_set_persistent_ctas, or every SM of the
// current device (defined in tc_gemm.cu).
int persistent_ctas();

// Squared distance in the op order nvcc emits for the reference expression
// (a-b)^2 summed x,y,z: t = rn(dy*dy); t = fma(dx,dx,t); d = fma(dz,dz,t)
// (PB/src/sampling_gpu.cu:132, PB/src/ball_query_gpu.cu:34; checked in SASS).
// Written with explicit intrinsics so no compiler version can reorder it.
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    float t = __fmul_rn(dy, dy);
    t = __fmaf_rn(dx, dx, t);
    return __fmaf_rn(dz, dz, t);
}

// ---- packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2): each half is the scalar IEEE operation (round to nearest, no flush), so
// a packed expression gives the bits of its scalar form; what it saves is issue slots.  A pair lives in a 64-bit register.
using f32x2 = unsigned long long;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &a, float &b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

}  // namespace pdab
