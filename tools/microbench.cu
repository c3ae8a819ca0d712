// Latency micro-benchmarks for the primitives on FPS's serial chain (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu && gpurun_out/microbench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int R = 256;

__global__ void k_redux(unsigned *out, long long *cyc) {
    unsigned v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) v = __reduce_max_sync(0xffffffffu, v + (threadIdx.x & 1));
    long long t1 = clock64();
    if (threadIdx.x == 0) { *cyc = t1 - t0; *out = v; }
}
__global__ void k_shfl(unsigned *out, long long *cyc) {
    unsigned v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1;
    long long t1 = clock64();
    if (threadIdx.x == 0) { *cyc = t1 - t0; *out = v; }
}
__global__ void k_ballot(unsigned *out, long long *cyc) {
    unsigned v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) v = __ballot_sync(0xffffffffu, (v >> (threadIdx.x & 31)) & 1) ^ threadIdx.x;
    long long t1 = clock64();
    if (threadIdx.x == 0) { *cyc = t1 - t0; *out = v; }
}
__global__ void k_lds(unsigned *out, long long *cyc) {
    __shared__ unsigned s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 37 + 11) & 1023;
    __syncthreads();
    unsigned v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) v = s[v];
    long long t1 = clock64();
    if (threadIdx.x == 0) { *cyc = t1 - t0; *out = v; }
}
__global__ void k_ffma(float *out, long long *cyc) {
    float v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) v = fmaf(v, 1.0001f, 0.5f);
    long long t1 = clock64();
    if (threadIdx.x == 0) { *cyc = t1 - t0; *out = v; }
}
__global__ void k_fmnmx(float *out, long long *cyc) {
    float v = threadIdx.x, w = 3.f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) { v = fminf(v, w); w = fmaxf(w, v) + 0.f; }
    long long t1 = clock64();
    if (threadIdx.x == 0) { *cyc = t1 - t0; *out = v + w; }
}
__global__ void k_bar(unsigned *out, long long *cyc) {
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) { *cyc = t1 - t0; *out = 0; }
}
// smem publish + barrier + read-back: the cross-warp exchange step of the argmax
__global__ void k_exchange(unsigned *out, long long *cyc) {
    __shared__ unsigned long long red[2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long v = threadIdx.x;
    int buf = 0;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < R; i++) {
        if (lane == 0) red[buf][warp] = v;
        __syncthreads();
        v = red[buf][lane & (blockDim.x / 32 - 1)] + 1;
        buf ^= 1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { *cyc = t1 - t0; *out = (unsigned)v; }
}

template <typename F>
void run(const char *name, F launch) {
    long long *cyc; unsigned *out;
    cudaMalloc(&cyc, 8); cudaMalloc(&out, 8);
    launch(out, cyc); launch(out, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %7.1f cycles/op\n", name, (double)h / R);
    cudaFree(cyc); cudaFree(out);
}

int main() {
    run("REDUX.max (dependent)", [](unsigned *o, long long *c) { k_redux<<<1, 32>>>(o, c); });
    run("SHFL.xor+add (dependent)", [](unsigned *o, long long *c) { k_shfl<<<1, 32>>>(o, c); });
    run("VOTE.ballot (dependent)", [](unsigned *o, long long *c) { k_ballot<<<1, 32>>>(o, c); });
    run("LDS (pointer chase)", [](unsigned *o, long long *c) { k_lds<<<1, 32>>>(o, c); });
    run("FFMA (dependent)", [](unsigned *o, long long *c) { k_ffma<<<1, 32>>>((float *)o, c); });
    run("FMNMX x2 + FADD (dependent)", [](unsigned *o, long long *c) { k_fmnmx<<<1, 32>>>((float *)o, c); });
    for (int w : {1, 4, 8, 16, 32}) {
        char nm[64];
        snprintf(nm, 64, "bar.sync, %d warps", w);
        run(nm, [w](unsigned *o, long long *c) { k_bar<<<1, 32 * w>>>(o, c); });
        snprintf(nm, 64, "STS+bar+LDS exchange, %d warps", w);
        run(nm, [w](unsigned *o, long long *c) { k_exchange<<<1, 32 * w>>>(o, c); });
    }
    return 0;
}
