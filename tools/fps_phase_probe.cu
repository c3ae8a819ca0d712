// Developer probe: where does a round of the pruned FPS chain spend its cycles?  Builds the kernel of csrc/fps_pruned.cu with
// PDAB_FPS_TIMERS (clock deltas of warp 0 of scene 0, accumulated per phase; round statistics) and prints cycles per round.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pdanet_b200/csrc -I include tools/fps_phase_probe.cu -o /tmp/fps_probe
//   add -DPDAB_FPS_TIMERS for the phase breakdown (the timers cost about 15 % themselves); without it only the event time.
//   usage: fps_probe B N m CL variant   (variant: see the switch in pdab::fps_pruned; 0 = the library's choice)
#define PDAB_FPS_PROBE 1
#include "../pdanet_b200/csrc/fps_pruned.cu"

#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char **argv) {
    const int b = argc > 1 ? atoi(argv[1]) : 16, n = argc > 2 ? atoi(argv[2]) : 16384, m = argc > 3 ? atoi(argv[3]) : 4096;
    const int cl = argc > 4 ? atoi(argv[4]) : 1;
    pdab::g_probe_variant = argc > 5 ? atoi(argv[5]) : 0;
    std::vector<float> h((size_t)b * n * 3);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.0f / 16777216.0f); };
    for (size_t i = 0; i < h.size(); i += 3) {   // a LiDAR-like slab: 70 x 80 x 4 m
        h[i] = rnd() * 70.f;
        h[i + 1] = rnd() * 80.f - 40.f;
        h[i + 2] = rnd() * 4.f - 3.f;
    }
    float *xyz, *temp;
    int *idx;
    cudaMalloc(&xyz, h.size() * 4);
    cudaMalloc(&temp, (size_t)b * n * 4);
    cudaMalloc(&idx, (size_t)b * m * 4);
    cudaMemcpy(xyz, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    std::vector<float> big((size_t)b * n, 1e10f);
    int L = 0;
    while ((1 << (L + 1)) <= n && L < 10) L++;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; rep++) {
        cudaMemcpy(temp, big.data(), big.size() * 4, cudaMemcpyHostToDevice);
#ifdef PDAB_FPS_TIMERS
        unsigned long long zero[16] = {};
        cudaMemcpyToSymbol(g_fps_phase, zero, sizeof(zero));
#endif
        cudaEventRecord(e0);
        const int rc = pdab::fps_pruned(b, n, m, xyz, temp, idx, L, cl, 0);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<int> hi((size_t)b * m);
        cudaMemcpy(hi.data(), idx, hi.size() * 4, cudaMemcpyDeviceToHost);
        unsigned long long sum = 0;
        for (size_t i = 0; i < hi.size(); i++) sum = sum * 1000003ull + (unsigned)hi[i];
        printf("idx hash %016llx ", sum);
        printf("variant %d rc=%d err=%s  %.3f ms  %.3f us/step", pdab::g_probe_variant, rc, cudaGetErrorString(err), ms,
               ms * 1e3 / (m - 1));
#ifdef PDAB_FPS_TIMERS
        unsigned long long ph[16];
        cudaMemcpyFromSymbol(ph, g_fps_phase, sizeof(ph));
        const double rounds = ph[4] > 0 ? (double)ph[4] : 1.0;
        const double merged = rounds - (double)ph[5] - (double)ph[6];
        printf(" | rounds %.0f (%.2f samples/round; fallbacks: empty %llu, overflow %llu; mean list %.1f) | cycles/round (warp 0): "
               "apply %.0f  refresh+push %.0f  barrier %.0f  merge %.0f",
               rounds, (m - 1) / rounds, ph[5], ph[6], merged > 0 ? (double)ph[7] / merged : 0.0, (double)ph[0] / rounds,
               (double)ph[1] / rounds, (double)ph[2] / rounds, (double)ph[3] / rounds);
#endif
        printf("\n");
    }
    return 0;
}
