// Latency micro-benchmark for the building blocks of a stencil-wavefront step (one warp, clock64 deltas).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/micro/dplat tools/micro/dplat.cu && tools/micro/dplat
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(double *out, long long *cyc, double seed, int n) {
    __shared__ double sm[8 * 64];
    const int lane = threadIdx.x;
    for (int i = lane; i < 8 * 64; i += 32) sm[i] = seed + i;
    __syncwarp();
    double a = seed + lane, b = seed * 0.5 + 1.0;
    long long t0, t1;
    // (0) dependent DADD chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a = __dadd_rn(a, b);
    }
    t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    // (1) dependent DMUL chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a = __dmul_rn(a, 1.0000001);
    }
    t1 = clock64();
    if (lane == 0) cyc[1] = t1 - t0;
    // (2) dependent divisions
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) a = __ddiv_rn(a, b) + 3.0;
    }
    t1 = clock64();
    if (lane == 0) cyc[2] = t1 - t0;
    // (3) STS -> __syncwarp -> LDS of the neighbour's value -> DADD, dependent (the ring hop)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            sm[(u & 7) * 64 + lane + 1] = a;
            __syncwarp();
            a = __dadd_rn(a, sm[(u & 7) * 64 + lane]);
        }
    }
    t1 = clock64();
    if (lane == 0) cyc[3] = t1 - t0;
    // (4) 26 independent LDS.64 + 13 mul + 13 dependent adds (a row's worth)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < 13; ++u) s = __dadd_rn(s, __dmul_rn(sm[u * 32 + lane], sm[(u & 7) * 64 + lane + 1]));
        a = __dadd_rn(a, s);
        sm[lane] = a;
        __syncwarp();
    }
    t1 = clock64();
    if (lane == 0) cyc[4] = t1 - t0;
    // (5) __any_sync in a dependent loop
    t0 = clock64();
    int p = lane;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) p += __any_sync(0xffffffffu, p == 12345 + u) ? 1 : 2;
    }
    t1 = clock64();
    if (lane == 0) cyc[5] = t1 - t0;
    // (6) cp.async commit + wait_group 6 (nothing outstanding)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 6;" ::: "memory");
        }
    }
    t1 = clock64();
    if (lane == 0) cyc[6] = t1 - t0;
    // (7) dependent DFMA chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a = fma(a, 1.0000001, b);
    }
    t1 = clock64();
    if (lane == 0) cyc[7] = t1 - t0;
    // (8) independent DADDs (throughput, 8 chains)
    double c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) c[u] = a + u;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int u = 0; u < 8; ++u) c[u] = __dadd_rn(c[u], b);
    }
    t1 = clock64();
    if (lane == 0) cyc[8] = t1 - t0;
#pragma unroll
    for (int u = 0; u < 8; ++u) a += c[u];
    out[lane] = a + p;
}

int main() {
    double *out;
    long long *cyc, h[16];
    cudaMalloc(&out, 32 * 8);
    cudaMalloc(&cyc, 16 * 8);
    const int n = 1000;
    for (int rep = 0; rep < 2; ++rep) {
        k<<<1, 32>>>(out, cyc, 1.0 + rep, n);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const char *names[] = {"dependent DADD", "dependent DMUL", "dependent (DDIV + DADD)", "STS -> syncwarp -> LDS -> DADD (ring hop)",
                           "row: 26 LDS + 13 DMUL + 13 dependent DADD + STS + syncwarp", "__any_sync (dependent)",
                           "cp.async commit + wait_group", "dependent DFMA", "independent DADD (8 chains), per instruction"};
    const int per[] = {16, 16, 4, 4, 1, 4, 4, 16, 16};
    for (int i = 0; i < 9; ++i) printf("%-62s %8.1f cycles\n", names[i], (double)h[i] / n / per[i]);
    printf("cudaError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
