// Micro-benchmarks behind DESIGN.md's SpTRSV analysis: dependent fp64 op latency and the cost of one
// hop of a cross-SM dependency chain (store to L2 -> polled load on another SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/latency tools/micro/latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chain_dadd(double *out, double a, int n, long long *cycles) {
    double s = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) s = __dadd_rn(s, a);
    long long t1 = clock64();
    out[0] = s;
    cycles[0] = t1 - t0;
}
__global__ void chain_dfma(double *out, double a, int n, long long *cycles) {
    double s = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) s = fma(s, a, a);
    long long t1 = clock64();
    out[0] = s;
    cycles[0] = t1 - t0;
}
__global__ void chain_ddiv(double *out, double a, int n, long long *cycles) {
    double s = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) s = __ddiv_rn(s, a);
    long long t1 = clock64();
    out[0] = s;
    cycles[0] = t1 - t0;
}
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// block 0 and block 1 (different SMs) bounce a counter through two L2 words
__global__ void pingpong(unsigned long long *w, int n, long long *cycles) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    long long t0 = clock64();
    for (int i = 1; i <= n; ++i) {
        if (me == 0) {
            __stcg(&w[0], (unsigned long long)i);
            while (ld_relaxed(&w[16]) != (unsigned long long)i) {}
        } else {
            while (ld_relaxed(&w[0]) != (unsigned long long)i) {}
            __stcg(&w[16], (unsigned long long)i);
        }
    }
    long long t1 = clock64();
    if (me == 0) cycles[0] = t1 - t0;
}
// a ring of `nb` blocks passes a token: hop = poll + store
__global__ void ring(unsigned long long *w, int laps, long long *cycles) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x, nb = gridDim.x;
    long long t0 = clock64();
    for (int lap = 0; lap < laps; ++lap) {
        const unsigned long long want = (unsigned long long)lap * nb + me;   // token value I wait for
        if (!(lap == 0 && me == 0))
            while (ld_relaxed(&w[me * 16]) != want) {}
        __stcg(&w[((me + 1) % nb) * 16], want + 1);
    }
    long long t1 = clock64();
    if (me == 0) cycles[0] = t1 - t0;
}

// how long does __nanosleep(t) really take?
__global__ void sleep_cost(unsigned int ns, int n, long long *cycles) {
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) __nanosleep(ns);
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}
// ping-pong with a sleep between two polls (the shape of the triangular solve's wait loop)
__global__ void pingpong_sleep(unsigned long long *w, int n, unsigned int ns, long long *cycles) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    long long t0 = clock64();
    for (int i = 1; i <= n; ++i) {
        if (me == 0) {
            __stcg(&w[0], (unsigned long long)i);
            while (ld_relaxed(&w[16]) != (unsigned long long)i) __nanosleep(ns);
        } else {
            while (ld_relaxed(&w[0]) != (unsigned long long)i) __nanosleep(ns);
            __stcg(&w[16], (unsigned long long)i);
        }
    }
    long long t1 = clock64();
    if (me == 0) cycles[0] = t1 - t0;
}

int main() {
    double *d;
    long long *c, h;
    unsigned long long *w;
    cudaMalloc(&d, 64);
    cudaMalloc(&c, 64);
    cudaMalloc(&w, 148 * 16 * 8 + 1024);
    cudaMemset(d, 0, 64);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int n = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        chain_dadd<<<1, 1>>>(d, 1e-9, n, c);
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("dependent DADD : %.1f cycles/op\n", (double)h / n);
        chain_dfma<<<1, 1>>>(d, 0.999, n, c);
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("dependent DFMA : %.1f cycles/op\n", (double)h / n);
        chain_ddiv<<<1, 1>>>(d, 1.0000001, n, c);
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("dependent DDIV : %.1f cycles/op\n", (double)h / n);
        cudaMemset(w, 0, 148 * 16 * 8);
        pingpong<<<2, 32>>>(w, 2000, c);
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("2-SM ping-pong : %.0f cycles per one-way hop (%.2f us at %d MHz)\n", (double)h / 4000, (double)h / 4000 / (clk / 1e3), clk / 1000);
        for (unsigned int ns : {0u, 1u, 20u, 100u, 400u, 1000u, 4000u}) {
            sleep_cost<<<1, 32>>>(ns, 1000, c);
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            if (rep) printf("__nanosleep(%4u): %.0f cycles (%.0f ns) per call\n", ns, (double)h / 1000, (double)h / 1000 / (clk / 1e6));
        }
        for (unsigned int ns : {1u, 20u, 100u, 400u}) {
            cudaMemset(w, 0, 148 * 16 * 8);
            pingpong_sleep<<<2, 32>>>(w, 2000, ns, c);
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            if (rep) printf("ping-pong, __nanosleep(%u) between polls: %.0f cycles per one-way hop\n", ns, (double)h / 4000);
        }
        for (int nb : {8, 64, 148}) {
            cudaMemset(w, 0, 148 * 16 * 8);
            ring<<<nb, 32>>>(w, 50, c);
            cudaError_t e = cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            if (rep) printf("ring of %3d SMs: %.0f cycles per hop (%s)\n", nb, (double)h / (50.0 * nb), cudaGetErrorString(e));
        }
    }
    return 0;
}
