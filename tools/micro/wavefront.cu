// Micro-benchmark behind DESIGN.md's SpTRSV analysis: what does ONE level-to-level hop of a sparse
// triangular solve cost when rows wait for three values of the previous level that other SMs publish?
// G persistent warps, warp j owns positions [32j, 32j+32) of every level; the value of (level, p)
// depends on (level-1, p-1), (level-1, p), (level-1, p+33) (mod width): three producers in up to three
// other warps, like the three level-1 operands of an HPCG row.  Values live in a sentinel-filled
// array, one 32-byte sector per value (scattered, as in the solve).  Variants of the wait loop:
//   0: every lane polls its missing operands and leaves the loop on its own (divergent exit)
//   1: warp-uniform loop, the warp leaves when all its lanes are ready
//   2: like 1, with two staggered polls in flight
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/wavefront tools/micro/wavefront.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr unsigned long long SENT = 0xFFF87E5E7E5E7E5EULL;
constexpr int STRIDE = 4;   // doubles between two values (one sector each)

__device__ __forceinline__ unsigned long long ld_relaxed(const double *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int VARIANT>
__global__ void wavefront(double *w, int width, int levels, unsigned int sleep_ns, long long *cycles) {
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int p = warp * 32 + lane;
    if (p >= width) return;
    const int d0 = (p + width - 1) % width, d1 = p, d2 = (p + 33) % width;
    long long t0 = clock64();
    const long long deadline = t0 + 4000000000LL;   // ~2 s: a bug must not hang the box
    double v = 1.0;
    for (int l = 1; l <= levels; ++l) {
        const double *prev = w + (size_t)(l - 1) * width * STRIDE;
        const double *a0 = prev + (size_t)d0 * STRIDE, *a1 = prev + (size_t)d1 * STRIDE, *a2 = prev + (size_t)d2 * STRIDE;
        unsigned long long x0 = SENT, x1 = SENT, x2 = SENT;
        if (VARIANT == 0) {
            for (;;) {
                if (x0 == SENT) x0 = ld_relaxed(a0);
                if (x1 == SENT) x1 = ld_relaxed(a1);
                if (x2 == SENT) x2 = ld_relaxed(a2);
                if ((x0 != SENT && x1 != SENT && x2 != SENT) || clock64() > deadline) break;
                if (sleep_ns) __nanosleep(sleep_ns);
            }
        } else if (VARIANT == 1) {
            for (;;) {
                unsigned long long y0 = ld_relaxed(a0), y1 = ld_relaxed(a1), y2 = ld_relaxed(a2);
                const bool ok = y0 != SENT && y1 != SENT && y2 != SENT;
                if (__all_sync(0xffffffffu, ok) || clock64() > deadline) { x0 = y0; x1 = y1; x2 = y2; break; }
                if (sleep_ns) __nanosleep(sleep_ns);
            }
        } else {
            unsigned long long y0 = ld_relaxed(a0), y1 = ld_relaxed(a1), y2 = ld_relaxed(a2);
            for (;;) {
                if (sleep_ns) __nanosleep(sleep_ns);
                unsigned long long z0 = ld_relaxed(a0), z1 = ld_relaxed(a1), z2 = ld_relaxed(a2);   // second poll in flight
                bool ok = y0 != SENT && y1 != SENT && y2 != SENT;
                if (__all_sync(0xffffffffu, ok) || clock64() > deadline) { x0 = y0; x1 = y1; x2 = y2; break; }
                if (sleep_ns) __nanosleep(sleep_ns);
                y0 = ld_relaxed(a0); y1 = ld_relaxed(a1); y2 = ld_relaxed(a2);
                ok = z0 != SENT && z1 != SENT && z2 != SENT;
                if (__all_sync(0xffffffffu, ok) || clock64() > deadline) { x0 = z0; x1 = z1; x2 = z2; break; }
            }
        }
        // a row's arithmetic: a short dependent fp64 chain and one division
        double s = __longlong_as_double((long long)x0);
        s = __dadd_rn(s, __dmul_rn(0.25, __longlong_as_double((long long)x1)));
        s = __dadd_rn(s, __dmul_rn(0.25, __longlong_as_double((long long)x2)));
        v = __ddiv_rn(__dsub_rn(1.0, s * 1e-3), 1.5);
        __stcg(w + ((size_t)l * width + p) * STRIDE, v);
    }
    long long t1 = clock64();
    if (p == 0) cycles[0] = t1 - t0;
}

__global__ void fill(double *w, size_t n, size_t first_level) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        reinterpret_cast<unsigned long long *>(w)[i] = i < first_level ? 0x3FF0000000000000ULL : SENT;
}

int main(int argc, char **argv) {
    const int levels = 400;
    long long *c, h;
    cudaMalloc(&c, 64);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int width : {32 * 74, 32 * 148, 32 * 296, 32 * 592}) {
        const size_t n = (size_t)(levels + 1) * width * STRIDE;
        double *w;
        cudaMalloc(&w, n * 8);
        for (int warps_per_block : {1, 2, 8}) {
            const int blocks = (width / 32 + warps_per_block - 1) / warps_per_block;
            if (blocks < 1 || blocks > 148 * 8) continue;
            for (int variant = 0; variant < 3; ++variant)
                for (unsigned int ns : {0u, 20u, 100u}) {
                    fill<<<1024, 256>>>(w, n, (size_t)width * STRIDE);
                    if (variant == 0) wavefront<0><<<blocks, 32 * warps_per_block>>>(w, width, levels, ns, c);
                    if (variant == 1) wavefront<1><<<blocks, 32 * warps_per_block>>>(w, width, levels, ns, c);
                    if (variant == 2) wavefront<2><<<blocks, 32 * warps_per_block>>>(w, width, levels, ns, c);
                    cudaError_t e = cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
                    printf("width %5d rows (%3d warps, %d per block) variant %d sleep %3u ns: %.0f cycles = %.2f us per level (%s)\n",
                           width, width / 32, warps_per_block, variant, ns, (double)h / levels, (double)h / levels / (clk / 1e3),
                           cudaGetErrorString(e));
                    fflush(stdout);
                }
        }
        cudaFree(w);
    }
    return 0;
}
