// bis_device.cuh -- device-side helpers: pinned-rounding arithmetic and the
// deterministic block/grid reduction used by every fused kernel.
#pragma once

#include "bis_internal.cuh"

// The library is compiled with --fmad=false: a*b+c is two roundings unless
// written as fma().  The helpers below name the intent at the call sites.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// streaming loads/stores: matrix data is read exactly once per SpMV
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }

// per-row operands an SpMV epilogue fetches ahead of time (bis_spmv.cu: Epi*::load)
struct EpiPre {
    double a, b, c;
};

__device__ __forceinline__ unsigned long long bis_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// a peer that never answers turns into an error (bis_context_synchronize), not a hung GPU
constexpr unsigned long long BIS_PEER_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of NRED accumulators, then the grid-wide deterministic
// finish.  Must be called by ALL threads of the block (blockDim.x <= 1024,
// multiple of 32).
template <int NRED>
__device__ __forceinline__ void block_reduce_finish(double (&acc)[NRED], const RedArgs &ra) {
    __shared__ double s_part[NRED][32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int q = 0; q < NRED; ++q) {
        double v = warp_sum(acc[q]);
        if (lane == 0) s_part[q][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int q = 0; q < NRED; ++q) {
            double v = (lane < nwarp) ? s_part[q][lane] : 0.0;
            v = warp_sum(v);
            if (lane == 0)
                ra.partials[q * BIS_MAX_RED_BLOCKS + ra.block_offset + blockIdx.x] = v;
        }
    }
    if (!ra.finalize) return;
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int ticket = atomicAdd(ra.counter, 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block: fixed-order sum of all partials
    __shared__ double s_tot[NRED];
#pragma unroll
    for (int q = 0; q < NRED; ++q) {
        double v = 0.0;
        for (int i = threadIdx.x; i < ra.total_blocks; i += blockDim.x)
            v += __ldcg(&ra.partials[q * BIS_MAX_RED_BLOCKS + i]);
        v = warp_sum(v);
        __syncthreads();
        if (lane == 0) s_part[q][warp] = v;
        __syncthreads();
        if (warp == 0) {
            double t = (lane < nwarp) ? s_part[q][lane] : 0.0;
            t = warp_sum(t);
            if (lane == 0) s_tot[q] = t;
        }
    }
    if (warp != 0) return;
    __syncwarp();
    double t0 = s_tot[0], t1 = NRED > 1 ? s_tot[NRED > 1 ? 1 : 0] : 0.0;
    if (ra.peer_n > 1) {
        // Sum over ranks through peer memory: lane p stores this rank's partial(s) into rank p's
        // bank (NVLink store, then the epoch as the ready flag), then waits for rank p's record in
        // the local bank.  Records alternate between two copies by epoch parity: a rank can only
        // be one reduction ahead of the slowest one, because finishing a reduction needs everybody's
        // record of it.  All ranks add the P partials in rank order => identical bits everywhere.
        const int par = (int)(ra.peer_epoch & 1ull);
        if (lane < ra.peer_n) {
            volatile double *rec = ra.peer_bank[lane] + (par * BIS_MAX_PEERS + ra.peer_rank) * 4;
            rec[0] = t0;
            rec[1] = t1;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(rec + 2) = ra.peer_epoch;
            volatile double *in = ra.peer_bank[ra.peer_rank] + (par * BIS_MAX_PEERS + lane) * 4;
            const unsigned long long t_start = bis_globaltimer();
            while (*reinterpret_cast<volatile unsigned long long *>(in + 2) != ra.peer_epoch) {
                if (bis_globaltimer() - t_start > BIS_PEER_TIMEOUT_NS) {
                    atomicExch(ra.errflag, 20 + lane);
                    break;
                }
            }
            __threadfence_system();
            t0 = in[0];
            t1 = in[1];
        }
        double a0 = 0.0, a1 = 0.0;
        for (int p = 0; p < ra.peer_n; ++p) {
            a0 += __shfl_sync(0xffffffffu, t0, p);
            a1 += __shfl_sync(0xffffffffu, t1, p);
        }
        t0 = a0;
        t1 = a1;
    }
    if (lane == 0) {
        if (ra.slot[0] >= 0) ra.scalars[ra.slot[0]] = t0;
        if (NRED > 1 && ra.slot[NRED > 1 ? 1 : 0] >= 0) ra.scalars[ra.slot[NRED > 1 ? 1 : 0]] = t1;
        *ra.counter = 0u;
    }
}
