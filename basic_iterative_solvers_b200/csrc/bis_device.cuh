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

// Fixed combination of the records of a reduction: 8 records as a balanced tree, otherwise in order.
__device__ __forceinline__ double combine_records(const double *r, int n) {
    if (n == BIS_NSLAB) return ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    double t = 0.0;
    for (int i = 0; i < n; ++i) t += r[i];
    return t;
}

// Block-wide sum of NRED accumulators (fixed shape: shuffle tree per warp, then a shuffle tree over the
// warp sums), then the grid-wide deterministic finish described in bis_internal.cuh.  Must be called by
// ALL threads of the block (blockDim.x <= 1024, multiple of 32).  `part_index`: where this block's
// partial goes (block_offset + blockIdx.x for plain launches).
template <int NRED> __device__ __forceinline__ void grid_reduce_finish(const RedArgs &ra);

template <int NRED>
__device__ __forceinline__ void block_reduce_finish(double (&acc)[NRED], const RedArgs &ra, int part_index = -1) {
    __shared__ double s_part[NRED][32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    if (part_index < 0) part_index = ra.block_offset + (int)blockIdx.x;
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
            if (lane == 0) ra.partials[q * BIS_MAX_RED_BLOCKS + part_index] = v;
        }
    }
    grid_reduce_finish<NRED>(ra);
}

// The grid-wide part: called by ALL threads of every block after thread 0 of the block has stored the
// block's partial(s).
template <int NRED>
__device__ __forceinline__ void grid_reduce_finish(const RedArgs &ra) {
    __shared__ double s_rec[NRED][BIS_NSLAB];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    if (!ra.finalize) return;
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int ticket = atomicAdd(ra.counter, 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block: one warp per local slab adds that slab's partials in a fixed order that depends on
    // nothing but the slab (lane l: l, l+32, ...; 8 independent loads in flight; then the shuffle tree)
    for (int i = warp; i < ra.n_slab; i += nwarp) {
        const int lo = ra.slab_off[i], hi = ra.slab_off[i + 1];
#pragma unroll
        for (int q = 0; q < NRED; ++q) {
            const double *p = ra.partials + q * BIS_MAX_RED_BLOCKS;
            double v = 0.0;
            int j = lo + lane;
            for (; j + 7 * 32 < hi; j += 8 * 32) {
                double t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = __ldcg(p + j + u * 32);
#pragma unroll
                for (int u = 0; u < 8; ++u) v += t[u];
            }
            for (; j < hi; j += 32) v += __ldcg(p + j);
            v = warp_sum(v);
            if (lane == 0) s_rec[q][i] = v;
        }
    }
    __syncthreads();
    if (warp != 0) return;
    double t0 = 0.0, t1 = 0.0;   // lane i: record i
    if (ra.peer_n > 1) {
        // Sum over ranks through peer memory.  The lanes of warp 0 carry (destination rank, local slab)
        // pairs: this rank's slab sums go into every rank's bank (NVLink stores, then the epoch as the ready
        // flag), then lane i waits for record i in the local bank.  Records alternate between two copies by
        // epoch parity: a rank can only be one reduction ahead of the slowest one, because finishing a
        // reduction needs everybody's records of it.  All ranks add the same records in the same order.
        const int par = (int)(ra.peer_epoch & 1ull);
        unsigned long long waited = 0ull;
        if (lane < ra.peer_n * ra.n_slab) {
            const int dst = lane / ra.n_slab, i = lane - dst * ra.n_slab;
            volatile double *rec = ra.peer_bank[dst] + (par * BIS_NSLAB + ra.rec_first + i) * 4;
            rec[0] = s_rec[0][i];
            rec[1] = NRED > 1 ? s_rec[NRED > 1 ? 1 : 0][i] : 0.0;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(rec + 2) = ra.peer_epoch;
        }
        if (lane < ra.n_rec) {
            volatile double *in = ra.peer_bank[ra.peer_rank] + (par * BIS_NSLAB + lane) * 4;
            const unsigned long long t_start = bis_globaltimer();
            while (*reinterpret_cast<volatile unsigned long long *>(in + 2) != ra.peer_epoch) {
                if (bis_globaltimer() - t_start > BIS_PEER_TIMEOUT_NS) {
                    atomicExch(ra.errflag, 20 + lane);
                    break;
                }
            }
            waited = bis_globaltimer() - t_start;
            __threadfence_system();
            t0 = in[0];
            t1 = in[1];
        }
        // how long this rank sat waiting for the slowest one: the number that names the scaling limiter
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, waited, o);
            waited = other > waited ? other : waited;
        }
        if (lane == 0 && ra.waitstat) {
            ra.waitstat[0] += waited;
            ra.waitstat[1] += 1ull;
        }
    } else if (lane < ra.n_rec) {
        t0 = s_rec[0][lane];
        t1 = NRED > 1 ? s_rec[NRED > 1 ? 1 : 0][lane] : 0.0;
    }
    double r0[BIS_NSLAB], r1[BIS_NSLAB];
#pragma unroll
    for (int i = 0; i < BIS_NSLAB; ++i) {
        r0[i] = __shfl_sync(0xffffffffu, t0, i);
        r1[i] = __shfl_sync(0xffffffffu, t1, i);
    }
    if (lane == 0) {
        if (ra.slot[0] >= 0) ra.scalars[ra.slot[0]] = combine_records(r0, ra.n_rec);
        if (NRED > 1 && ra.slot[NRED > 1 ? 1 : 0] >= 0) ra.scalars[ra.slot[NRED > 1 ? 1 : 0]] = combine_records(r1, ra.n_rec);
        *ra.counter = 0u;
    }
}
