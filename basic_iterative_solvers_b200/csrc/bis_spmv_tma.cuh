// bis_spmv_tma.cuh -- SpMV variant 2: TMA-staged, one thread per row.
//
// The nonzeros of a block of consecutive rows are contiguous in CRS, so a tile
// of R rows is fetched by TWO 1-D bulk copies (cp.async.bulk, the TMA engine:
// SASS UBLKCP) -- val[s..e) and col[s..e) -- into shared memory, completion
// signalled on an mbarrier.  HBM is therefore read in large, perfectly
// sequential, fully used bursts that no thread has to wait on instruction by
// instruction; NSTAGE tiles are in flight per CTA.  Each thread then walks ITS
// row out of shared memory strictly left to right with an unfused multiply and
// add per nonzero: exactly the reference's summation order (the compiled
// reference adds the separately rounded products in storage order,
// kernels.hpp:31-36 under GCC's in-order `omp simd` reduction), so y is
// bit-identical to native_spmv.  x is gathered through L1/L2 (consecutive
// threads of a stencil matrix read consecutive x: coalesced).
//
// Grid: persistent, one CTA per SM (x occupancy), each CTA owns a CONTIGUOUS
// run of tiles (static assignment => fused dot products are bit-reproducible,
// and neighbouring tiles reuse x lines in L1).
//
// Shared memory per stage: cap * 12 bytes, cap = R * max_row + 8 (the copies
// start and end on 16-byte boundaries; the arrays are allocated with 4 spare
// elements for that).
#pragma once

#include "bis_device.cuh"

namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 1-D bulk copy global -> shared, bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

constexpr int MAX_STAGES = 8;

struct Plan {
    int threads;          // rows per tile
    int cap;              // elements per stage
    int nstage;
    int grid;
    int64_t tiles_per_cta;
    size_t smem_bytes;
};

}  // namespace tma

struct SpmvTmaIn {
    const void *rp;
    const int *col;
    const double *val;
    const double *x;
    const double *ghost;
    int64_t n_owned;
    int64_t lo, cnt;          // contiguous row range of this launch
    int64_t tiles_per_cta;
    int cap;
    int nstage;
};

template <typename RP, bool GHOST, class Epi>
__global__ void __launch_bounds__(256) spmv_tma_kernel(SpmvTmaIn in, Epi epi, RedArgs ra) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);                       // [MAX_STAGES]
    double *sval = reinterpret_cast<double *>(smem_raw + 128);                     // [nstage][cap]
    int *scol = reinterpret_cast<int *>(smem_raw + 128 + (size_t)in.nstage * in.cap * sizeof(double));

    const RP *__restrict__ rp = static_cast<const RP *>(in.rp);
    const double *__restrict__ x = in.x;
    const int R = blockDim.x;
    const int64_t n_tiles = (in.cnt + R - 1) / R;
    const int64_t t_begin = (int64_t)blockIdx.x * in.tiles_per_cta;
    int64_t t_end = t_begin + in.tiles_per_cta;
    if (t_end > n_tiles) t_end = n_tiles;
    const int64_t my_tiles = t_end > t_begin ? t_end - t_begin : 0;
    const int64_t row_end = in.lo + in.cnt;

    double acc[Epi::NRED > 0 ? Epi::NRED : 1];
#pragma unroll
    for (int q = 0; q < (Epi::NRED > 0 ? Epi::NRED : 1); ++q) acc[q] = 0.0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < in.nstage; ++s) tma::mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint64_t policy = 0;
    if (threadIdx.x == 0) policy = tma::policy_evict_first();

    // producer (thread 0): fetch tile `j` (CTA-local index) into its stage
    auto issue = [&](int64_t j) {
        const int64_t r0 = in.lo + (t_begin + j) * R;
        int64_t r1 = r0 + R;
        if (r1 > row_end) r1 = row_end;
        const int64_t s = (int64_t)rp[r0], e = (int64_t)rp[r1];
        const int64_t s_al = s & ~(int64_t)3;
        const int64_t e_al = (e + 3) & ~(int64_t)3;
        const uint32_t n_el = (uint32_t)(e_al - s_al);
        const int st = (int)(j % in.nstage);
        if (n_el == 0) {
            tma::mbar_arrive(&bars[st]);
            return;
        }
        tma::mbar_expect_tx(&bars[st], n_el * 12u);
        tma::bulk_g2s(sval + (size_t)st * in.cap, in.val + s_al, n_el * 8u, &bars[st], policy);
        tma::bulk_g2s(scol + (size_t)st * in.cap, in.col + s_al, n_el * 4u, &bars[st], policy);
    };

    if (threadIdx.x == 0) {
        const int64_t pre = my_tiles < in.nstage ? my_tiles : in.nstage;
        for (int64_t j = 0; j < pre; ++j) issue(j);
    }

    // this thread's row bounds of the current tile, loaded one tile ahead
    auto row_bounds = [&](int64_t j, int64_t &ks, int64_t &ke, int64_t &base, int64_t &row) {
        row = in.lo + (t_begin + j) * R + threadIdx.x;
        const int64_t r0 = in.lo + (t_begin + j) * R;
        if (j < my_tiles && row < row_end) {
            ks = (int64_t)rp[row];
            ke = (int64_t)rp[row + 1];
        } else {
            ks = ke = 0;
            row = -1;
        }
        base = (j < my_tiles) ? ((int64_t)rp[r0] & ~(int64_t)3) : 0;
    };
    int64_t ks, ke, base, row;
    row_bounds(0, ks, ke, base, row);

    for (int64_t j = 0; j < my_tiles; ++j) {
        const int st = (int)(j % in.nstage);
        const uint32_t parity = (uint32_t)((j / in.nstage) & 1);
        int64_t nks, nke, nbase, nrow;
        row_bounds(j + 1, nks, nke, nbase, nrow);   // overlaps with the wait and the row walk
        tma::mbar_wait(&bars[st], parity);
        if (row >= 0) {
            const double *__restrict__ sv = sval + (size_t)st * in.cap - base;
            const int *__restrict__ sc = scol + (size_t)st * in.cap - base;
            double sum = 0.0;
            int64_t k = ks;
            for (; k + 4 <= ke; k += 4) {
                double a[4], xv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = sc[k + u];
                    a[u] = sv[k + u];
                    if (GHOST && c >= in.n_owned) xv[u] = __ldg(in.ghost + (c - in.n_owned));
                    else xv[u] = __ldg(x + c);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) sum = add_rn(sum, mul_rn(a[u], xv[u]));
            }
            for (; k < ke; ++k) {
                const int c = sc[k];
                double xv;
                if (GHOST && c >= in.n_owned) xv = __ldg(in.ghost + (c - in.n_owned));
                else xv = __ldg(x + c);
                sum = add_rn(sum, mul_rn(sv[k], xv));
            }
            epi(row, sum, acc);
        }
        __syncthreads();   // every thread is done with stage `st`
        if (threadIdx.x == 0 && j + in.nstage < my_tiles) issue(j + in.nstage);
        ks = nks; ke = nke; base = nbase; row = nrow;
    }
    if constexpr (Epi::NRED > 0) block_reduce_finish<Epi::NRED>(acc, ra);
}
