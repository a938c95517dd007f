// bis_factor.cu -- device-side preprocessing of the triangular solves:
//   * strict L/U split of a device CRS             split_LU_new, LU_factors.hpp:122-309
//   * ILU(0), row-wise IKJ on A's pattern          factor_ILU0_old, LU_factors.hpp:320-539
//   * level analysis + level-ordered copy + gates  (what bis_sptrsv.cu consumes)
// Nothing is downloaded: the matrix that was generated or uploaded once stays where it is.
//
// Both the level analysis and the factorisation are the same dataflow problem as the triangular solve
// itself: row i needs rows k < i of its pattern (k > i for an upper factor) to be finished.  They run
// as ONE launch each, one thread per row, rows handed out in dependency-compatible order through a
// ticket counter (every dependency belongs to a block that already runs), and a thread that still
// waits publishes nothing and just looks again -- the publishing store sits INSIDE the wait loop, so a
// lane never has to leave the loop before the lanes of its own warp that depend on it can see its
// result (rows i and i-1 of a stencil matrix sit in the same warp).
//
// ILU(0) rounding follows the host restatement that is pinned to the compiled reference
// (host/lu_factors.hpp: factor by one division, row update by one fused multiply-add, updates only
// where the working value is != 0.0, pivots below 1e-16 skipped, small diagonals replaced).
#include "bis_device.cuh"

#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/fill.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>

#include <algorithm>

namespace {

constexpr int FB = 256;   // threads per block of the row-parallel kernels

template <typename T> int dalloc(T **p, size_t count) {
    BIS_CUDA(bis_cuda_malloc(p, sizeof(T) * (count > 0 ? count : 1)));
    return 0;
}

__device__ __forceinline__ int ld_relaxed_s32(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_acquire_s32(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_s32(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- level analysis ---------------------------------------------------------------------------
template <typename RP>
__global__ void tri_validate_kernel(int64_t n, const RP *rp, const int *col, int upper, unsigned long long *bad) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        for (RP k = rp[r]; k < rp[r + 1]; ++k) {
            const int c = col[k];
            const bool ok = upper ? (c > r && c < n) : (c >= 0 && c < r);
            if (!ok) atomicMin(bad, (unsigned long long)r);
        }
}

// level(r) = 1 + max level of the rows r reads (0 without dependencies); level[] starts at -1.
template <typename RP>
__global__ void __launch_bounds__(FB)
levels_dataflow_kernel(int64_t n, const RP *rp, const int *col, int upper, int *level, unsigned int *ticket, int *errflag) {
    __shared__ unsigned int s_chunk;
    if (threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1u);
    __syncthreads();
    const int64_t pos = (int64_t)s_chunk * FB + threadIdx.x;
    if (pos >= n) return;
    const int64_t r = upper ? n - 1 - pos : pos;
    RP k = rp[r];
    const RP e = rp[r + 1];
    int lv = 0;
    unsigned int spins = 0;
    unsigned long long t_wd = 0;
    bool finished = false;
    while (!finished) {
        const RP k_before = k;
        while (k < e) {
            const int l = ld_relaxed_s32(level + col[k]);
            if (l < 0) break;
            lv = max(lv, l + 1);
            ++k;
        }
        if (k != k_before) t_wd = 0;
        if (k == e) {
            __stcg(level + r, lv);   // publish inside the loop (see the file header)
            finished = true;
        } else if ((++spins & 0x3fffu) == 0) {   // watchdog: another kernel's error, or 60 s without progress
            if (t_wd == 0) t_wd = bis_globaltimer();
            if (*reinterpret_cast<volatile int *>(errflag)) {
                finished = true;
            } else if (bis_globaltimer() - t_wd > 60000000000ull) {
                atomicExch(errflag, 3);
                finished = true;
            }
        }
    }
}

__global__ void level_hist_kernel(int64_t n, const int *level, int *level_size) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(level_size + level[r], 1);
}

template <typename RP>
__global__ void slot_meta_kernel(int64_t n, const RP *rp, const int *slot_row, int *slot_of, int64_t *len) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s <= n; s += (int64_t)gridDim.x * blockDim.x) {
        if (s == n) {
            len[s] = 0;
            continue;
        }
        const int r = slot_row[s];
        slot_of[r] = (int)s;
        len[s] = (int64_t)(rp[r + 1] - rp[r]);
    }
}

// level-ordered copy (operands named by slot, storage order kept) and the per-row gates; the device
// twin of the analysis that bis_matrix_upload_triangular used to run on the host
template <typename RP>
__global__ void slot_fill_kernel(int64_t n, const RP *rp, const int *col, const double *val, const int *slot_row,
                                 const int *slot_of, const int *level, const int64_t *rp2, int *col2, double *val2,
                                 int *gate) {
    for (int64_t sl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; sl < n; sl += (int64_t)gridDim.x * blockDim.x) {
        const int r = slot_row[sl];
        const int lvr = level[r];
        int64_t o = rp2[sl];
        long long best1 = -1, best2 = -1, oldest = 0x7fffffffffffffffLL;
        int g[3] = {-1, -1, -1}, gl[3] = {0, 0, 0};
        for (RP k = rp[r]; k < rp[r + 1]; ++k, ++o) {
            const int cc = col[k];
            const long long so = slot_of[cc];
            col2[o] = (int)so;
            val2[o] = val[k];
            const int d = lvr - level[cc];
            if (so / 32 == sl / 32) continue;   // own warp: handled in lockstep, never a gate
            if (d >= 4 && so < oldest) { oldest = so; g[0] = (int)so; gl[0] = level[cc]; }
            if (d >= 3 && so > best1) { best1 = so; g[1] = (int)so; gl[1] = level[cc]; }
            if (d == 2 && so > best2) { best2 = so; g[2] = (int)so; gl[2] = level[cc]; }
        }
        if (g[0] == g[1]) g[0] = -1;
        int packed = 0;
        for (int i = 0; i < 3; ++i) {
            gate[4 * sl + i] = g[i];
            packed |= (g[i] >= 0 ? min(lvr - gl[i], 15) : 0) << (8 * i);
        }
        gate[4 * sl + 3] = packed;
    }
}

template <typename RP>
int build_levels(bis_context *c, bis_matrix *T) {
    const int64_t n = T->n_rows;
    const int upper = T->triangular == 2 ? 1 : 0;
    const RP *rp = static_cast<const RP *>(T->d_rp);
    cudaStream_t st = c->stream;
    auto pol = thrust::cuda::par.on(st);
    LevelSets &lv = T->lv;
    lv.n_slots = n;
    lv.n_levels = 0;
    BIS_CHECK(dalloc(&lv.d_ticket, 1));
    BIS_CHECK(dalloc(&lv.d_w, (size_t)n));
    BIS_CHECK(dalloc(&lv.d_w2, (size_t)n));
    // 0xFFF87E5E7E5E7E5E ("not ready", bis_sptrsv.cu) in both working vectors; every solve re-arms the other one
    if (n > 0) {
        thrust::fill(pol, thrust::device_pointer_cast(reinterpret_cast<unsigned long long *>(lv.d_w)),
                     thrust::device_pointer_cast(reinterpret_cast<unsigned long long *>(lv.d_w)) + n, 0xFFF87E5E7E5E7E5EULL);
        thrust::fill(pol, thrust::device_pointer_cast(reinterpret_cast<unsigned long long *>(lv.d_w2)),
                     thrust::device_pointer_cast(reinterpret_cast<unsigned long long *>(lv.d_w2)) + n, 0xFFF87E5E7E5E7E5EULL);
    }
    BIS_CHECK(dalloc(&lv.d_slot_row, (size_t)n));
    BIS_CHECK(dalloc(&lv.d_slot_level, (size_t)n));
    BIS_CHECK(dalloc(&lv.d_slot_gate, (size_t)n * 4 + 4));
    BIS_CHECK(dalloc(&lv.d_rp, (size_t)n + 1));
    BIS_CHECK(dalloc(&lv.d_col, (size_t)T->nnz));
    BIS_CHECK(dalloc(&lv.d_val, (size_t)T->nnz));
    lv.level_start.assign(1, 0);
    if (n == 0) {
        BIS_CHECK(dalloc(&lv.d_level_size, 1));
        BIS_CHECK(dalloc(&lv.d_level_done, 1));
        BIS_CUDA(cudaMemsetAsync(lv.d_rp, 0, sizeof(int64_t), st));
        return 0;
    }
    const int grid = bis_blocks_for(n, FB, c->sm_count * 16);
    // (1) strictly triangular?
    unsigned long long *d_bad = nullptr;
    BIS_CHECK(dalloc(&d_bad, 1));
    BIS_CUDA(cudaMemsetAsync(d_bad, 0xFF, sizeof(unsigned long long), st));
    tri_validate_kernel<RP><<<grid, FB, 0, st>>>(n, rp, T->d_col, upper, d_bad);
    BIS_LAUNCH_CHECK(c);
    unsigned long long bad = 0;
    BIS_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, st));
    BIS_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_bad);
    BIS_REQUIRE(bad == ~0ull, "bis_matrix_upload_triangular: entry not strictly %s the diagonal at row %lld",
                upper ? "above" : "below", (long long)bad);
    // (2) levels by dataflow
    int *d_level = nullptr, *d_slot_of = nullptr;
    BIS_CHECK(dalloc(&d_level, (size_t)n));
    BIS_CHECK(dalloc(&d_slot_of, (size_t)n));
    BIS_CUDA(cudaMemsetAsync(d_level, 0xFF, sizeof(int) * (size_t)n, st));
    BIS_CUDA(cudaMemsetAsync(lv.d_ticket, 0, sizeof(unsigned int), st));
    // a flag left behind by an earlier kernel's watchdog would make waiting rows give up at once: it is
    // reported (and cleared) here, before the analysis starts
    int stale = 0;
    BIS_CUDA(cudaMemcpyAsync(&stale, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, st));
    BIS_CUDA(cudaStreamSynchronize(st));
    if (stale) {
        cudaMemsetAsync(c->d_errflag, 0, sizeof(int), st);
        cudaFree(d_level);
        cudaFree(d_slot_of);
        bis_set_error("level analysis: an earlier kernel's device watchdog had fired (code %d)", stale);
        return 3;
    }
    levels_dataflow_kernel<RP><<<(unsigned)((n + FB - 1) / FB), FB, 0, st>>>(n, rp, T->d_col, upper, d_level, lv.d_ticket, c->d_errflag);
    BIS_LAUNCH_CHECK(c);
    // every row must have a level: a row that gave up keeps -1 and would index level_size[-1] below
    auto lvl_b = thrust::device_pointer_cast(d_level);
    const int max_level = thrust::reduce(pol, lvl_b, lvl_b + n, -1, thrust::maximum<int>());
    const int min_level = thrust::reduce(pol, lvl_b, lvl_b + n, 0x7fffffff, thrust::minimum<int>());
    int fired = 0;
    BIS_CUDA(cudaMemcpyAsync(&fired, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, st));
    BIS_CUDA(cudaStreamSynchronize(st));
    if (min_level < 0 || max_level < 0 || fired) {
        if (fired) cudaMemsetAsync(c->d_errflag, 0, sizeof(int), st);
        cudaFree(d_level);
        cudaFree(d_slot_of);
        bis_set_error("level analysis did not finish for every row (device watchdog code %d)", fired);
        return 3;
    }
    lv.n_levels = max_level + 1;
    // (3) rows by (level, row): a stable sort of the row ids by level
    BIS_CUDA(cudaMemcpyAsync(lv.d_slot_level, d_level, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    thrust::sequence(pol, thrust::device_pointer_cast(lv.d_slot_row), thrust::device_pointer_cast(lv.d_slot_row) + n);
    thrust::stable_sort_by_key(pol, thrust::device_pointer_cast(lv.d_slot_level),
                               thrust::device_pointer_cast(lv.d_slot_level) + n, thrust::device_pointer_cast(lv.d_slot_row));
    BIS_CHECK(dalloc(&lv.d_level_size, (size_t)lv.n_levels));
    BIS_CHECK(dalloc(&lv.d_level_done, (size_t)lv.n_levels));
    BIS_CUDA(cudaMemsetAsync(lv.d_level_size, 0, sizeof(int) * (size_t)lv.n_levels, st));
    level_hist_kernel<<<grid, FB, 0, st>>>(n, d_level, lv.d_level_size);
    BIS_LAUNCH_CHECK(c);
    // (4) level-ordered copy + gates
    slot_meta_kernel<RP><<<grid, FB, 0, st>>>(n, rp, lv.d_slot_row, d_slot_of, lv.d_rp);
    BIS_LAUNCH_CHECK(c);
    thrust::exclusive_scan(pol, thrust::device_pointer_cast(lv.d_rp), thrust::device_pointer_cast(lv.d_rp) + n + 1,
                           thrust::device_pointer_cast(lv.d_rp));
    slot_fill_kernel<RP><<<grid, FB, 0, st>>>(n, rp, T->d_col, T->d_val, lv.d_slot_row, d_slot_of, d_level, lv.d_rp,
                                              lv.d_col, lv.d_val, lv.d_slot_gate);
    BIS_LAUNCH_CHECK(c);
    std::vector<int> level_size((size_t)lv.n_levels);
    BIS_CUDA(cudaMemcpyAsync(level_size.data(), lv.d_level_size, sizeof(int) * (size_t)lv.n_levels, cudaMemcpyDeviceToHost, st));
    BIS_CUDA(cudaStreamSynchronize(st));
    lv.level_start.assign((size_t)lv.n_levels + 1, 0);
    for (int l = 0; l < lv.n_levels; ++l) lv.level_start[l + 1] = lv.level_start[l] + level_size[l];
    lv.d_level = d_level;   // kept: the chain format (bis_sptrsv_chain.cuh) is built from it on demand
    cudaFree(d_slot_of);
    return 0;
}

// ---- split ---------------------------------------------------------------------------------------
template <typename RP>
__global__ void split_count_kernel(int64_t n, const RP *rp, const int *col, int *nl, int *nu) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (int64_t)gridDim.x * blockDim.x) {
        int a = 0, b = 0;
        if (r < n)
            for (RP k = rp[r]; k < rp[r + 1]; ++k) {
                const int c = col[k];
                a += c < r;
                b += c > r;
            }
        nl[r] = a;
        nu[r] = b;
    }
}

// entries keep A's within-row order (split_LU_new) unless `sorted` (factor_ILU0_old walks the columns
// of a row in ascending order and stores its factors that way); diag receives A[r][r] (0.0 when absent)
template <typename RP>
__global__ void split_fill_kernel(int64_t n, const RP *rp, const int *col, const double *val, const int *lrp,
                                  const int *urp, int *lcol, double *lval, int *ucol, double *uval, double *diag,
                                  int sorted) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        int pl = lrp[r], pu = urp[r];
        double d = 0.0;
        for (RP k = rp[r]; k < rp[r + 1]; ++k) {
            const int c = col[k];
            const double v = val[k];
            if (c < r) { lcol[pl] = c; lval[pl++] = v; }
            else if (c > r) { ucol[pu] = c; uval[pu++] = v; }
            else d = v;
        }
        if (diag) diag[r] = d;
        if (sorted) {
            for (int part = 0; part < 2; ++part) {   // insertion sort: rows are short and mostly sorted already
                int *cc = part ? ucol : lcol;
                double *vv = part ? uval : lval;
                const int b = part ? urp[r] : lrp[r], e = part ? pu : pl;
                for (int i = b + 1; i < e; ++i) {
                    const int ci = cc[i];
                    const double vi = vv[i];
                    int j = i - 1;
                    while (j >= b && cc[j] > ci) { cc[j + 1] = cc[j]; vv[j + 1] = vv[j]; --j; }
                    cc[j + 1] = ci;
                    vv[j + 1] = vi;
                }
            }
        }
    }
}

int new_triangular(bis_context *c, int64_t n, int64_t nnz, int kind, bis_matrix **out) {
    bis_matrix *T = new bis_matrix;
    T->n_rows = T->n_cols = T->n_rows_global = n;
    T->nnz = T->nnz_global = nnz;
    T->rp_bytes = 4;
    T->triangular = kind;
    int *rp = nullptr;
    if (dalloc(&rp, (size_t)n + 1 + 8) || dalloc(&T->d_col, (size_t)nnz + 8) || dalloc(&T->d_val, (size_t)nnz + 8)) {
        cudaFree(rp);
        cudaFree(T->d_col);
        cudaFree(T->d_val);
        delete T;
        return 1;
    }
    T->d_rp = rp;
    (void)c;
    *out = T;
    return 0;
}

// L and U of A with their level sets still missing (the caller factors in place first, or not)
template <typename RP>
int split_device(bis_context *c, const bis_matrix *A, int sorted, double *diag, bis_matrix **Lo, bis_matrix **Uo) {
    const int64_t n = A->n_rows;
    cudaStream_t st = c->stream;
    auto pol = thrust::cuda::par.on(st);
    const RP *rp = static_cast<const RP *>(A->d_rp);
    int *nl = nullptr, *nu = nullptr;
    BIS_CHECK(dalloc(&nl, (size_t)n + 1));
    BIS_CHECK(dalloc(&nu, (size_t)n + 1));
    const int grid = bis_blocks_for(n + 1, FB, c->sm_count * 16);
    split_count_kernel<RP><<<grid, FB, 0, st>>>(n, rp, A->d_col, nl, nu);
    BIS_LAUNCH_CHECK(c);
    auto pl = thrust::device_pointer_cast(nl), pu = thrust::device_pointer_cast(nu);
    const int64_t tl = thrust::reduce(pol, pl, pl + n, (int64_t)0), tu = thrust::reduce(pol, pu, pu + n, (int64_t)0);
    const int ml = thrust::reduce(pol, pl, pl + n, 0, thrust::maximum<int>());
    const int mu = thrust::reduce(pol, pu, pu + n, 0, thrust::maximum<int>());
    if (tl >= INT32_MAX || tu >= INT32_MAX) {
        cudaFree(nl);
        cudaFree(nu);
        bis_set_error("bis_matrix_split_triangular: a factor has more than 2^31-1 nonzeros");
        return 2;
    }
    bis_matrix *L = nullptr, *U = nullptr;
    if (new_triangular(c, n, tl, 1, &L) != 0 || new_triangular(c, n, tu, 2, &U) != 0) {
        cudaFree(nl);
        cudaFree(nu);
        if (L) bis_matrix_free(c, L);
        bis_set_error("bis_matrix_split_triangular: out of device memory");
        return 1;
    }
    L->max_row = ml; L->mean_row = n ? (double)tl / (double)n : 0.0;
    U->max_row = mu; U->mean_row = n ? (double)tu / (double)n : 0.0;
    L->grid_nx = U->grid_nx = A->grid_nx;   // the factors of a grid matrix live on the same grid
    L->grid_ny = U->grid_ny = A->grid_ny;
    L->grid_nz = U->grid_nz = A->grid_nz;
    int *lrp = static_cast<int *>(L->d_rp), *urp = static_cast<int *>(U->d_rp);
    thrust::exclusive_scan(pol, pl, pl + n + 1, thrust::device_pointer_cast(lrp));
    thrust::exclusive_scan(pol, pu, pu + n + 1, thrust::device_pointer_cast(urp));
    split_fill_kernel<RP><<<grid, FB, 0, st>>>(n, rp, A->d_col, A->d_val, lrp, urp, L->d_col, L->d_val, U->d_col,
                                               U->d_val, diag, sorted);
    BIS_LAUNCH_CHECK(c);
    BIS_CUDA(cudaStreamSynchronize(st));
    cudaFree(nl);
    cudaFree(nu);
    *Lo = L;
    *Uo = U;
    return 0;
}

// ---- ILU(0) --------------------------------------------------------------------------------------
// In place on the sorted split: lval[] starts as A's strictly lower entries and ends as L's, uval[] and
// ud[] start as A's upper entries / diagonal and end as U's.  done[k] != 0: row k of U and ud[k] are final.
__device__ __forceinline__ int find_col(const int *cols, int lo, int hi, int j) {
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        const int cm = cols[m];
        if (cm < j) lo = m + 1;
        else hi = m;
    }
    return lo;
}

__global__ void __launch_bounds__(FB)
ilu0_dataflow_kernel(int64_t n, const int *lrp, const int *lcol, double *lval, const int *urp, const int *ucol,
                     double *uval, double *ud, double *ld, int *done, unsigned int *ticket, double pivot_tol,
                     double pivot_repl, int *errflag) {
    __shared__ unsigned int s_chunk;
    if (threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1u);
    __syncthreads();
    const int64_t i = (int64_t)s_chunk * FB + threadIdx.x;
    if (i >= n) return;
    int kk = lrp[i];
    const int ke = lrp[i + 1], ub = urp[i], ue = urp[i + 1];
    unsigned int spins = 0;
    unsigned long long t_wd = 0;
    bool finished = false;
    while (!finished) {
        if (kk < ke) {
            const int k = lcol[kk];
            if (ld_acquire_s32(done + k)) {
                const double pivot = __ldcg(ud + k);
                if (fabs(pivot) >= 1e-16) {   // an unusable pivot skips this elimination (LU_factors.hpp:370)
                    const double factor = div_rn(__ldcg(lval + kk), pivot);
                    __stcg(lval + kk, factor);
                    for (int t = urp[k]; t < urp[k + 1]; ++t) {
                        const int j = ucol[t];
                        double *p = nullptr;
                        if (j < i) {
                            const int q = find_col(lcol, kk + 1, ke, j);
                            if (q < ke && lcol[q] == j) p = lval + q;
                        } else if (j == i) {
                            p = ud + i;
                        } else {
                            const int q = find_col(ucol, ub, ue, j);
                            if (q < ue && ucol[q] == j) p = uval + q;
                        }
                        if (p) {   // only on A's pattern, and only where the working value is non-zero (:384)
                            const double w = __ldcg(p);
                            if (w != 0.0) __stcg(p, fma(-factor, __ldcg(uval + t), w));
                        }
                    }
                }
                ++kk;
                spins = 0;
                t_wd = 0;
            } else if ((++spins & 0x3fffu) == 0) {   // watchdog: another kernel's error, or 60 s without progress
                if (t_wd == 0) t_wd = bis_globaltimer();
                if (*reinterpret_cast<volatile int *>(errflag)) {
                    finished = true;
                } else if (bis_globaltimer() - t_wd > 60000000000ull) {
                    atomicExch(errflag, 4);
                    finished = true;
                }
            }
        } else {
            double u = __ldcg(ud + i);
            if (fabs(u) < pivot_tol) u = (u >= 0.0 ? 1.0 : -1.0) * pivot_repl;   // :410-412
            __stcg(ud + i, u);
            if (ld) ld[i] = 1.0;
            __threadfence();
            st_release_s32(done + i, 1);   // publish inside the loop (see the file header)
            finished = true;
        }
    }
}

} // namespace

static int build_levels_now(bis_context *c, bis_matrix *T) {
    BIS_CUDA(cudaSetDevice(c->device));
    int rc = T->rp_bytes == 8 ? build_levels<int64_t>(c, T) : build_levels<int32_t>(c, T);
    if (rc == 0) T->lv.built = true;
    return rc;
}

// What a triangular solve needs beyond the CRS arrays: the level sets of the dataflow solve and -- when the
// factor is a stencil on a structured grid and the wavefront is expected to be faster (trsv_variant = 0, the
// default: the cost model in wave_build_t; = 5 forces it and then skips the level analysis) -- the records of
// the stencil wavefront (bis_sptrsv_wave.cuh).
int bis_build_levels_device(bis_context *c, bis_matrix *T) {
    T->lv.n_slots = T->n_rows;
    if (c->opt_trsv_variant == 5) {
        BIS_CHECK(bis_wave_build(c, T));
        if (T->lv.wave.state == 1) {
            // validation is the wavefront's own (strictly triangular, ascending columns, stencil slots)
            return 0;
        }
    }
    BIS_CHECK(build_levels_now(c, T));
    if (c->opt_trsv_variant == 0) BIS_CHECK(bis_wave_build(c, T));
    return 0;
}

// Level sets (and, where chosen, wavefront records) of both factors.  With factor_keep_crs = 0 a factor's
// natural-order CRS is freed as soon as its level-ordered copy exists -- before the other factor's copy is built, so the
// peak is one factor lower too.  The triangular solves only read the level-ordered copy; the Gauss-Seidel SWEEPS
// (b - T x) and the two-stage preconditioners read the natural one, so the host asks for this only when a Krylov
// method uses the factors as a gs / bgs / sgs / ilu0 preconditioner.  HPCG-512 -cg -p sgs: 177 -> 131 GB, fits one GPU.
static int build_levels_pair(bis_context *c, bis_matrix *l, bis_matrix *u) {
    for (bis_matrix *T : {l, u}) {
        if (bis_build_levels_device(c, T) != 0) return 1;
        if (!c->opt_factor_keep_crs && T->lv.built) {
            BIS_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(T->d_rp);
            cudaFree(T->d_col);
            cudaFree(T->d_val);
            T->d_rp = nullptr;
            T->d_col = nullptr;
            T->d_val = nullptr;
            T->crs_released = true;
        }
    }
    return 0;
}

int bis_ensure_levels(bis_context *c, const bis_matrix *T) {
    if (T->lv.built) return 0;
    return build_levels_now(c, const_cast<bis_matrix *>(T));
}

// split_LU_new (LU_factors.hpp:122-309), strict parts, on the device.
extern "C" int bis_matrix_split_triangular(bis_context *c, const bis_matrix *A, bis_matrix **L, bis_matrix **U) {
    BIS_REQUIRE(c && A && L && U, "null argument");
    BIS_REQUIRE(!A->distributed && c->nranks == 1, "bis_matrix_split_triangular: single-GPU only");
    BIS_REQUIRE_CRS(A);
    BIS_CUDA(cudaSetDevice(c->device));
    bis_vector_cache_trim(c);
    bis_matrix *l = nullptr, *u = nullptr;
    if (A->rp_bytes == 8) BIS_CHECK(split_device<int64_t>(c, A, 0, nullptr, &l, &u));
    else BIS_CHECK(split_device<int32_t>(c, A, 0, nullptr, &l, &u));
    if (build_levels_pair(c, l, u) != 0) {
        bis_matrix_free(c, l);
        bis_matrix_free(c, u);
        return 1;
    }
    *L = l;
    *U = u;
    return 0;
}

// factor_ILU0_old (LU_factors.hpp:320-539) on the device: L_strict, U_strict (columns ascending), L_D = 1, U_D.
extern "C" int bis_matrix_ilu0(bis_context *c, const bis_matrix *A, double pivot_tolerance, double pivot_replacement,
                               bis_matrix **L, bis_matrix **U, double *L_D, double *U_D) {
    BIS_REQUIRE(c && A && L && U && U_D, "null argument");
    BIS_REQUIRE(!A->distributed && c->nranks == 1, "bis_matrix_ilu0: single-GPU only");
    BIS_REQUIRE_CRS(A);
    BIS_CUDA(cudaSetDevice(c->device));
    bis_vector_cache_trim(c);
    const int64_t n = A->n_rows;
    bis_matrix *l = nullptr, *u = nullptr;
    if (A->rp_bytes == 8) BIS_CHECK(split_device<int64_t>(c, A, 1, U_D, &l, &u));
    else BIS_CHECK(split_device<int32_t>(c, A, 1, U_D, &l, &u));
    int rc = 0;
    if (n > 0) {
        int *done = nullptr;
        unsigned int *ticket = nullptr;
        if (dalloc(&done, (size_t)n) || dalloc(&ticket, 1)) rc = 1;
        if (!rc) {
            cudaMemsetAsync(done, 0, sizeof(int) * (size_t)n, c->stream);
            cudaMemsetAsync(ticket, 0, sizeof(unsigned int), c->stream);
            ilu0_dataflow_kernel<<<(unsigned)((n + FB - 1) / FB), FB, 0, c->stream>>>(
                n, static_cast<const int *>(l->d_rp), l->d_col, l->d_val, static_cast<const int *>(u->d_rp), u->d_col,
                u->d_val, U_D, L_D, done, ticket, pivot_tolerance, pivot_replacement, c->d_errflag);
            c->launches++;
            if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) {
                bis_set_error("bis_matrix_ilu0: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = 1;
            }
            int flag = 0;
            cudaMemcpy(&flag, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost);
            if (!rc && flag) {
                cudaMemset(c->d_errflag, 0, sizeof(int));
                bis_set_error("bis_matrix_ilu0: the factorisation did not finish (device watchdog %d)", flag);
                rc = 3;
            }
        }
        cudaFree(done);
        cudaFree(ticket);
    }
    if (!rc && build_levels_pair(c, l, u) != 0) rc = 1;
    if (rc) {
        bis_matrix_free(c, l);
        bis_matrix_free(c, u);
        return rc;
    }
    *L = l;
    *U = u;
    return 0;
}
