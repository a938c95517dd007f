// bis_perm.cu -- the permutation seam of the reference (preprocessing.hpp:52-65, utilities/smax_helpers.hpp:44-80:
// generate_perm / apply_mat_perm / apply_vec_perm of the absent SMAX library, PERM_MODE = C "colouring"), as a
// LABELLED mode of the build: A <- P A P^T, b <- P b, x_0 <- P x_0 with P a multicolouring of A's graph.
//
// Why it is worth having on a GPU: rows of one colour do not read one another, so a triangular factor of the
// permuted matrix has as many levels as there are colours (8 for a 27-point stencil instead of 7n-6): the
// level-scheduled solve becomes a bandwidth-bound kernel.  The price is a weaker Gauss-Seidel / ILU
// preconditioner (iteration counts change), which is why the mode is opt-in and has its own fixtures: the
// oracle for it is the CPU restatement run on the explicitly permuted system (tests/test_perm_gpu.py).
//
// Colouring: on a structured grid (the generators' hint) colour = (x&1) + 2 (y&1) + 4 (z&1) -- exact for every
// stencil that stays inside the 3 x 3 x 3 box; otherwise Luby / Jones-Plassmann rounds on the device (a vertex
// takes the round's colour when its hashed priority beats all of its still uncoloured neighbours).  The
// permutation lists the rows by (colour, row); inside a row the entries keep their stored order (the summation
// order), with their columns renumbered.
#include "bis_device.cuh"

#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#include <thrust/binary_search.h>
#include <thrust/count.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/transform_reduce.h>
#include <thrust/functional.h>
#include <thrust/reverse.h>

#include <vector>

namespace {

__device__ __forceinline__ unsigned int prio_hash(unsigned int v) {
    v ^= v >> 16;
    v *= 0x7feb352du;
    v ^= v >> 15;
    v *= 0x846ca68bu;
    v ^= v >> 16;
    return v;
}

__global__ void colour_grid_kernel(int64_t n, int nx, int ny, int *colour) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(r % nx), y = (int)((r / nx) % ny), z = (int)(r / ((int64_t)nx * ny));
        colour[r] = (x & 1) | ((y & 1) << 1) | ((z & 1) << 2);
    }
}

// one round: uncoloured vertices whose (hash, index) beats every uncoloured neighbour take colour `round`
template <typename RP>
__global__ void colour_round_kernel(int64_t n, const RP *rp, const int *col, int round, const int *colour_in, int *colour_out,
                                    int *remaining) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        int cr = colour_in[r];
        if (cr < 0) {
            const unsigned int pr = prio_hash((unsigned int)r);
            bool top = true;
            for (RP k = rp[r]; k < rp[r + 1] && top; ++k) {
                const int c = col[k];
                if (c == r || c < 0 || c >= n || colour_in[c] >= 0) continue;
                const unsigned int pc = prio_hash((unsigned int)c);
                if (pc > pr || (pc == pr && c > r)) top = false;
            }
            if (top) cr = round;
            else atomicAdd(remaining, 1);
        }
        colour_out[r] = cr;
    }
}

// does any row read a row of its own colour?  (structurally nonsymmetric patterns can defeat the rounds above)
template <typename RP>
__global__ void colour_check_kernel(int64_t n, const RP *rp, const int *col, const int *colour, int *bad) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        for (RP k = rp[r]; k < rp[r + 1]; ++k) {
            const int c = col[k];
            if (c != r && c >= 0 && c < n && colour[c] == colour[r]) atomicAdd(bad, 1);
        }
}

__global__ void invert_perm_kernel(int64_t n, const int *perm, int *inv) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) inv[perm[i]] = (int)i;
}

template <typename RP>
__global__ void perm_len_kernel(int64_t n, const RP *rp, const int *perm, int64_t *len) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x)
        len[i] = i < n ? (int64_t)(rp[perm[i] + 1] - rp[perm[i]]) : 0;
}

template <typename RP>
__global__ void perm_fill_kernel(int64_t n, const RP *rp, const int *col, const double *val, const int *perm, const int *inv,
                                 const int64_t *rp2, RP *rp_out, int *col2, double *val2) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
        rp_out[i] = (RP)rp2[i];
        if (i == n) continue;
        const int old = perm[i];
        int64_t o = rp2[i];
        for (RP k = rp[old]; k < rp[old + 1]; ++k, ++o) {   // stored order kept: it is the summation order
            col2[o] = inv[col[k]];
            val2[o] = val[k];
        }
    }
}

__global__ void vec_perm_kernel(int64_t n, const double *in, const int *perm, double *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[perm[i]];
}


// ---- breadth-first / (reverse) Cuthill-McKee orderings ------------------------------------------------------
template <typename RP>
__global__ void degree_kernel(int64_t n, const RP *rp, const int *col, int *deg) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        int d = 0;
        for (RP k = rp[r]; k < rp[r + 1]; ++k) d += (col[k] != r && col[k] >= 0 && col[k] < n) ? 1 : 0;
        deg[r] = d;
    }
}

// one round of the level-synchronous search: rows of level `cur` give level cur + 1 to the rows they read that have none
// yet (every writer stores the same number: the race is benign)
template <typename RP>
__global__ void bfs_round_kernel(int64_t n, const RP *rp, const int *col, int cur, int *level, int *found) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        if (level[r] != cur) continue;
        bool any = false;
        for (RP k = rp[r]; k < rp[r + 1]; ++k) {
            const int c = col[k];
            if (c < 0 || c >= n || c == r) continue;
            if (level[c] < 0) {
                level[c] = cur + 1;
                any = true;
            }
        }
        if (any) *reinterpret_cast<volatile int *>(found) = 1;
    }
}

// Cuthill-McKee key of the rows ids[s..e) of level l: (position of the earliest-numbered row of level l - 1 the row
// reads, degree); rows that read none of them (roots of further components) come by degree alone
template <typename RP>
__global__ void cm_key_kernel(int64_t s, int64_t e, int l, const RP *rp, const int *col, const int *level, const int *deg,
                              const int *pos, const int *ids, unsigned long long *key, int64_t n) {
    for (int64_t i = s + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = ids[i];
        unsigned int parent = 0xFFFFFFFFu;
        for (RP k = rp[v]; k < rp[v + 1]; ++k) {
            const int c = col[k];
            if (c < 0 || c >= n || c == v) continue;
            if (level[c] == l - 1) parent = min(parent, (unsigned int)pos[c]);
        }
        const unsigned int d = min((unsigned int)deg[v], 0xFFFFFFu);
        key[i - s] = ((unsigned long long)parent << 24) | d;
    }
}

__global__ void cm_pos_kernel(int64_t s, int64_t e, const int *ids, int *pos) {
    for (int64_t i = s + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) pos[ids[i]] = (int)i;
}

struct UnvisitedKey {
    const int *level, *deg;
    __device__ unsigned long long operator()(int v) const {
        return level[v] >= 0 ? 0xFFFFFFFFFFFFFFFFull : (((unsigned long long)(unsigned int)deg[v] << 32) | (unsigned int)v);
    }
};
struct IsVisited {
    __device__ bool operator()(int l) const { return l >= 0; }
};

template <typename T> int dalloc(T **p, size_t count) {
    BIS_CUDA(bis_cuda_malloc(p, sizeof(T) * (count > 0 ? count : 1)));
    return 0;
}

} // namespace

// generate_perm (smax_helpers.hpp:51-53) for PERM_MODE = C: perm[new] = old, inv_perm[old] = new, both [dev] int32[n]
extern "C" int bis_matrix_colouring_permutation(bis_context *c, const bis_matrix *A, int *d_perm, int *d_inv_perm, int *n_colours) {
    BIS_REQUIRE(c && A && d_perm && d_inv_perm, "bis_matrix_colouring_permutation: null argument");
    BIS_REQUIRE_CRS(A);
    BIS_REQUIRE(!A->distributed && c->nranks == 1, "bis_matrix_colouring_permutation: single-GPU only");
    BIS_CUDA(cudaSetDevice(c->device));
    const int64_t n = A->n_rows;
    if (n_colours) *n_colours = 0;
    if (n == 0) return 0;
    cudaStream_t st = c->stream;
    auto pol = thrust::cuda::par.on(st);
    const int grid = bis_blocks_for(n, 256, c->sm_count * 16);
    int *colour = nullptr, *colour2 = nullptr, *d_cnt = nullptr;
    BIS_CHECK(dalloc(&colour, (size_t)n));
    BIS_CHECK(dalloc(&colour2, (size_t)n));
    BIS_CHECK(dalloc(&d_cnt, 1));
    auto cleanup = [&]() {
        cudaFree(colour);
        cudaFree(colour2);
        cudaFree(d_cnt);
    };
    int colours = 0;
    bool done = false;
    if (A->grid_nx > 0 && A->grid_nx * A->grid_ny * A->grid_nz == n && A->max_row <= 27) {
        colour_grid_kernel<<<grid, 256, 0, st>>>(n, (int)A->grid_nx, (int)A->grid_ny, colour);
        c->launches++;
        cudaMemsetAsync(d_cnt, 0, sizeof(int), st);
        if (A->rp_bytes == 8) colour_check_kernel<int64_t><<<grid, 256, 0, st>>>(n, static_cast<const int64_t *>(A->d_rp), A->d_col, colour, d_cnt);
        else colour_check_kernel<int32_t><<<grid, 256, 0, st>>>(n, static_cast<const int32_t *>(A->d_rp), A->d_col, colour, d_cnt);
        c->launches++;
        int bad = 0;
        cudaMemcpyAsync(&bad, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        if (bad == 0) {
            colours = 8;
            done = true;
        }
    }
    if (!done) {
        cudaMemsetAsync(colour, 0xFF, sizeof(int) * (size_t)n, st);
        for (int round = 0; round < 4096; ++round) {
            cudaMemsetAsync(d_cnt, 0, sizeof(int), st);
            if (A->rp_bytes == 8)
                colour_round_kernel<int64_t><<<grid, 256, 0, st>>>(n, static_cast<const int64_t *>(A->d_rp), A->d_col, round, colour, colour2, d_cnt);
            else
                colour_round_kernel<int32_t><<<grid, 256, 0, st>>>(n, static_cast<const int32_t *>(A->d_rp), A->d_col, round, colour, colour2, d_cnt);
            c->launches++;
            std::swap(colour, colour2);
            int remaining = 0;
            cudaMemcpyAsync(&remaining, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            colours = round + 1;
            if (remaining == 0) {
                done = true;
                break;
            }
        }
    }
    if (!done || cudaGetLastError() != cudaSuccess) {
        cleanup();
        bis_set_error("bis_matrix_colouring_permutation: colouring did not finish");
        return 3;
    }
    // rows by (colour, row): a stable sort of the row ids by colour
    thrust::sequence(pol, thrust::device_pointer_cast(d_perm), thrust::device_pointer_cast(d_perm) + n);
    thrust::stable_sort_by_key(pol, thrust::device_pointer_cast(colour), thrust::device_pointer_cast(colour) + n,
                               thrust::device_pointer_cast(d_perm));
    invert_perm_kernel<<<grid, 256, 0, st>>>(n, d_perm, d_inv_perm);
    BIS_LAUNCH_CHECK(c);
    BIS_CUDA(cudaStreamSynchronize(st));
    cleanup();
    if (n_colours) *n_colours = colours;
    return 0;
}

// generate_perm (smax_helpers.hpp:51-53) for the breadth-first family of PERM_MODEs.  mode 2: rows by (BFS level, row);
// mode 4: Cuthill-McKee (inside a level by the position of the earliest-numbered row of the previous level the row
// reads, then by degree, then by row), mode 3: that order reversed (RCM).  The numbers are those of the context option
// "perm_mode" (1 is the multicolouring above).  The search starts at the row of smallest
// degree (smallest index among equals) and restarts there among the rows not reached yet; neighbours are the columns
// of a row (the matrices of this code are structurally symmetric; for others this is the search on the directed graph).
// Deterministic: tests/test_perm_gpu.py restates it in numpy and compares the permutations entry for entry.
extern "C" int bis_matrix_bfs_permutation(bis_context *c, const bis_matrix *A, int mode, int *d_perm, int *d_inv_perm, int *n_levels) {
    BIS_REQUIRE(c && A && d_perm && d_inv_perm, "bis_matrix_bfs_permutation: null argument");
    BIS_REQUIRE(mode >= 2 && mode <= 4, "bis_matrix_bfs_permutation: mode must be 2 (BFS), 3 (reverse Cuthill-McKee) or 4 (Cuthill-McKee)");
    BIS_REQUIRE_CRS(A);
    BIS_REQUIRE(!A->distributed && c->nranks == 1, "bis_matrix_bfs_permutation: single-GPU only");
    BIS_CUDA(cudaSetDevice(c->device));
    const int64_t n = A->n_rows;
    if (n_levels) *n_levels = 0;
    if (n == 0) return 0;
    BIS_REQUIRE(n < ((int64_t)1 << 31), "bis_matrix_bfs_permutation: more than 2^31 rows");
    cudaStream_t st = c->stream;
    auto pol = thrust::cuda::par.on(st);
    const int grid = bis_blocks_for(n, 256, c->sm_count * 16);
    int *level = nullptr, *deg = nullptr, *d_found = nullptr, *pos = nullptr, *lkey = nullptr, *lstart = nullptr;
    unsigned long long *key = nullptr;
    auto cleanup = [&]() {
        cudaFree(level); cudaFree(deg); cudaFree(d_found); cudaFree(pos); cudaFree(lkey); cudaFree(lstart); cudaFree(key);
    };
    BIS_CHECK(dalloc(&level, (size_t)n));
    BIS_CHECK(dalloc(&deg, (size_t)n));
    BIS_CHECK(dalloc(&d_found, 1));
    BIS_CHECK(dalloc(&lkey, (size_t)n));
    const bool rp64 = A->rp_bytes == 8;
    const int64_t *rp8 = static_cast<const int64_t *>(A->d_rp);
    const int32_t *rp4 = static_cast<const int32_t *>(A->d_rp);
    if (rp64) degree_kernel<int64_t><<<grid, 256, 0, st>>>(n, rp8, A->d_col, deg);
    else degree_kernel<int32_t><<<grid, 256, 0, st>>>(n, rp4, A->d_col, deg);
    c->launches++;
    cudaMemsetAsync(level, 0xFF, sizeof(int) * (size_t)n, st);
    int cur = 0;
    int64_t visited = 0;
    bool ok = true;
    for (int comp = 0; visited < n && ok; ++comp) {
        if (comp >= 65536) {     // one search (with host round trips) per component: not meant for graphs of fragments
            ok = false;
            break;
        }
        const unsigned long long rk = thrust::transform_reduce(pol, thrust::counting_iterator<int>(0), thrust::counting_iterator<int>((int)n),
                                                               UnvisitedKey{level, deg}, 0xFFFFFFFFFFFFFFFFull, thrust::minimum<unsigned long long>());
        const int root = (int)(rk & 0xFFFFFFFFull);
        cudaMemcpyAsync(level + root, &cur, sizeof(int), cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);      // `cur` is a host variable that changes below
        for (;;) {
            cudaMemsetAsync(d_found, 0, sizeof(int), st);
            if (rp64) bfs_round_kernel<int64_t><<<grid, 256, 0, st>>>(n, rp8, A->d_col, cur, level, d_found);
            else bfs_round_kernel<int32_t><<<grid, 256, 0, st>>>(n, rp4, A->d_col, cur, level, d_found);
            c->launches++;
            int found = 0;
            cudaMemcpyAsync(&found, d_found, sizeof(int), cudaMemcpyDeviceToHost, st);
            if (cudaStreamSynchronize(st) != cudaSuccess) {
                ok = false;
                break;
            }
            ++cur;
            if (!found) break;
        }
        visited = thrust::count_if(pol, thrust::device_pointer_cast(level), thrust::device_pointer_cast(level) + n, IsVisited());
    }
    if (!ok || cudaGetLastError() != cudaSuccess) {
        cleanup();
        bis_set_error("bis_matrix_bfs_permutation: the search did not finish (more than 65536 components, or a device error)");
        return 3;
    }
    const int L = cur;      // levels 0 .. L-1 (a component's levels follow those of the one before it)
    // rows by (level, row)
    cudaMemcpyAsync(lkey, level, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, st);
    thrust::sequence(pol, thrust::device_pointer_cast(d_perm), thrust::device_pointer_cast(d_perm) + n);
    thrust::stable_sort_by_key(pol, thrust::device_pointer_cast(lkey), thrust::device_pointer_cast(lkey) + n, thrust::device_pointer_cast(d_perm));
    if (mode >= 3) {
        BIS_CHECK(dalloc(&pos, (size_t)n));
        BIS_CHECK(dalloc(&lstart, (size_t)L + 1));
        BIS_CHECK(dalloc(&key, (size_t)n));
        thrust::lower_bound(pol, thrust::device_pointer_cast(lkey), thrust::device_pointer_cast(lkey) + n, thrust::counting_iterator<int>(0),
                            thrust::counting_iterator<int>(L + 1), thrust::device_pointer_cast(lstart));
        std::vector<int> hs((size_t)L + 1);
        cudaMemcpyAsync(hs.data(), lstart, sizeof(int) * ((size_t)L + 1), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        for (int l = 0; l < L; ++l) {
            const int64_t s = hs[l], e = hs[l + 1];
            if (e <= s) continue;
            const int g = bis_blocks_for(e - s, 256, c->sm_count * 16);
            if (e - s > 1) {
                if (rp64) cm_key_kernel<int64_t><<<g, 256, 0, st>>>(s, e, l, rp8, A->d_col, level, deg, pos, d_perm, key, n);
                else cm_key_kernel<int32_t><<<g, 256, 0, st>>>(s, e, l, rp4, A->d_col, level, deg, pos, d_perm, key, n);
                c->launches++;
                thrust::stable_sort_by_key(pol, thrust::device_pointer_cast(key), thrust::device_pointer_cast(key) + (e - s),
                                           thrust::device_pointer_cast(d_perm) + s);
            }
            cm_pos_kernel<<<g, 256, 0, st>>>(s, e, d_perm, pos);
            c->launches++;
        }
        if (mode == 3) thrust::reverse(pol, thrust::device_pointer_cast(d_perm), thrust::device_pointer_cast(d_perm) + n);
    }
    invert_perm_kernel<<<grid, 256, 0, st>>>(n, d_perm, d_inv_perm);
    BIS_LAUNCH_CHECK(c);
    BIS_CUDA(cudaStreamSynchronize(st));
    cleanup();
    if (n_levels) *n_levels = L;
    return 0;
}

// apply_mat_perm (smax_helpers.hpp:56-58): B = P A P^T, a new matrix (the generators' grid hint does not survive)
extern "C" int bis_matrix_permute_symmetric(bis_context *c, const bis_matrix *A, const int *d_perm, const int *d_inv_perm,
                                            bis_matrix **out) {
    BIS_REQUIRE(c && A && d_perm && d_inv_perm && out, "bis_matrix_permute_symmetric: null argument");
    BIS_REQUIRE_CRS(A);
    BIS_REQUIRE(!A->distributed && c->nranks == 1 && A->triangular == 0 && A->n_rows == A->n_cols,
                "bis_matrix_permute_symmetric: square general single-GPU matrices only");
    BIS_CUDA(cudaSetDevice(c->device));
    bis_vector_cache_trim(c);
    const int64_t n = A->n_rows;
    cudaStream_t st = c->stream;
    bis_matrix *B = new bis_matrix;
    B->n_rows = B->n_cols = B->n_rows_global = n;
    B->nnz = B->nnz_global = A->nnz;
    B->rp_bytes = A->rp_bytes;
    B->max_row = A->max_row;
    B->mean_row = A->mean_row;
    int64_t *rp2 = nullptr;
    int rc = dalloc(&rp2, (size_t)n + 1) | dalloc(&B->d_col, (size_t)A->nnz + 8) | dalloc(&B->d_val, (size_t)A->nnz + 8);
    if (A->rp_bytes == 8) rc |= dalloc(reinterpret_cast<int64_t **>(&B->d_rp), (size_t)n + 1 + 8);
    else rc |= dalloc(reinterpret_cast<int32_t **>(&B->d_rp), (size_t)n + 1 + 8);
    if (rc) {
        cudaFree(rp2);
        bis_matrix_free(c, B);
        return 1;
    }
    const int grid = bis_blocks_for(n + 1, 256, c->sm_count * 16);
    if (A->rp_bytes == 8) perm_len_kernel<int64_t><<<grid, 256, 0, st>>>(n, static_cast<const int64_t *>(A->d_rp), d_perm, rp2);
    else perm_len_kernel<int32_t><<<grid, 256, 0, st>>>(n, static_cast<const int32_t *>(A->d_rp), d_perm, rp2);
    c->launches++;
    thrust::exclusive_scan(thrust::cuda::par.on(st), thrust::device_pointer_cast(rp2), thrust::device_pointer_cast(rp2) + n + 1,
                           thrust::device_pointer_cast(rp2));
    if (A->rp_bytes == 8)
        perm_fill_kernel<int64_t><<<grid, 256, 0, st>>>(n, static_cast<const int64_t *>(A->d_rp), A->d_col, A->d_val, d_perm, d_inv_perm, rp2,
                                                        static_cast<int64_t *>(B->d_rp), B->d_col, B->d_val);
    else
        perm_fill_kernel<int32_t><<<grid, 256, 0, st>>>(n, static_cast<const int32_t *>(A->d_rp), A->d_col, A->d_val, d_perm, d_inv_perm, rp2,
                                                        static_cast<int32_t *>(B->d_rp), B->d_col, B->d_val);
    c->launches++;
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(rp2);
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) {
        bis_set_error("bis_matrix_permute_symmetric: %s", cudaGetErrorString(e));
        bis_matrix_free(c, B);
        return 1;
    }
    bis_partition_set(c, n, 0, 0, n);
    *out = B;
    return bis_spmv_prepare(c, B);
}

// apply_vec_perm (smax_helpers.hpp:60-70): out[i] = in[perm[i]]; out must not alias in
extern "C" int bis_vector_permute(bis_context *c, double *out, const double *in, const int *d_perm, int64_t n) {
    BIS_REQUIRE(c && out && in && d_perm && out != in, "bis_vector_permute: bad argument");
    vec_perm_kernel<<<bis_blocks_for(n, 256, c->sm_count * 8), 256, 0, c->stream>>>(n, in, d_perm, out);
    BIS_LAUNCH_CHECK(c);
    return 0;
}

extern "C" int bis_index_alloc(bis_context *c, int64_t n, int **p) {
    BIS_REQUIRE(c && p && n >= 0, "bis_index_alloc: bad argument");
    BIS_CUDA(cudaSetDevice(c->device));
    BIS_CUDA(bis_cuda_malloc(p, sizeof(int) * (size_t)(n > 0 ? n : 1)));
    return 0;
}
extern "C" int bis_index_free(bis_context *c, int *p) {
    (void)c;
    if (p) cudaFree(p);
    return 0;
}
extern "C" int bis_index_download(bis_context *c, int32_t *dst, const int *src, int64_t n) {
    BIS_REQUIRE(c && (n == 0 || (dst && src)), "bis_index_download: null pointer");
    if (n == 0) return 0;
    BIS_CUDA(cudaMemcpyAsync(dst, src, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    BIS_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
