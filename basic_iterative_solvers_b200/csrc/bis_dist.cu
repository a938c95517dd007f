// bis_dist.cu -- row partitioning across GPUs: halo index lists and the NCCL
// halo exchange that precedes the boundary rows of every SpMV.
//
// New in the build: the reference is single-process.  Each rank
// owns a contiguous block of rows [row_begin, row_begin + n_rows).  Columns it
// owns are renumbered c - row_begin; every other column referenced by its rows
// becomes a "ghost" stored after the owned part, ordered by global id (so the
// ghosts of one owner are contiguous and the within-row order of the nonzeros,
// hence the summation order, is untouched: row-local kernels are bit-identical
// for every rank count).
//
// NCCL is used for exactly two things: this halo exchange (grouped
// ncclSend/ncclRecv on a dedicated communicator and stream, overlapped with the
// interior rows) and the allreduce of dot products (bis_context.cu).
#include "bis_device.cuh"

#include <thrust/copy.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include <algorithm>

namespace {

struct OutsideRange {
    int lo, hi;
    __host__ __device__ bool operator()(int c) const { return c < lo || c >= hi; }
};

__global__ void remap_cols_kernel(int64_t nnz, int *col, int lo, int hi, const int *ghost,
                                  int n_ghost, int n_owned) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz;
         k += (int64_t)gridDim.x * blockDim.x) {
        const int c = col[k];
        if (c >= lo && c < hi) {
            col[k] = c - lo;
        } else {
            int a = 0, b = n_ghost;   // lower_bound
            while (a < b) {
                int m = (a + b) >> 1;
                if (ghost[m] < c) a = m + 1;
                else b = m;
            }
            col[k] = n_owned + a;
        }
    }
}

template <typename RP>
__global__ void interior_range_kernel(int64_t n, const RP *rp, const int *col, int n_owned,
                                      int n_low_ghost, unsigned long long *ib, unsigned long long *ie) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x) {
        bool low = false, high = false;
        for (RP k = rp[r]; k < rp[r + 1]; ++k) {
            int c = col[k];
            if (c >= n_owned) {
                if (c - n_owned < n_low_ghost) low = true;
                else high = true;
            }
        }
        if (low) atomicMax(ib, (unsigned long long)(r + 1));
        if (high) atomicMin(ie, (unsigned long long)r);
    }
}

__global__ void pack_kernel(int64_t n_send, const int *idx, const double *x, double *buf) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_send;
         i += (int64_t)gridDim.x * blockDim.x)
        buf[i] = x[idx[i]];
}

__global__ void sub_offset_kernel(int64_t n, int *v, int off) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        v[i] -= off;
}

} // namespace

// A->d_col holds GLOBAL ids on entry (d_col_global == A->d_col) and local ids
// (ghosts >= n_cols) on return.
int bis_matrix_finalize_distributed(bis_context *c, bis_matrix *A, int *d_col_global) {
    BIS_REQUIRE(c->nranks > 1 && c->comm, "finalize_distributed needs a distributed context");
    (void)d_col_global;
    const int P = c->nranks, me = c->rank;
    A->distributed = true;
    cudaStream_t st = c->stream;
    auto pol = thrust::cuda::par.on(st);

    // (a) who owns what: allgather {row_begin, n_rows, nnz}
    int64_t *d_meta = nullptr;
    BIS_CUDA(cudaMalloc(&d_meta, sizeof(int64_t) * 3 * (size_t)(P + 1)));
    int64_t mine[3] = {A->row_begin, A->n_rows, A->nnz};
    BIS_CUDA(cudaMemcpyAsync(d_meta + 3 * P, mine, sizeof mine, cudaMemcpyHostToDevice, st));
    BIS_NCCL(ncclAllGather(d_meta + 3 * P, d_meta, 3, ncclInt64, c->comm, st));
    std::vector<int64_t> meta(3 * (size_t)P);
    BIS_CUDA(cudaMemcpyAsync(meta.data(), d_meta, sizeof(int64_t) * 3 * (size_t)P, cudaMemcpyDeviceToHost, st));
    BIS_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_meta);
    std::vector<int64_t> first(P + 1);
    int64_t nnz_g = 0;
    for (int p = 0; p < P; ++p) {
        first[p] = meta[3 * p];
        nnz_g += meta[3 * p + 2];
        if (p > 0)
            BIS_REQUIRE(meta[3 * (p - 1)] + meta[3 * (p - 1) + 1] == meta[3 * p],
                        "row blocks of ranks %d and %d are not contiguous", p - 1, p);
    }
    first[P] = meta[3 * (P - 1)] + meta[3 * (P - 1) + 1];
    BIS_REQUIRE(first[P] == A->n_rows_global, "row blocks cover %lld rows, matrix has %lld",
                (long long)first[P], (long long)A->n_rows_global);
    A->nnz_global = nnz_g;

    // (b) ghost columns: the distinct global ids outside [lo, hi), ascending
    const int lo = (int)A->row_begin, hi = (int)(A->row_begin + A->n_rows);
    int *d_tmp = nullptr;
    // upper bound on candidates is nnz; collect in chunks to bound the scratch
    const int64_t chunk = 1 << 26;
    std::vector<int> ghost_h;
    BIS_CUDA(cudaMalloc(&d_tmp, sizeof(int) * (size_t)std::min<int64_t>(std::max<int64_t>(A->nnz, 1), chunk)));
    for (int64_t off = 0; off < A->nnz; off += chunk) {
        const int64_t len = std::min<int64_t>(chunk, A->nnz - off);
        int *e = thrust::copy_if(pol, A->d_col + off, A->d_col + off + len, d_tmp, OutsideRange{lo, hi});
        thrust::sort(pol, d_tmp, e);
        e = thrust::unique(pol, d_tmp, e);
        const size_t cnt = (size_t)(e - d_tmp);
        const size_t old = ghost_h.size();
        ghost_h.resize(old + cnt);
        if (cnt)
            BIS_CUDA(cudaMemcpyAsync(ghost_h.data() + old, d_tmp, sizeof(int) * cnt, cudaMemcpyDeviceToHost, st));
        BIS_CUDA(cudaStreamSynchronize(st));
    }
    cudaFree(d_tmp);
    std::sort(ghost_h.begin(), ghost_h.end());
    ghost_h.erase(std::unique(ghost_h.begin(), ghost_h.end()), ghost_h.end());
    HaloPlan &h = A->halo;
    h.n_ghost = (int64_t)ghost_h.size();
    BIS_CUDA(cudaMalloc(&h.d_ghost_global, sizeof(int) * std::max<size_t>(ghost_h.size(), 1)));
    BIS_CUDA(cudaMalloc(&h.d_ghost, sizeof(double) * (std::max<size_t>(ghost_h.size(), 1) + 2)));
    if (!ghost_h.empty())
        BIS_CUDA(cudaMemcpyAsync(h.d_ghost_global, ghost_h.data(), sizeof(int) * ghost_h.size(), cudaMemcpyHostToDevice, st));

    // (c) renumber
    if (A->nnz) {
        remap_cols_kernel<<<c->sm_count * 8, 256, 0, st>>>(A->nnz, A->d_col, lo, hi, h.d_ghost_global,
                                                         (int)h.n_ghost, (int)A->n_rows);
        BIS_LAUNCH_CHECK(c);
    }

    // (d) segments of the ghost list per owner
    h.recv_off.assign(P + 1, 0);
    for (int p = 0; p <= P; ++p)
        h.recv_off[p] = std::lower_bound(ghost_h.begin(), ghost_h.end(), (int)std::min<int64_t>(first[p], INT32_MAX)) - ghost_h.begin();
    BIS_REQUIRE(h.recv_off[me + 1] == h.recv_off[me], "internal: own columns classified as ghosts");

    // (e) tell every owner how many of its rows we need: allgather the count rows
    int64_t *d_cnt = nullptr;
    BIS_CUDA(cudaMalloc(&d_cnt, sizeof(int64_t) * (size_t)P * (P + 1)));
    std::vector<int64_t> my_cnt(P);
    for (int p = 0; p < P; ++p) my_cnt[p] = h.recv_off[p + 1] - h.recv_off[p];
    BIS_CUDA(cudaMemcpyAsync(d_cnt + (size_t)P * P, my_cnt.data(), sizeof(int64_t) * P, cudaMemcpyHostToDevice, st));
    BIS_NCCL(ncclAllGather(d_cnt + (size_t)P * P, d_cnt, P, ncclInt64, c->comm, st));
    std::vector<int64_t> all_cnt((size_t)P * P);
    BIS_CUDA(cudaMemcpyAsync(all_cnt.data(), d_cnt, sizeof(int64_t) * (size_t)P * P, cudaMemcpyDeviceToHost, st));
    BIS_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_cnt);
    h.send_off.assign(P + 1, 0);
    for (int q = 0; q < P; ++q) h.send_off[q + 1] = h.send_off[q] + all_cnt[(size_t)q * P + me];
    h.n_send = h.send_off[P];
    BIS_CUDA(cudaMalloc(&h.d_send_idx, sizeof(int) * std::max<int64_t>(h.n_send, 1)));
    BIS_CUDA(cudaMalloc(&h.d_sendbuf, sizeof(double) * std::max<int64_t>(h.n_send, 1)));

    // (f) exchange the index lists (global ids), then make them local
    BIS_NCCL(ncclGroupStart());
    for (int p = 0; p < P; ++p) {
        if (p == me) continue;
        const int64_t nr = h.recv_off[p + 1] - h.recv_off[p];
        const int64_t ns = h.send_off[p + 1] - h.send_off[p];
        if (nr) BIS_NCCL(ncclSend(h.d_ghost_global + h.recv_off[p], (size_t)nr, ncclInt32, p, c->comm, st));
        if (ns) BIS_NCCL(ncclRecv(h.d_send_idx + h.send_off[p], (size_t)ns, ncclInt32, p, c->comm, st));
    }
    BIS_NCCL(ncclGroupEnd());
    if (h.n_send) {
        sub_offset_kernel<<<bis_blocks_for(h.n_send, 256, c->sm_count * 8), 256, 0, st>>>(h.n_send, h.d_send_idx, lo);
        BIS_LAUNCH_CHECK(c);
    }

    // (g) rows that touch no ghost: [interior_begin, interior_end)
    unsigned long long *d_range = nullptr;
    BIS_CUDA(cudaMalloc(&d_range, 2 * sizeof(unsigned long long)));
    unsigned long long init[2] = {0ull, (unsigned long long)A->n_rows};
    BIS_CUDA(cudaMemcpyAsync(d_range, init, sizeof init, cudaMemcpyHostToDevice, st));
    const int n_low = (int)h.recv_off[me];
    const int blocks = bis_blocks_for(A->n_rows, 256, c->sm_count * 8);
    if (A->rp_bytes == 8)
        interior_range_kernel<int64_t><<<blocks, 256, 0, st>>>(A->n_rows, static_cast<const int64_t *>(A->d_rp), A->d_col, (int)A->n_rows, n_low, d_range, d_range + 1);
    else
        interior_range_kernel<int32_t><<<blocks, 256, 0, st>>>(A->n_rows, static_cast<const int32_t *>(A->d_rp), A->d_col, (int)A->n_rows, n_low, d_range, d_range + 1);
    BIS_LAUNCH_CHECK(c);
    unsigned long long range[2];
    BIS_CUDA(cudaMemcpyAsync(range, d_range, sizeof range, cudaMemcpyDeviceToHost, st));
    BIS_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_range);
    h.interior_begin = (int64_t)range[0];
    h.interior_end = (int64_t)range[1];
    if (h.interior_end < h.interior_begin) h.interior_end = h.interior_begin;   // no interior
    return 0;
}

// Start the halo exchange of x on the comm stream (after everything already
// queued on the main stream, which produced x and consumed the old ghosts).
int bis_halo_exchange_begin(bis_context *c, const bis_matrix *A, const double *x) {
    const HaloPlan &h = A->halo;
    BIS_CUDA(cudaEventRecord(c->ev_main, c->stream));
    BIS_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_main, 0));
    if (h.n_send) {
        pack_kernel<<<bis_blocks_for(h.n_send, 256, c->sm_count * 4), 256, 0, c->comm_stream>>>(
            h.n_send, h.d_send_idx, x, h.d_sendbuf);
        BIS_LAUNCH_CHECK(c);
    }
    BIS_NCCL(ncclGroupStart());
    for (int p = 0; p < c->nranks; ++p) {
        if (p == c->rank) continue;
        const int64_t ns = h.send_off[p + 1] - h.send_off[p];
        const int64_t nr = h.recv_off[p + 1] - h.recv_off[p];
        if (ns) BIS_NCCL(ncclSend(h.d_sendbuf + h.send_off[p], (size_t)ns, ncclDouble, p, c->comm_halo, c->comm_stream));
        if (nr) BIS_NCCL(ncclRecv(h.d_ghost + h.recv_off[p], (size_t)nr, ncclDouble, p, c->comm_halo, c->comm_stream));
    }
    BIS_NCCL(ncclGroupEnd());
    BIS_CUDA(cudaEventRecord(c->ev_comm, c->comm_stream));
    return 0;
}

int bis_halo_exchange_end(bis_context *c, const bis_matrix *A) {
    (void)A;
    BIS_CUDA(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
    return 0;
}
