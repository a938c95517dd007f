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
#include <cstdlib>
#include <cstring>

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

// Peer-memory halo exchange ------------------------------------------------------------------
// The pack kernel of rank r stores x[send_idx[i]] straight into the ghost buffer of the rank
// that needs it (NVLink stores through the CUDA-IPC mapping), then publishes the exchange's
// epoch in that rank's bank; the receiver's boundary rows start behind halo_wait_kernel.
// Ghost buffers exist twice (epoch parity).  Before overwriting the copy last read by exchange
// e-2 the sender checks the receiver's ack: a rank acks epoch e (to every rank) at the start
// of its own pack e, which stream order places after all of its SpMVs before e.
struct HaloPeerArgs {
    int n_dst, n_ranks, me;
    int dst_rank[BIS_MAX_PEERS];
    int64_t seg_off[BIS_MAX_PEERS + 1];            // segments of the send list, one per destination
    double *dst[BIS_MAX_PEERS];                    // where that segment lands (destination's ghost copy of this parity)
    unsigned long long *dst_flag[BIS_MAX_PEERS];   // destination's bank: HALO_FLAG + me
    unsigned long long *ack_out[BIS_MAX_PEERS];    // rank p's bank: HALO_ACK + me
    const unsigned long long *ack_in;              // my bank: HALO_ACK
    unsigned long long epoch;
    unsigned int *ticket;
    int *errflag;
};

__global__ void __launch_bounds__(256) pack_peer_kernel(HaloPeerArgs a, const int *idx, const double *x) {
    __shared__ bool s_last;
    const int t = threadIdx.x;
    if (blockIdx.x == 0 && t < a.n_ranks && t != a.me)
        *reinterpret_cast<volatile unsigned long long *>(a.ack_out[t]) = a.epoch;
    if (t < a.n_dst && a.epoch >= 2) {
        const volatile unsigned long long *ack = a.ack_in + a.dst_rank[t];
        const unsigned long long t0 = bis_globaltimer();
        while (*ack + 1 < a.epoch) {
            if (bis_globaltimer() - t0 > BIS_PEER_TIMEOUT_NS) {
                atomicExch(a.errflag, 30 + a.dst_rank[t]);
                break;
            }
        }
    }
    __syncthreads();
    const int64_t n_send = a.seg_off[a.n_dst];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + t; i < n_send; i += (int64_t)gridDim.x * blockDim.x) {
        int d = 0;
        while (i >= a.seg_off[d + 1]) ++d;
        a.dst[d][i - a.seg_off[d]] = x[idx[i]];
    }
    __threadfence_system();
    __syncthreads();
    if (t == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    if (t < a.n_dst) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(a.dst_flag[t]) = a.epoch;
    }
    if (t == 0) *a.ticket = 0u;
}

struct HaloWaitArgs {
    int n_src;
    int src_rank[BIS_MAX_PEERS];
    const unsigned long long *flag;   // my bank: HALO_FLAG
    unsigned long long epoch;
    int *errflag;
};

__global__ void halo_wait_kernel(HaloWaitArgs a) {
    const int t = threadIdx.x;
    if (t < a.n_src) {
        const volatile unsigned long long *f = a.flag + a.src_rank[t];
        const unsigned long long t0 = bis_globaltimer();
        while (*f < a.epoch) {
            if (bis_globaltimer() - t0 > BIS_PEER_TIMEOUT_NS) {
                atomicExch(a.errflag, 40 + a.src_rank[t]);
                break;
            }
        }
        __threadfence_system();
    }
}

__global__ void sub_offset_kernel(int64_t n, int *v, int off) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        v[i] -= off;
}

} // namespace

// ---- peer-memory link ------------------------------------------------------------------------
// Collective.  Returns 0 with *ok = 1 when every rank mapped every other rank's buffer,
// 0 with *ok = 0 when some rank could not (the callers then keep NCCL), non-zero on a hard error.
static int peer_map_try(bis_context *c, void *mine, void **out, int *ok) {
    const int P = c->nranks, me = c->rank;
    cudaStream_t st = c->stream;
    *ok = 0;
    cudaIpcMemHandle_t hmine;
    int good = cudaIpcGetMemHandle(&hmine, mine) == cudaSuccess ? 1 : 0;
    if (!good) { cudaGetLastError(); memset(&hmine, 0, sizeof hmine); }
    unsigned char *d_h = nullptr;
    const size_t hb = sizeof(cudaIpcMemHandle_t);
    BIS_CUDA(bis_cuda_malloc(&d_h, hb * (size_t)(P + 1)));
    BIS_CUDA(cudaMemcpyAsync(d_h + hb * P, &hmine, hb, cudaMemcpyHostToDevice, st));
    BIS_NCCL(ncclAllGather(d_h + hb * P, d_h, hb, ncclUint8, c->comm, st));
    std::vector<cudaIpcMemHandle_t> all(P);
    BIS_CUDA(cudaMemcpyAsync(all.data(), d_h, hb * P, cudaMemcpyDeviceToHost, st));
    BIS_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_h);
    std::vector<void *> opened;
    for (int p = 0; p < P; ++p) {
        out[p] = nullptr;
        if (p == me) { out[p] = mine; continue; }
        if (!good) continue;
        void *q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            good = 0;
            continue;
        }
        out[p] = q;
        opened.push_back(q);
    }
    // agree: the link is used only if it is up everywhere
    int *d_ok = nullptr;
    BIS_CUDA(bis_cuda_malloc(&d_ok, sizeof(int)));
    BIS_CUDA(cudaMemcpyAsync(d_ok, &good, sizeof(int), cudaMemcpyHostToDevice, st));
    BIS_NCCL(ncclAllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, c->comm, st));
    int all_good = 0;
    BIS_CUDA(cudaMemcpyAsync(&all_good, d_ok, sizeof(int), cudaMemcpyDeviceToHost, st));
    BIS_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_ok);
    if (!all_good) {
        for (void *q : opened) cudaIpcCloseMemHandle(q);
        for (int p = 0; p < P; ++p) out[p] = (p == me) ? mine : nullptr;
        return 0;
    }
    for (void *q : opened) c->ipc_opened.push_back(q);
    *ok = 1;
    return 0;
}

int bis_peer_map(bis_context *c, void *mine, void **out) {
    int ok = 0;
    BIS_CHECK(peer_map_try(c, mine, out, &ok));
    return ok ? 0 : -1;
}

int bis_peer_link_setup(bis_context *c) {
    c->peer_on = 0;
    if (c->nranks <= 1 || c->nranks > BIS_MAX_PEERS) return 0;
    const char *env = getenv("BIS_P2P");
    if (env && env[0] == '0') return 0;   // must be set on every rank alike
    BIS_CUDA(bis_cuda_malloc(&c->d_bank, BIS_BANK_BYTES));
    BIS_CUDA(cudaMemset(c->d_bank, 0, BIS_BANK_BYTES));
    BIS_CUDA(bis_cuda_malloc(&c->d_pack_ticket, sizeof(unsigned int)));
    BIS_CUDA(cudaMemset(c->d_pack_ticket, 0, sizeof(unsigned int)));
    BIS_CUDA(bis_cuda_malloc(&c->d_waitstat, 4 * sizeof(unsigned long long)));
    BIS_CUDA(cudaMemset(c->d_waitstat, 0, 4 * sizeof(unsigned long long)));
    BIS_CUDA(cudaDeviceSynchronize());   // banks are zero before the allgather below lets anyone write
    void *banks[BIS_MAX_PEERS] = {};
    int ok = 0;
    BIS_CHECK(peer_map_try(c, c->d_bank, banks, &ok));
    if (!ok) return 0;
    for (int p = 0; p < c->nranks; ++p) c->peer_bank[p] = static_cast<double *>(banks[p]);
    c->peer_on = 1;
    return 0;
}

void bis_peer_link_teardown(bis_context *c) {
    for (void *q : c->ipc_opened) cudaIpcCloseMemHandle(q);
    c->ipc_opened.clear();
    cudaFree(c->d_bank);
    cudaFree(c->d_pack_ticket);
    cudaFree(c->d_waitstat);
    c->d_waitstat = nullptr;
    c->d_bank = nullptr;
    c->d_pack_ticket = nullptr;
    c->peer_on = 0;
}

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
    BIS_CUDA(bis_cuda_malloc(&d_meta, sizeof(int64_t) * 3 * (size_t)(P + 1)));
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
    // the partition-invariant sum needs EVERY rank's row block to be a union of virtual slabs
    if (c->part.n_global == A->n_rows_global) {
        bool all = BIS_NSLAB % P == 0;
        if (all) {
            int64_t vb[BIS_NSLAB + 1];
            int ch;
            bis_partition_rule(A->n_rows_global, 0, vb, &ch);
            for (int p = 0; p < P; ++p)
                if (meta[3 * p] != vb[p * (BIS_NSLAB / P)] || meta[3 * p] + meta[3 * p + 1] != vb[(p + 1) * (BIS_NSLAB / P)]) all = false;
        }
        if (!all) {
            c->part.invariant = false;
            c->part.n_slab = 1;
            c->part.slab_first = 0;
            c->part.slab_row[0] = 0;
            c->part.slab_row[1] = A->n_rows;
        }
    }
    BIS_REQUIRE(first[P] == A->n_rows_global, "row blocks cover %lld rows, matrix has %lld",
                (long long)first[P], (long long)A->n_rows_global);
    A->nnz_global = nnz_g;

    // (b) ghost columns: the distinct global ids outside [lo, hi), ascending
    const int lo = (int)A->row_begin, hi = (int)(A->row_begin + A->n_rows);
    int *d_tmp = nullptr;
    // upper bound on candidates is nnz; collect in chunks to bound the scratch
    const int64_t chunk = 1 << 26;
    std::vector<int> ghost_h;
    BIS_CUDA(bis_cuda_malloc(&d_tmp, sizeof(int) * (size_t)std::min<int64_t>(std::max<int64_t>(A->nnz, 1), chunk)));
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
    BIS_CUDA(bis_cuda_malloc(&h.d_ghost_global, sizeof(int) * std::max<size_t>(ghost_h.size(), 1)));
    h.ghost_stride = (int64_t)((ghost_h.size() + 2 + 15) & ~(size_t)15);
    {
        size_t bytes = sizeof(double) * 2 * (size_t)h.ghost_stride;
        bytes = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
        BIS_CUDA(bis_cuda_malloc(&h.d_ghost, bytes));
        BIS_CUDA(cudaMemsetAsync(h.d_ghost, 0, bytes, st));
    }
    h.cur_ghost = h.d_ghost;
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
    BIS_CUDA(bis_cuda_malloc(&d_cnt, sizeof(int64_t) * (size_t)P * (P + 1)));
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
    BIS_CUDA(bis_cuda_malloc(&h.d_send_idx, sizeof(int) * std::max<int64_t>(h.n_send, 1)));
    BIS_CUDA(bis_cuda_malloc(&h.d_sendbuf, sizeof(double) * std::max<int64_t>(h.n_send, 1)));

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
    BIS_CUDA(bis_cuda_malloc(&d_range, 2 * sizeof(unsigned long long)));
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

    // (h) peer-memory transport: map every rank's ghost buffer, learn where my values land in it
    h.peer_ready = false;
    if (c->peer_on) {
        BIS_CUDA(cudaStreamSynchronize(st));
        void *pg[BIS_MAX_PEERS] = {};
        int ok = 0;
        BIS_CHECK(peer_map_try(c, h.d_ghost, pg, &ok));
        int64_t *d_ro = nullptr;
        BIS_CUDA(bis_cuda_malloc(&d_ro, sizeof(int64_t) * (size_t)(P + 1) * (P + 2)));
        std::vector<int64_t> mine_ro(P + 2);
        for (int p = 0; p <= P; ++p) mine_ro[p] = h.recv_off[p];
        mine_ro[P + 1] = ok ? h.ghost_stride : -1;
        BIS_CUDA(cudaMemcpyAsync(d_ro + (size_t)P * (P + 2), mine_ro.data(), sizeof(int64_t) * (P + 2), cudaMemcpyHostToDevice, st));
        BIS_NCCL(ncclAllGather(d_ro + (size_t)P * (P + 2), d_ro, P + 2, ncclInt64, c->comm, st));
        std::vector<int64_t> all_ro((size_t)P * (P + 2));
        BIS_CUDA(cudaMemcpyAsync(all_ro.data(), d_ro, sizeof(int64_t) * all_ro.size(), cudaMemcpyDeviceToHost, st));
        BIS_CUDA(cudaStreamSynchronize(st));
        cudaFree(d_ro);
        if (ok) {
            h.peer_recv_off.assign(P, 0);
            h.peer_stride.assign(P, 0);
            for (int p = 0; p < P; ++p) {
                h.peer_ghost[p] = static_cast<double *>(pg[p]);
                h.peer_recv_off[p] = all_ro[(size_t)p * (P + 2) + me];
                h.peer_stride[p] = all_ro[(size_t)p * (P + 2) + P + 1];
            }
            h.peer_ready = true;
        }
    }
    return 0;
}

// Start the halo exchange of x.  Peer-memory transport: one pack kernel on the main stream that
// stores into the neighbours' ghost buffers.  NCCL transport: pack + grouped send/recv on the
// comm stream (after everything already queued on the main stream, which produced x and
// consumed the old ghosts).
int bis_halo_exchange_begin(bis_context *c, const bis_matrix *A, const double *x) {
    const HaloPlan &h = A->halo;
    if (c->peer_on && c->opt_dist_p2p && h.peer_ready) {
        const unsigned long long e = ++c->halo_epoch;
        const int par = (int)(e & 1ull);
        h.cur_ghost = h.d_ghost + (size_t)par * h.ghost_stride;
        HaloPeerArgs a;
        a.n_dst = 0; a.n_ranks = c->nranks; a.me = c->rank;
        a.seg_off[0] = 0;
        unsigned long long *mybank = reinterpret_cast<unsigned long long *>(c->d_bank);
        for (int p = 0; p < c->nranks; ++p) {
            unsigned long long *pb = reinterpret_cast<unsigned long long *>(c->peer_bank[p]);
            a.ack_out[p] = pb + BIS_BANK_HALO_ACK + c->rank;
            const int64_t ns = h.send_off[p + 1] - h.send_off[p];
            if (p == c->rank || ns == 0) continue;
            const int d = a.n_dst++;
            a.dst_rank[d] = p;
            a.seg_off[d] = h.send_off[p];      // the send list is ordered by destination rank
            a.seg_off[d + 1] = h.send_off[p + 1];
            a.dst[d] = h.peer_ghost[p] + (size_t)par * h.peer_stride[p] + h.peer_recv_off[p];
            a.dst_flag[d] = pb + BIS_BANK_HALO_FLAG + c->rank;
        }
        for (int d = a.n_dst; d < BIS_MAX_PEERS; ++d) {
            a.dst_rank[d] = 0; a.dst[d] = nullptr; a.dst_flag[d] = nullptr;
            a.seg_off[d + 1] = a.seg_off[a.n_dst];
        }
        for (int p = c->nranks; p < BIS_MAX_PEERS; ++p) a.ack_out[p] = nullptr;
        a.ack_in = mybank + BIS_BANK_HALO_ACK;
        a.epoch = e;
        a.ticket = c->d_pack_ticket;
        a.errflag = c->d_errflag;
        pack_peer_kernel<<<bis_blocks_for(h.n_send, 256 * 4, c->sm_count * 2), 256, 0, c->stream>>>(a, h.d_send_idx, x);
        BIS_LAUNCH_CHECK(c);
        return 0;
    }
    h.cur_ghost = h.d_ghost;
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

int bis_halo_fuse_args(bis_context *c, const bis_matrix *A, HaloFuse *hf) {
    const HaloPlan &h = A->halo;
    BIS_REQUIRE(c->peer_on && c->opt_dist_p2p && h.peer_ready, "internal: fused halo exchange without the peer-memory link");
    const unsigned long long e = ++c->halo_epoch;
    const int par = (int)(e & 1ull);
    h.cur_ghost = h.d_ghost + (size_t)par * h.ghost_stride;
    HaloFuse &a = *hf;
    a.n_dst = 0; a.n_src = 0; a.n_ranks = c->nranks; a.me = c->rank;
    a.seg_off[0] = 0;
    unsigned long long *mybank = reinterpret_cast<unsigned long long *>(c->d_bank);
    for (int p = 0; p < BIS_MAX_PEERS; ++p) {
        a.ack_out[p] = nullptr; a.dst[p] = nullptr; a.dst_flag[p] = nullptr; a.dst_rank[p] = 0; a.src_rank[p] = 0;
    }
    for (int p = 0; p < c->nranks; ++p) {
        unsigned long long *pb = reinterpret_cast<unsigned long long *>(c->peer_bank[p]);
        a.ack_out[p] = pb + BIS_BANK_HALO_ACK + c->rank;
        if (p == c->rank) continue;
        if (h.recv_off[p + 1] > h.recv_off[p]) a.src_rank[a.n_src++] = p;
        const int64_t ns = h.send_off[p + 1] - h.send_off[p];
        if (ns == 0) continue;
        const int d = a.n_dst++;
        a.dst_rank[d] = p;
        a.seg_off[d] = h.send_off[p];
        a.seg_off[d + 1] = h.send_off[p + 1];
        a.dst[d] = h.peer_ghost[p] + (size_t)par * h.peer_stride[p] + h.peer_recv_off[p];
        a.dst_flag[d] = pb + BIS_BANK_HALO_FLAG + c->rank;
    }
    for (int d = a.n_dst; d < BIS_MAX_PEERS; ++d) a.seg_off[d + 1] = a.seg_off[a.n_dst];
    a.ack_in = mybank + BIS_BANK_HALO_ACK;
    a.flag_in = mybank + BIS_BANK_HALO_FLAG;
    a.send_idx = h.d_send_idx;
    a.ticket = c->d_pack_ticket;
    a.epoch = e;
    a.errflag = c->d_errflag;
    a.waitstat = c->d_waitstat;
    return 0;
}

// Aligns the ranks' streams: everything enqueued after it starts only when every rank's stream has got here
// (one 4-byte NCCL allreduce; a measurement aid -- a timed region opened right after it does not count the time a
// late rank took to arrive -- not part of any solver path).
extern "C" int bis_dist_stream_barrier(bis_context *c) {
    BIS_REQUIRE(c, "null context");
    if (c->nranks <= 1) return 0;
    BIS_CUDA(cudaSetDevice(c->device));
    if (!c->d_barrier_word) {
        BIS_CUDA(bis_cuda_malloc(&c->d_barrier_word, sizeof(int)));
        BIS_CUDA(cudaMemsetAsync(c->d_barrier_word, 0, sizeof(int), c->stream));
    }
    BIS_NCCL(ncclAllReduce(c->d_barrier_word, c->d_barrier_word, 1, ncclInt32, ncclSum, c->comm, c->stream));
    return 0;
}

// In-kernel wait accounting (ns): out[0] total time the finalising blocks of reductions waited for the
// other ranks' records, out[1] reductions, out[2] time CTA 0's producer warp of the fused SpMV waited
// for the senders' halo flags, out[3] exchanges.
extern "C" int bis_dist_wait_read(bis_context *c, double out[4], int reset) {
    BIS_REQUIRE(c && out, "null argument");
    for (int i = 0; i < 4; ++i) out[i] = 0.0;
    if (!c->d_waitstat) return 0;
    unsigned long long h[4];
    BIS_CUDA(cudaMemcpyAsync(h, c->d_waitstat, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    BIS_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 4; ++i) out[i] = (double)h[i];
    if (reset) BIS_CUDA(cudaMemsetAsync(c->d_waitstat, 0, sizeof h, c->stream));
    return 0;
}

// Everything queued on the main stream after this sees the ghosts of the current exchange.
int bis_halo_exchange_end(bis_context *c, const bis_matrix *A) {
    const HaloPlan &h = A->halo;
    if (c->peer_on && c->opt_dist_p2p && h.peer_ready) {
        HaloWaitArgs w;
        w.n_src = 0;
        for (int p = 0; p < c->nranks; ++p)
            if (p != c->rank && h.recv_off[p + 1] > h.recv_off[p]) w.src_rank[w.n_src++] = p;
        for (int d = w.n_src; d < BIS_MAX_PEERS; ++d) w.src_rank[d] = 0;
        if (w.n_src == 0) return 0;
        w.flag = reinterpret_cast<const unsigned long long *>(c->d_bank) + BIS_BANK_HALO_FLAG;
        w.epoch = c->halo_epoch;
        w.errflag = c->d_errflag;
        halo_wait_kernel<<<1, 32, 0, c->stream>>>(w);
        BIS_LAUNCH_CHECK(c);
        return 0;
    }
    BIS_CUDA(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
    return 0;
}
