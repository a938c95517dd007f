// bis_sptrsv_chain.cuh -- triangular solve, variant 4 ("chains"): opt-in (trsv_variant = 4), see DESIGN.md 3.2.
//
// The dataflow solve of bis_sptrsv.cu pays one L2 hop per LEVEL (7n-6 of them for HPCG-n).  Most of
// those hops connect a row to the row just before it: in a stencil factor row r reads row r-1 (the
// x-neighbour), so the rows of an x-line form a CHAIN that is sequential by nature.  Here a lane owns a
// whole chain (a maximal run of consecutive rows each of which reads its predecessor); the 32 chains of
// a warp advance in lockstep, skewed by their levels (step = level - first level of the warp), so every
// operand produced inside the warp at most seven steps earlier -- the chain predecessor first of all --
// comes from a small shared-memory ring, and only operands of OTHER warps travel
// through L2 -- and those only need the producing warp to run a hop AHEAD, not a hop per step.
//
// Storage (built once per factor, on the device): per warp and step one record holding the 32 lanes'
// rows side by side (sliced ELL): val[K][32], code[K][32], row[32].  A record is one contiguous block:
// a bulk copy (cp.async.bulk, SASS UBLKCP) brings it into shared memory, NST records ahead.  code:
//   <= CODE_RING0   ring entry CODE_RING0 - code = (producer step & 7) * 32 + producer lane (the chain
//                   predecessor is the most frequent of these)
//   >= 0            position in the working vector w (laid out record-major: the 32 results of a warp
//                   step are one 256-byte store, and the operands that 32 neighbouring chains read from
//                   another warp are 32 neighbouring words)
//   CODE_NONE       padding
// Products are added in the row's storage order with separate roundings: bit-identical to the
// reference's loop, as in every other variant.  Rows are handed out warp by warp in chain order through
// a ticket: every operand of another warp belongs to a warp that already runs (chains are contiguous
// row ranges in dependency-compatible order), so there is no deadlock and no co-residency assumption.
#pragma once

#include "bis_device.cuh"
#include "bis_spmv_tma.cuh"

#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>

namespace chain {

constexpr int CODE_NONE = INT32_MIN;
constexpr int CODE_RING0 = -2;
constexpr int NST = 4;                 // records in flight per warp
constexpr int WARPS = 8;               // warps (chain groups) per block
constexpr unsigned long long SENT = 0xFFF87E5E7E5E7E5EULL;

__host__ __device__ inline size_t rec_bytes(int K) { return (size_t)K * 256 + (size_t)(K + 1) * 128; }

// ---- build ------------------------------------------------------------------------------------------
// position p <-> row: lower factors walk the rows upwards, upper factors downwards
__device__ __forceinline__ int row_of_pos(int64_t p, int64_t n, int upper) { return (int)(upper ? n - 1 - p : p); }

template <typename RP>
__global__ void head_kernel(int64_t n, const RP *rp, const int *col, int upper, int *head) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        int h = 1;
        if (p > 0) {
            const int r = row_of_pos(p, n, upper), target = row_of_pos(p - 1, n, upper);
            for (RP k = rp[r]; k < rp[r + 1]; ++k)
                if (col[k] == target) h = 0;
        }
        head[p] = h;
    }
}

__global__ void chain_start_kernel(int64_t n, const int *head, const int *cid, int *chain_start, int n_chains) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p <= n; p += (int64_t)gridDim.x * blockDim.x) {
        if (p == n) chain_start[n_chains] = (int)n;
        else if (head[p]) chain_start[cid[p] - 1] = (int)p;   // cid is the inclusive scan of head
    }
}

__global__ void group_steps_kernel(int n_groups, int n_chains, int64_t n, int upper, const int *chain_start,
                                   const int *level, int *lmin, long long *steps) {
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_groups; w += gridDim.x * blockDim.x) {
        int lo = 0x7fffffff, hi = -1;
        for (int c = w * 32; c < min(n_chains, w * 32 + 32); ++c) {
            lo = min(lo, level[row_of_pos(chain_start[c], n, upper)]);
            hi = max(hi, level[row_of_pos(chain_start[c + 1] - 1, n, upper)]);
        }
        lmin[w] = lo;
        steps[w] = (long long)(hi - lo + 1);
    }
}

__global__ void rec_init_kernel(long long n_recs, int K, unsigned char *recs) {
    const size_t rb = rec_bytes(K);
    const long long total = n_recs * (long long)(K + 1) * 32;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long g = i / ((K + 1) * 32);
        const int e = (int)(i % ((K + 1) * 32));
        int *meta = reinterpret_cast<int *>(recs + (size_t)g * rb + (size_t)K * 256);
        meta[e] = e >= K * 32 ? -1 : CODE_NONE;          // row index -1: idle lane
        if (e < K * 32) reinterpret_cast<double *>(recs + (size_t)g * rb)[e] = 0.0;
    }
}

__global__ void pos_kernel(int64_t n, int upper, const int *cid, const int *level, const int *lmin,
                           const long long *slice_off, int *pos) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int r = row_of_pos(p, n, upper);
        const int c = cid[p] - 1, w = c >> 5, lane = c & 31;
        pos[r] = (int)((slice_off[w] + (level[r] - lmin[w])) * 32 + lane);
    }
}

template <typename RP>
__global__ void rec_fill_kernel(int64_t n, const RP *rp, const int *col, const double *val, int upper, int K,
                                const int *cid, const int *level, const int *lmin, const int *pos,
                                unsigned char *recs) {
    const size_t rb = rec_bytes(K);
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int r = row_of_pos(p, n, upper);
        const int c = cid[p] - 1, w = c >> 5, lane = c & 31;
        const int s = level[r] - lmin[w];
        const int g = pos[r] >> 5;
        double *rv = reinterpret_cast<double *>(recs + (size_t)g * rb);
        int *meta = reinterpret_cast<int *>(recs + (size_t)g * rb + (size_t)K * 256);
        meta[K * 32 + lane] = r;
        int k = 0;
        for (RP q = rp[r]; q < rp[r + 1]; ++q, ++k) {     // storage order is the summation order
            const int cc = col[q];
            const int64_t pc = upper ? n - 1 - cc : cc;
            const int c2 = cid[pc] - 1;
            int code = pos[cc];
            if ((c2 >> 5) == w) {   // produced by this warp (the chain predecessor included): ring if recent enough
                const int sc = level[cc] - lmin[w];
                const int age = s - sc;
                if (age >= 1 && age <= 7) code = CODE_RING0 - (((sc & 7) << 5) | (c2 & 31));
            }
            rv[k * 32 + lane] = val[q];
            meta[k * 32 + lane] = code;
        }
    }
}

// ---- solve ------------------------------------------------------------------------------------------
struct Args {
    int n_groups;
    int K;
    const unsigned char *recs;
    const long long *slice_off;     // [n_groups + 1]
    unsigned int *ticket;
    int *errflag;
    double *x;
    const double *D;
    const double *b;
};

__device__ __forceinline__ unsigned long long ld_relaxed(const double *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void fill_sentinel_kernel(long long n, double *w) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        reinterpret_cast<unsigned long long *>(w)[i] = SENT;
}

constexpr int KMAX = 16;

// One warp step.  cur: this step's (row, b, D, operands of other warps), requested during the previous
// step; nxt: the same for step s+1, requested here.  Everything of a step is loaded before anything is
// consumed and operands are steered by predicated loads into one register per operand, not by
// selects: with one or two warps per scheduler a step is a chain of instruction latencies, and the
// first version of this kernel spent 850 instructions per step.
struct StepOps {
    unsigned long long x[KMAX];
    int row;
    double b, d;
};

constexpr unsigned int SENT_HI = (unsigned int)(SENT >> 32);

template <int KT>
__device__ __forceinline__ void request(const Args &a, const double *w, const unsigned char *rec, int lane, StepOps &o) {
    const int *meta = reinterpret_cast<const int *>(rec + (size_t)KT * 256);
    int code[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) code[k] = meta[k * 32 + lane];
    o.row = meta[KT * 32 + lane];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        o.x[k] = 0ull;
        if (code[k] >= 0) o.x[k] = ld_relaxed(w + code[k]);
    }
    o.b = 0.0;
    o.d = 1.0;
    if (o.row >= 0) {
        o.b = a.b[o.row];
        o.d = a.D[o.row];
    }
}

template <int KT>
__global__ void __launch_bounds__(WARPS * 32) chain_kernel(Args a, double *w) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned int s_chunk;
    if (threadIdx.x == 0) s_chunk = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wg = (int)s_chunk * WARPS + warp;
    if (wg >= a.n_groups) return;
    constexpr size_t rb = (size_t)KT * 256 + (size_t)(KT + 1) * 128;
    // per warp: NST records, the ring ([8 steps][32 lanes]: conflict-free), the barriers
    constexpr size_t per_warp = NST * rb + 32 * 8 * sizeof(double) + 64;
    unsigned char *base = smem + (size_t)warp * per_warp;
    double *ring = reinterpret_cast<double *>(base + NST * rb);
    uint64_t *full = reinterpret_cast<uint64_t *>(base + NST * rb + 32 * 8 * sizeof(double));
    const long long g0 = a.slice_off[wg];
    const int S = (int)(a.slice_off[wg + 1] - g0);
    const uint64_t pol = tma::policy_evict_first();
    if (lane == 0) {
        for (int i = 0; i < NST; ++i) tma::mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto fetch = [&](int s) {   // lane 0: record of step s into its stage
        const int st = s % NST;
        tma::fence_reads_before_bulk_write();
        tma::mbar_expect_tx(&full[st], (uint32_t)rb);
        tma::bulk_g2s(base + (size_t)st * rb, a.recs + (size_t)(g0 + s) * rb, (uint32_t)rb, &full[st], pol);
    };
    if (lane == 0)
        for (int s = 0; s < NST && s < S; ++s) fetch(s);

    auto step = [&](int s, StepOps &cur, StepOps &nxt) {
        const int st = s % NST;
        const unsigned char *rec = base + (size_t)st * rb;
        if (s + 1 < S) {
            const int sn = (s + 1) % NST;
            tma::mbar_wait(&full[sn], (uint32_t)(((s + 1) / NST) & 1));
            request<KT>(a, w, base + (size_t)sn * rb, lane, nxt);
        }
        const double *rv = reinterpret_cast<const double *>(rec);
        const int *meta = reinterpret_cast<const int *>(rec + (size_t)KT * 256);
        int code[KT];
        double av[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            code[k] = meta[k * 32 + lane];
            av[k] = rv[k * 32 + lane];
        }
        // operands made by this warp: from the ring, into the operand's own register
        // (unconditional loads from a clamped index plus a select: as `if` bodies ptxas turned them into
        // fourteen branch / reconvergence regions)
        double rg[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const unsigned int e = (unsigned int)(CODE_RING0 - code[k]);   // 0..255 for a ring operand
            rg[k] = ring[e < 256u ? e : 0u];
        }
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if ((unsigned int)(CODE_RING0 - code[k]) < 256u) cur.x[k] = (unsigned long long)__double_as_longlong(rg[k]);
        // operands of other warps that were not there yet (warp-uniform loop; rare once the producing
        // warp runs a hop ahead).  The sentinel's high word cannot occur in a result.
        bool missing = false;
#pragma unroll
        for (int k = 0; k < KT; ++k) missing |= (unsigned int)(cur.x[k] >> 32) == SENT_HI;
        if (__any_sync(0xffffffffu, missing)) {
            unsigned int miss = 0u, spins = 0;
            unsigned long long t_wd = 0;
#pragma unroll
            for (int k = 0; k < KT; ++k) miss |= ((unsigned int)(cur.x[k] >> 32) == SENT_HI ? 1u : 0u) << k;
            do {
#pragma unroll
                for (int k = 0; k < KT; ++k)
                    if (miss & (1u << k)) cur.x[k] = ld_relaxed(w + code[k]);
#pragma unroll
                for (int k = 0; k < KT; ++k)
                    if ((unsigned int)(cur.x[k] >> 32) != SENT_HI) miss &= ~(1u << k);
                bool give_up = false;
                if ((++spins & 1023u) == 0) {
                    if (t_wd == 0) t_wd = bis_globaltimer();
                    give_up = *reinterpret_cast<volatile int *>(a.errflag) != 0 || bis_globaltimer() - t_wd > 3000000000ull;
                }
                if (__any_sync(0xffffffffu, give_up)) {
                    atomicExch(a.errflag, 5);
                    miss = 0u;
                }
            } while (!__all_sync(0xffffffffu, miss == 0u));
        }
        // the row: all products first (independent), then the sum in storage order, separately rounded
        double pr[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) pr[k] = mul_rn(av[k], __longlong_as_double((long long)cur.x[k]));
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (code[k] != CODE_NONE) sum = add_rn(sum, pr[k]);     // padding sits behind the row's last nonzero
        if (cur.row >= 0) {
            const double r = div_rn(sub_rn(cur.b, sum), cur.d);
            ring[((s & 7) << 5) | lane] = r;
            __stcg(w + (g0 + s) * 32 + lane, r);
            a.x[cur.row] = r;
        }
        __syncwarp();   // ring written (read by the next steps), this stage's record no longer needed
        if (lane == 0 && s + NST < S) fetch(s + NST);
    };
    StepOps oa, ob;
    tma::mbar_wait(&full[0], 0u);
    request<KT>(a, w, base, lane, oa);
    for (int s = 0; s < S; s += 2) {
        step(s, oa, ob);
        if (s + 1 < S) step(s + 1, ob, oa);
    }
}

}  // namespace chain
