// bis_sptrsv_wave.cuh -- triangular solve, variant 5 ("stencil wavefront"): chosen automatically for factors of
// matrices on a structured grid (<= 27-point stencils in natural ordering: HPCG, 2-D 9-point, ...) when the cost
// model in bis_sptrsv.cu expects it to beat the dataflow solve, see DESIGN.md 3.2.  Same row semantics and the
// same bits as native_sptrsv / native_bsptrsv (kernels.hpp:54-76, 88-107): products separately rounded and added
// in storage order.
//
// Why: the general dataflow solve (bis_sptrsv.cu) pays one L2 round trip per LEVEL, and HPCG-n has 7n-6
// of them (1.28 us per level measured, 2.3 ms per sweep at n = 256).  In a stencil factor the rows of an
// x-line form a chain (row r reads r-1), and everything else a row reads lies in the neighbouring lines
// y-1 (same plane) and y-1, y, y+1 (previous plane).  So:
//   * a LANE owns an x-line and walks it one row per step; 32 consecutive lines of a plane form a block, lane j
//     running 2j steps behind lane 0 (row (x,y) needs (x+1,y-1): the level function is x+2y+4z); a CTA owns a
//     whole plane (up to 8 blocks, block w running 64 steps behind block w-1), two warps per block;
//   * what a row needs from its own plane are the results of at most three steps ago: a shared-memory ring
//     indexed by STEP, so that for every lane the operand of stencil slot (dx,dy) sits in ring row
//     (step + dx + 2 dy) -- no shuffles, no selects;
//   * the values of the previous plane (same lines: 256 contiguous bytes per block and step) come from the working
//     vector in L2: an asynchronous copy (cp.async.cg) requested GA steps before they enter the ring, checked
//     against the "not ready" pattern there (the value is its own ready flag, as in the dataflow solve) and polled
//     only if the producer has not got that far: a producer only has to run a few steps AHEAD, the L2 latency is
//     off the critical path, and the dependency chain of a step is register/shared-memory only;
//   * matrix values arrive as one bulk copy (cp.async.bulk, SASS UBLKCP) per block and step from a record
//     layout built once per factor ([group][step][slot][lane], zeros where a neighbour does not exist),
//     b and D by 8-byte cp.async eight steps ahead.
// CTAs take planes in order from a ticket: everything a plane waits for belongs to an earlier ticket, i.e. to a
// CTA that already runs -- no co-residency assumption, no deadlock.
//
// Absent neighbours are stored as value +0.0; their products (+-0.0) do not change the running sum
// (it starts at +0.0 and can never be -0.0), so the result is bit-identical to skipping them -- as long
// as the solution is finite (0 * inf would spread a NaN differently from the reference; a diverged
// solve is stopped by the harness on its NaN norm either way).
#pragma once

#include "bis_device.cuh"
#include "bis_spmv_tma.cuh"

namespace wave {

constexpr int K = 13;          // slots of a lower (<= 27-point) stencil, lexicographic (dz, dy, dx)
constexpr int RING = 8;        // ring rows (steps); the step loop is unrolled by it
#ifndef WAVE_NST
#define WAVE_NST 4
#endif
#ifndef WAVE_PF
#define WAVE_PF 24
#endif
constexpr int NST = WAVE_NST;  // matrix records in flight per warp (shared memory); divides RING
constexpr int PF = WAVE_PF;    // steps a record is prefetched into L2 ahead of its bulk copy (0: off)
#ifndef WAVE_GA
#define WAVE_GA 2
#endif
#ifndef WAVE_BD_REG
#define WAVE_BD_REG 0          // b and D by plain loads two own steps ahead (registers) instead of cp.async rings, see solve
#endif
constexpr int GA = WAVE_GA;    // steps between the request of a value of the previous plane and its entry into a ring
constexpr int BD = 8;          // steps b and D are requested ahead
constexpr int PH = 2;          // warps per 32-line block; they take its steps in turn (see solve / prepare)
static_assert(GA % PH == 0 && GA >= PH && RING % PH == 0 && BD % PH == 0 && BD - PH - 1 >= 0, "a request is consumed by the warp that made it");
constexpr int MAX_WARPS = 8;   // 32-line blocks per plane a CTA can hold (ny <= 256)
constexpr int REC_DOUBLES = K * 32;
constexpr unsigned long long SENT = 0xFFF87E5E7E5E7E5EULL;
constexpr unsigned long long WATCHDOG_NS = 60000000000ull;


struct Grid {
    int nx, ny, nz;            // lines of nx rows, planes of ny lines
    int W;                     // warps (32-line blocks) per plane
    int S;                     // steps per group: nx + 62
    long long n;               // rows
    int upper;                 // 1: backward solve; all coordinates are those of position p = n-1-row
};

__host__ __device__ inline long long n_groups(const Grid &g) { return (long long)g.nz * g.W; }

// ---- build ------------------------------------------------------------------------------------------
// slot of the entry (row, col) or -1 when it does not fit the stencil
__device__ __forceinline__ int slot_of(const Grid &g, long long p, long long pc) {
    const long long P = (long long)g.nx * g.ny;
    const int x = (int)(p % g.nx), y = (int)((p / g.nx) % g.ny);
    const long long z = p / P;
    const int xc = (int)(pc % g.nx), yc = (int)((pc / g.nx) % g.ny);
    const long long zc = pc / P;
    const int dx = xc - x, dy = yc - y;
    const long long dz = zc - z;
    if (dx < -1 || dx > 1 || dy < -1 || dy > 1 || dz < -1 || dz > 0) return -1;
    const int k = (((int)dz + 1) * 3 + (dy + 1)) * 3 + (dx + 1);
    return k < K ? k : -1;     // k == 13 is the diagonal, above it the other triangle
}

// status[0] != 0: some entry does not fit; also requires ascending columns inside a row (then the slot
// order is the storage order, i.e. the reference's summation order)
template <typename RP>
__global__ void validate_kernel(Grid g, const RP *rp, const int *col, int *status) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < g.n; r += (long long)gridDim.x * blockDim.x) {
        const long long p = g.upper ? g.n - 1 - r : r;
        int prev = -1;
        bool bad = false;
        for (RP q = rp[r]; q < rp[r + 1]; ++q) {
            const int c = col[q];
            if (c <= prev) bad = true;
            prev = c;
            const long long pc = g.upper ? g.n - 1 - c : c;
            if (pc < 0 || pc >= p || slot_of(g, p, pc) < 0) bad = true;
        }
        if (bad) atomicExch(status, 1);
    }
}

template <typename RP>
__global__ void fill_kernel(Grid g, const RP *rp, const int *col, const double *val, double *rec) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < g.n; r += (long long)gridDim.x * blockDim.x) {
        const long long p = g.upper ? g.n - 1 - r : r;
        const long long P = (long long)g.nx * g.ny;
        const int x = (int)(p % g.nx), y = (int)((p / g.nx) % g.ny);
        const long long z = p / P;
        const int lane = y & 31;
        const long long grp = z * g.W + (y >> 5);
        const int ls = x + 2 * lane;
        double *base = rec + ((size_t)grp * g.S + ls) * REC_DOUBLES + lane;
        for (RP q = rp[r]; q < rp[r + 1]; ++q) {
            const int c = col[q];
            const long long pc = g.upper ? g.n - 1 - c : c;
            const int k = slot_of(g, p, pc);
            // an upper factor's storage order (ascending column) is DESCENDING in position space: the kernel
            // always adds planes 0..12, so its planes are stored mirrored
            base[(size_t)(g.upper ? K - 1 - k : k) * 32] = val[q];
        }
    }
}

__global__ void fill_u64_kernel(long long n, unsigned long long *p, unsigned long long v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// distinct |row - col| of a triangular factor (a stencil has at most 13): tab[0] = count, tab[1..]
constexpr int OFFS_CAP = 32;
template <typename RP>
__global__ void offsets_kernel(long long n, const RP *rp, const int *col, int *tab) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x)
        for (RP q = rp[r]; q < rp[r + 1]; ++q) {
            const long long d = r - col[q];
            const int o = (int)(d < 0 ? -d : d);
            bool found = false;
            for (int guard = 0; guard < 4 * OFFS_CAP && !found; ++guard) {
                const int cnt = *reinterpret_cast<volatile int *>(tab);
                for (int i = 0; i < cnt && i < OFFS_CAP; ++i)
                    if (*reinterpret_cast<volatile int *>(tab + 1 + i) == o) found = true;
                if (found || cnt >= OFFS_CAP) break;
                // append: claim slot cnt
                if (atomicCAS(tab + 1 + cnt, 0, o) == 0) {
                    atomicMax(tab, cnt + 1);
                    found = true;
                } else if (*reinterpret_cast<volatile int *>(tab + 1 + cnt) == o) {
                    atomicMax(tab, cnt + 1);
                    found = true;
                } else {
                    atomicMax(tab, cnt + 1);   // somebody else's offset sits there: look again
                }
            }
            if (!found) atomicExch(tab + 1 + OFFS_CAP, 1);   // more distinct offsets than a stencil has
        }
}

// ---- solve ------------------------------------------------------------------------------------------
struct Args {
    Grid g;
    const double *rec;
    double *w;                 // working vector of this solve, [group][step][lane]
    double *w_clean;           // its twin: marked "not ready" here for the next solve
    unsigned int *ticket;
    int *errflag;
    double *x;
    const double *D;
    const double *b;
    int post_mul_d;
    unsigned int backoff_ns;   // a plane fed through L2 that had to poll falls back by this much (clusters only, see prepare)
#ifdef BIS_PERF_DEBUG
    int dbg;                   // perf experiments (results invalid): 1 no record wait, 2 no b/D wait, 4 no x store,
                               // 8 no L2 operand requests, 16 no w stores, 32 no record copies
    unsigned long long *stamps;   // [2 * nz] globaltimer at the first / after the last step of every plane, or nullptr
#endif
};
#ifdef BIS_PERF_DEBUG
#define WAVE_DBG(a, bit) ((a).dbg & (bit))
#else
#define WAVE_DBG(a, bit) 0
#endif

__device__ __forceinline__ void prefetch_l2_bulk(const void *gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

__device__ __forceinline__ unsigned long long ld_relaxed(const double *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tma::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// 16 bytes global -> shared through L2 only (cp.async.cg: coherent with what other SMs have stored, unlike .ca)
__device__ __forceinline__ void cp_async16_cg(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tma::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_inval(uint64_t *bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tma::smem_u32(bar)) : "memory");
}
// ---- thread-block cluster: consecutive planes in the CTAs of one cluster hand their values over through distributed
// shared memory (a remote store is visible after ~0.1 us; the way through L2 costs ~1.5 us and a polling consumer)
__device__ __forceinline__ unsigned int cluster_ctarank() {
    unsigned int r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned int cluster_nctarank() {
    unsigned int r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the shared::cluster address of `saddr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t cluster_map(uint32_t saddr, unsigned int rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u64(uint32_t caddr, unsigned long long v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(caddr), "l"(v) : "memory");
}
// The progress word of the hand-over: plain (weak) remote store / volatile local load on purpose.  A release / acquire
// pair at cluster scope also orders the warp's global stores (w, x: an L2 round trip) in front of the word and cost
// more than the hand-over saves (measured: 2.6 instead of 1.96 ms per HPCG-256 sweep).  What has to be ordered is only
// the consumer's own re-arming store to its LOCAL shared memory before the producer's next remote store to the same
// row; the producer issues that store after it has seen the word, i.e. two trips through the cluster network after the
// local store was issued.  Were it ever violated the consumer would wait for a value that has been overwritten by "not
// ready": the watchdog reports it, no wrong value can be read.
#ifndef WAVE_PROG_FENCED
#define WAVE_PROG_FENCED 0
#endif
__device__ __forceinline__ void st_cluster_prog(uint32_t caddr, int v) {
    if (WAVE_PROG_FENCED) asm volatile("st.release.cluster.shared::cluster.s32 [%0], %1;" ::"r"(caddr), "r"(v) : "memory");
    else asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(caddr), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_prog(uint32_t saddr) {
    int v;
    if (WAVE_PROG_FENCED) asm volatile("ld.acquire.cluster.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    else asm volatile("ld.volatile.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_volatile_shared_u64(uint32_t saddr) {
    unsigned long long v;
    asm volatile("ld.volatile.shared::cta.u64 %0, [%1];" : "=l"(v) : "r"(saddr) : "memory");
    return v;
}

// ---- exact division off the critical path ---------------------------------------------------------------------
// x[r] = (b[r] - sum) / D[r] must be the IEEE quotient (bit parity with the reference's `/`), but D[r] is known a
// step ahead: the reciprocal r = RN(1/D) is formed there with a full IEEE division, and the dependent chain only
// carries   q0 = RN(a r) ; e0 = fma(-d, q0, a) ; q1 = fma(e0, r, q0) ; e1 = fma(-d, q1, a) ; q = fma(e1, r, q1)
// (Markstein: a correctly rounded reciprocal and a faithful quotient make the last fma the correctly rounded
// quotient).  Exponents far from 0, subnormal numerators and a divisor whose significand is all ones take the
// IEEE division instead; tools/micro/divcheck.c checks the sequence against `/` on 3e10 operand pairs (random,
// next to rounding ties, long runs of ones) under exactly this guard: no mismatch.
__device__ __forceinline__ bool div_guard_d(double d) {
    const unsigned int hi = (unsigned int)__double2hiint(d), lo = (unsigned int)__double2loint(d);
    const unsigned int e = (hi >> 20) & 0x7ffu;
    return (e - 623u > 800u) || ((hi & 0xfffffu) == 0xfffffu && lo == 0xffffffffu);
}
__device__ __forceinline__ bool div_guard_a(double a) {   // also zero: the sequence would lose the sign of -0
    const unsigned int e = ((unsigned int)__double2hiint(a) >> 20) & 0x7ffu;
    return e - 623u > 800u;
}
// RN(1/d) for a divisor that passed div_guard_d: the hardware's 2^-23 estimate, three Newton steps, and the
// rounding step r = fma(x, fma(-d, x, 1), x) (checked against 1.0 / d on 3e9 divisors, tools/micro/divcheck.c).
// Branch-free, unlike the IEEE division: it can sit in the same basic block as the step's dependent chain.
__device__ __forceinline__ double rcp_exact(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
}
__device__ __forceinline__ double div_by_rcp(double a, double d, double r) {
    const double q0 = mul_rn(a, r);
    const double e0 = fma(-d, q0, a);
    const double q1 = fma(e0, r, q0);
    const double e1 = fma(-d, q1, a);
    return fma(e1, r, q1);
}

// ---- shared memory of a CTA (in doubles) ---------------------------------------------------------------------
// per 32-line block: the record stages, the b / D rings, the staging of requested previous-plane values, the
// stages' mbarriers; then the CTA's two rings.  A ring has RING logical rows (steps mod RING) and is stored TWICE
// (physical rows j and j + RING are copies): a reader whose base is the physical row of an even step reaches every
// logical row it needs (base - 3 .. base + 6) at a compile-time offset, without wrap-around arithmetic.
constexpr int RWC = 32 * MAX_WARPS + 32;     // ring width: column y + 1 of the plane's line y, one spare column either side
constexpr int RROWS = 2 * RING;
constexpr int SROWS = 8;                     // staging rows: a requested value sits there GA = 2 steps (rows 0..3); a plane whose
                                             // predecessor runs in the same cluster uses all eight as its INBOX (see below)
constexpr int IB = SROWS;                    // inbox depth in steps
constexpr int O_RINGB = NST * REC_DOUBLES;
constexpr int O_RINGD = O_RINGB + BD * 32;
constexpr int O_RINGS = O_RINGD + BD * 32;
constexpr int O_FULL = O_RINGS + SROWS * 32;
constexpr int BLK_DOUBLES = O_FULL + 8;
constexpr int O_GHOST = RROWS * RWC;         // ring of the previous plane's values, relative to the ring of results
static_assert(NST == 4 && BD == 8 && (GA == 2 || GA == 4) && PH == 2 && RING == 8 && SROWS == 8, "the slot arithmetic below is written for these");
__host__ __device__ inline size_t smem_bytes(int W) { return ((size_t)W * BLK_DOUBLES + 2 * (size_t)RROWS * RWC) * 8; }

extern __shared__ __align__(128) double wave_sm[];   // the CTA's dynamic shared memory (addressed directly: no generic pointers)

// everything a warp keeps across the steps of a plane.  The step loop is unrolled by PH = 2 only (one step in each
// role, see below): the whole loop body is ~4 KB of code per warp and stays in the instruction caches -- unrolled by
// the ring depth it was 90 KB, and the warps spent a quarter of their cycles waiting for instructions (ncu,
// profiles/r02_wave_*).  Addresses are those of the pair's even step; step U adds a compile-time offset.
struct State {
    int blk;                   // the block's shared memory (offset in doubles), + lane
    int ring;                  // ring of results, physical row 0, column of the line LEFT of this lane's line
    int lane;
    int nx_eff;                // nx, or 0 for a lane without a line (y >= ny): "0 <= xp < nx_eff" is "active"
    int ls;                    // local step of the block at U = 0 (even)
    int xp;                    // position of this lane in its line at U = 0
    // what prepare<U-1> left for this warp's solve<U>
    double pre;                // lower: sum of the products of slots 0..10, in order
    double pp[K - 2];          // upper: the products of the slots that come AFTER the two late ones in the sum
    double v_own, v_nb;        // matrix values of the two late slots: own predecessor (x-1) and (x+1, y-1)
    // what solve<U-2> left for solve<U>
    double bb, dd, rcp;        // b, D and RN(1/D) of the step's row (0, 1, 1 for a lane without a row)
    double bb_n, dd_n;         // b and D of this warp's step after that (loads in flight)
    bool ieee;                 // the step's divisor needs the IEEE division
    const double *pm;          // request at U = 0 (+ U * 32), WITHOUT the lane offset
    int pm_lo, pm_hi;          // ... valid while pm_lo <= ls < pm_hi
    double *w_out;             // w[grp][ls][lane] at U = 0 (+ U * 32)
    unsigned long long *wc_out;
    double *x_out;             // x[row of this lane] at U = 0 (+- U)
    const double *b_req, *d_req;   // b / D of the row this lane has BD steps later (+- U)
    uint64_t pol;
    const double *rec_next;    // record of local step ls + 1 + NST at U = 0 (+ U records)
    // hand-over inside a cluster (see wave_kernel)
    bool in_push;              // the previous plane runs in this cluster: its values arrive in this block's inbox
    bool out_push;             // the next plane runs in this cluster: results also go into its inbox
    uint32_t inbox;            // shared::cta address of this lane's inbox column, row 0
    uint32_t push_to;          // shared::cluster address of the same place in the next plane's CTA
    uint32_t prog;             // shared::cta address of this block's progress word (the next plane's CTA writes it)
    uint32_t prog_to;          // shared::cluster address of the previous plane's progress word for this block
#ifdef BIS_PERF_DEBUG
    long long polls;
#endif
};

// A CTA owns a whole plane: its 32-line blocks, PH = 2 warps each, block w running 64 steps behind block w-1.
// Every operand that is not a value of the previous plane comes out of the CTA's shared-memory rings (the lines
// left and right of a block belong to its neighbour blocks); ring rows are indexed by the CTA's GLOBAL step (the
// blocks' local steps differ by multiples of 64 = 0 mod RING).
//
// Of the 13 products of a row only TWO depend on the previous step (the own predecessor x-1 and (x+1, y-1), both
// results of step s-1).  So the two warps of a block take its steps in turn: while one SOLVES step s (solve<U>:
// the two late products, the end of the sum, the division, the stores), the other PREPARES step s+1 (prepare<U>:
// the record of s+1, its eleven early operands out of the rings, their products and, for a lower factor, their
// in-order sum; the arrival of the previous plane's values).  What a warp prepares it solves itself one step
// later: the hand-over is in registers, and the CTA's one barrier per step orders the rings.
// (A lower factor adds the two late products LAST, so its chain is two adds long.  An upper factor's storage
// order starts with them: every other product has to be added after them, thirteen dependent adds -- the
// summation order is the reference's and is not negotiable.)
// So that the eleven early operands of step s+1 are in the rings during step s, the values of the previous plane
// enter one step earlier than a plain schedule would need them (ring row s + 5 at the end of step s).
template <int U, bool UPPER>
__device__ __forceinline__ void solve(const Args &a, State &st, const int Sw) {
    const int ls = st.ls + U;
    const int xp = st.xp + U;
    const bool act = (unsigned)xp < (unsigned)st.nx_eff;
    const bool in_range = ls >= 0 && ls < Sw;                    // warp-uniform
    double *rp = wave_sm + st.ring + (st.ls & (RING - 1)) * RWC;      // physical row of the pair's even step
    // results of step s - 1 (the block's other warp stored them before the barrier): the line to the left, the own line
    constexpr int RM1 = (U == 0 ? RING - 1 : 0) * RWC;
    const double op_nb = rp[RM1 + 0];
    const double r_prev = rp[RM1 + 1];
    double sum;
    if (!UPPER) {
        sum = add_rn(st.pre, mul_rn(st.v_nb, op_nb));
        sum = add_rn(sum, mul_rn(st.v_own, r_prev));
    } else {
        sum = add_rn(0.0, mul_rn(st.v_own, r_prev));
        sum = add_rn(sum, mul_rn(st.v_nb, op_nb));
#pragma unroll
        for (int k = 0; k < K - 2; ++k) sum = add_rn(sum, st.pp[k]);
    }
    // a lane without a row divides 1 by 1 (a zero numerator takes the IEEE division) and publishes 0.0
    const double dd = st.dd;
    const double num = act ? sub_rn(st.bb, sum) : 1.0;
    double q = div_by_rcp(num, dd, st.rcp);
    if (st.ieee || div_guard_a(num)) q = div_rn(num, dd);    // rare: exponents far from 0, an exactly zero numerator, ...
    const double r = act ? q : 0.0;
    rp[U * RWC + 1] = r;
    rp[(U + RING) * RWC + 1] = r;
    // ---- results out ----------------------------------------------------------------------------------------
    if (in_range && st.out_push) {
        // inbox row ls % IB of the next plane's CTA is free once that CTA has taken step ls - IB out of it
        if (ld_prog(st.prog) < ls - IB) {
            unsigned int spins = 0;
            unsigned long long t_wd = 0;
            while (ld_prog(st.prog) < ls - IB) {
                if ((++spins & 1023u) == 0) {
                    if (t_wd == 0) t_wd = bis_globaltimer();
                    if (*reinterpret_cast<volatile int *>(a.errflag) != 0 || bis_globaltimer() - t_wd > WATCHDOG_NS) {
                        atomicExch(a.errflag, 6);
                        break;
                    }
                }
            }
        }
        st_cluster_u64(st.push_to + (uint32_t)((ls & (IB - 1)) * 256), (unsigned long long)__double_as_longlong(r));
    }
    // (a plane whose successor gets the values pushed needs no copy in the working vector; neither vector is touched
    // there, so both stay "not ready" at these places -- the host resets them when the cluster size changes)
    if (in_range && !st.out_push && !WAVE_DBG(a, 16)) {
        __stcg(st.w_out + U * 32, r);                                      // inactive lanes publish 0.0
        st.wc_out[U * 32] = SENT;
    }
    if (act && !WAVE_DBG(a, 4)) st.x_out[UPPER ? -U : U] = a.post_mul_d ? mul_rn(r, dd) : r;
    // ---- b, D and 1/D of this warp's next step (ls + PH), off the chain ----------------------------------
    // b / D were requested BD steps ahead: five younger copy groups of this thread may still be in flight
    // (prepare / solve / prepare / solve / prepare)
#if WAVE_BD_REG
    // b and D are loaded straight into registers two of this warp's steps (four steps) ahead.  They used to come
    // through cp.async rings, eight steps ahead -- but a thread's copy groups complete in order, so prepare's wait for
    // its two-step-old request of previous-plane values also waited for the b / D copies committed in between, and under
    // load those took longer than that (prepare 800 -> 1400 cycles in mid-sweep planes, profiles/r02b_wave_*).
    {
        const bool act2 = (unsigned)(xp + PH) < (unsigned)st.nx_eff;
        const double d2 = act2 ? st.dd_n : 1.0;
        st.bb = act2 ? st.bb_n : 0.0;
        st.dd = d2;
        st.ieee = div_guard_d(d2);
        st.rcp = rcp_exact(st.ieee ? 1.0 : d2);
        if ((unsigned)(xp + 2 * PH) < (unsigned)st.nx_eff) {
            st.bb_n = __ldcs(st.b_req + (UPPER ? -(U + 2 * PH - BD) : (U + 2 * PH - BD)));
            st.dd_n = __ldcs(st.d_req + (UPPER ? -(U + 2 * PH - BD) : (U + 2 * PH - BD)));
        }
    }
    if ((ls & 2) == 0 && (unsigned)(xp + PH + BD + 16) < (unsigned)st.nx_eff) {   // the sectors of later steps into L2
        asm volatile("prefetch.global.L2 [%0];" ::"l"(st.b_req + (UPPER ? -(U + PH + 16) : (U + PH + 16))));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(st.d_req + (UPPER ? -(U + PH + 16) : (U + PH + 16))));
    }
    cp_async_commit();
    return;
#endif
    if (!WAVE_DBG(a, 2)) cp_async_wait<BD - PH - 1>();
    double *bd = wave_sm + st.blk + O_RINGB + ((st.ls + PH) & (BD - 1)) * 32 + U * 32;    // slot (ls + PH) % BD: no wrap, st.ls is even
    {
        const bool act2 = (unsigned)(xp + PH) < (unsigned)st.nx_eff;
        st.bb = act2 ? bd[0] : 0.0;
        double d2 = act2 ? bd[O_RINGD - O_RINGB] : 1.0;
        if (WAVE_DBG(a, 512) && act2) {    // experiment: straight from global memory
            st.bb = __ldcg(st.b_req + (UPPER ? -(U + PH - BD) : (U + PH - BD)));
            d2 = __ldcg(st.d_req + (UPPER ? -(U + PH - BD) : (U + PH - BD)));
        }
        st.dd = d2;
        st.ieee = div_guard_d(d2);
        st.rcp = rcp_exact(st.ieee ? 1.0 : d2);
    }
    if ((unsigned)(xp + PH + BD) < (unsigned)st.nx_eff && !WAVE_DBG(a, 2)) {   // b and D of step ls + PH + BD, into the slots just read
        cp_async8(bd, st.b_req + (UPPER ? -(U + PH) : (U + PH)));
        cp_async8(bd + (O_RINGD - O_RINGB), st.d_req + (UPPER ? -(U + PH) : (U + PH)));
    }
    if ((ls & 2) == 0 && (unsigned)(xp + PH + BD + 16) < (unsigned)st.nx_eff) {   // ... and the sectors of sixteen steps later into L2
        asm volatile("prefetch.global.L2 [%0];" ::"l"(st.b_req + (UPPER ? -(U + PH + 16) : (U + PH + 16))));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(st.d_req + (UPPER ? -(U + PH + 16) : (U + PH + 16))));
    }
    cp_async_commit();
}

// shared-memory places of the record of local step ls + 1 (stage (ls + 1) % NST), for st.ls even
template <int U>
__device__ __forceinline__ int rec_stage(const State &st) {
    return U == 0 ? (st.ls & 2) + 1 : ((st.ls & 2) ^ 2);
}

// the block's other warp, during the same step: everything of step ls + 1 that does not need the result of step ls
template <int U, bool UPPER>
__device__ __forceinline__ void prepare(const Args &a, State &st, const int Sw) {
    const int ls = st.ls + U;
    const int stage = rec_stage<U>(st);
    // ---- waits (rarely taken loops), before the straight-line part ---------------------------------------
    if (ls + 1 >= 0 && ls + 1 < Sw && !WAVE_DBG(a, 1 | 32))
        tma::mbar_wait(reinterpret_cast<uint64_t *>(wave_sm + (st.blk - st.lane) + O_FULL) + stage, (uint32_t)(((ls + 1) / NST) & 1));
    // the value of the previous plane this warp requested GA steps ago: global -> shared by an asynchronous copy
    // through L2 only (no register is tied up while it is in flight, sixteen lanes fetch the row's 256 bytes);
    // GA - 1 younger copy groups of this thread may still be in flight
    const int s8 = st.ls & (RING - 1);
    const double *rp = wave_sm + st.ring + s8 * RWC;
    {
        const double *rv = wave_sm + st.blk + stage * REC_DOUBLES;
        double v[K];
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = rv[k * 32];
        double pre = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int kk = UPPER ? K - 1 - k : k;       // plane k of the record is stencil slot kk
            const int dz = kk / 9 - 1, dy = (kk / 3) % 3 - 1, dx = kk % 3 - 1;
            if (kk == K - 1) {
                st.v_own = v[k];
            } else if (kk == K - 2) {
                st.v_nb = v[k];
            } else {
                // produced at step (s + 1) + dx + 2 dy of the producing line: logical ring row U + 1 + dx + 2 dy
                // relative to the pair's even step, i.e. -2 .. 5: a copy of it sits at a fixed physical distance
                const int d = U + 1 + dx + 2 * dy;
                const double o = rp[(dz < 0 ? O_GHOST : 0) + (d < 0 ? d + RING : d) * RWC + 1 + dy];
                if (!UPPER) pre = add_rn(pre, mul_rn(v[k], o));
                else st.pp[k - 2] = mul_rn(v[k], o);
            }
        }
        st.pre = pre;
    }
    // ---- arrival of the previous plane's value for ring row ls + 5: looked at only now, after the products, so that
    // a value that is still on its way is waited for while this warp had something else to do
    unsigned long long vm = 0ull;
    if (st.in_push) {
        // the previous plane's CTA stores its step ls + 5 into row (ls + 5) % IB of this block's inbox (remote shared
        // memory; the value is its own ready flag): take it, mark the row "not ready" again and tell the producer
        const int q = ls + 5;
        if (q >= 0 && q < Sw) {                                  // warp-uniform
            const uint32_t slot = st.inbox + (uint32_t)((q & (IB - 1)) * 256);
            vm = ld_volatile_shared_u64(slot);
            if (__any_sync(0xffffffffu, vm == SENT)) {
#ifdef BIS_PERF_DEBUG
                st.polls += 1;
#endif
                unsigned int spins = 0;
                unsigned long long t_wd = 0;
                for (;;) {
                    if (vm == SENT) vm = ld_volatile_shared_u64(slot);
                    if (!__any_sync(0xffffffffu, vm == SENT)) break;
                    bool give_up = false;
                    if ((++spins & 1023u) == 0) {
                        if (t_wd == 0) t_wd = bis_globaltimer();
                        give_up = *reinterpret_cast<volatile int *>(a.errflag) != 0 || bis_globaltimer() - t_wd > WATCHDOG_NS;
                    }
                    if (__any_sync(0xffffffffu, give_up)) {
                        atomicExch(a.errflag, 6);
                        vm = 0ull;
                        break;
                    }
                }
            }
            asm volatile("st.shared::cta.u64 [%0], %1;" ::"r"(slot), "l"(SENT) : "memory");
            __syncwarp();
            if (st.lane == 0) st_cluster_prog(st.prog_to, q);
        }
    } else {
    cp_async_wait<GA - 1>();
    __syncwarp();
    const int o_stage = st.blk + O_RINGS + U * 32;              // + slot * 32
    vm = reinterpret_cast<const unsigned long long *>(wave_sm)[o_stage + (((st.ls - GA) & (SROWS - 2))) * 32];   // slot (ls - GA) % SROWS
    if (__any_sync(0xffffffffu, vm == SENT)) {
#ifdef BIS_PERF_DEBUG
        st.polls += 1;
#endif
        const double *pm = st.pm + (U - GA) * 32 + st.lane;   // where it was requested from
        unsigned int spins = 0;
        unsigned long long t_wd = 0;
        for (;;) {
            if (vm == SENT) vm = ld_relaxed(pm);
            if (!__any_sync(0xffffffffu, vm == SENT)) break;
            if (spins > 32) __nanosleep(200);   // a plane far ahead of the wavefront: stay off the L2
            bool give_up = false;
            if ((++spins & 1023u) == 0) {
                if (t_wd == 0) t_wd = bis_globaltimer();
                give_up = *reinterpret_cast<volatile int *>(a.errflag) != 0 || bis_globaltimer() - t_wd > WATCHDOG_NS;
            }
            if (__any_sync(0xffffffffu, give_up)) {
                atomicExch(a.errflag, 6);
                vm = 0ull;
                break;
            }
        }
        // Having had to poll means this plane sits right behind its predecessor, where every later request misses too
        // and every step pays this round trip (a stable state: the plane then runs at step + round trip, and so does
        // everything behind it).  Falling back puts the next requests into the zone where they hit.  Without clusters
        // that is no gain (every plane falls back: the same lag per plane, measured); with clusters only one plane in
        // CL is fed through L2, and the CL - 1 planes behind it follow at the pace of a step without a round trip.
        if (a.backoff_ns) __nanosleep(a.backoff_ns);
    }
    }
    // ---- the value of the previous plane requested GA steps ago enters the ring (logical row ls + 5) ------
    {
        int row = s8 + U + 5;
        if (row >= RING) row -= RING;
        double *gp = wave_sm + st.ring + O_GHOST + row * RWC + 1;
        const double gv = __longlong_as_double((long long)vm);
        gp[0] = gv;
        gp[RING * RWC] = gv;
    }
    // ---- request for a later step (consumed by this same warp: GA is a multiple of PH) ---------------------
    const int o_req = (st.blk - st.lane) + O_RINGS + ((st.ls & (SROWS - 2)) + U) * 32;     // slot ls % SROWS
    if (st.in_push) {
        // nothing to request: the values are pushed
    } else if (ls >= st.pm_lo && ls < st.pm_hi && !WAVE_DBG(a, 8)) {
        if (st.lane < 16) cp_async16_cg(wave_sm + o_req + 2 * st.lane, st.pm + U * 32 + 2 * st.lane);
    } else {
        wave_sm[o_req + st.lane] = 0.0;     // no such value: 0.0 (never "not ready")
    }
    cp_async_commit();
}

// after the CTA's barrier of the step: the record stage read in it (that of local step ls + 1) is free again
template <int U>
__device__ __forceinline__ void refill(const Args &a, State &st, const int Sw) {
    const int ls = st.ls + U;
    if (st.lane == 0 && !WAVE_DBG(a, 32)) {
        const int stage = rec_stage<U>(st);
        uint64_t *bar = reinterpret_cast<uint64_t *>(wave_sm + st.blk + O_FULL) + stage;
        if (ls + 1 + NST >= 0 && ls + 1 + NST < Sw) {
            // the stage was read through the generic proxy (ld.shared, ordered before this point by the CTA's
            // barrier); the bulk copy writes it through the async proxy.  Without this fence the copy of record
            // ls + 1 + NST has been seen to overtake the last loads of record ls + 1 (a wrong row every ~100 solves
            // at 8 blocks per plane; visible only where consecutive records differ, i.e. at line ends)
            tma::fence_reads_before_bulk_write();
            tma::mbar_expect_tx(bar, (uint32_t)(REC_DOUBLES * 8));
            tma::bulk_g2s(wave_sm + st.blk + stage * REC_DOUBLES, st.rec_next + (size_t)U * REC_DOUBLES, (uint32_t)(REC_DOUBLES * 8), bar, st.pol);
        }
        // the record stream comes from HBM with nothing but this block asking for it: it is pulled into L2 well ahead
        if (PF > 0 && ls + 1 + NST + PF < Sw)
            prefetch_l2_bulk(st.rec_next + (size_t)(U + PF) * REC_DOUBLES, (uint32_t)(REC_DOUBLES * 8));
    }
}

// the bases move by one pair of steps
template <bool UPPER>
__device__ __forceinline__ void advance_pair(State &st) {
    st.ls += PH;
    st.xp += PH;
    st.pm += PH * 32;
    st.w_out += PH * 32;
    st.wc_out += PH * 32;
    st.x_out += UPPER ? -PH : PH;
    st.b_req += UPPER ? -PH : PH;
    st.d_req += UPPER ? -PH : PH;
    st.rec_next += (size_t)PH * REC_DOUBLES;
}

template <bool UPPER>
__global__ void __launch_bounds__(MAX_WARPS * PH * 32 + 32, 1) wave_kernel(Args a) {
    __shared__ long long s_plane;
    __shared__ int s_prog[MAX_WARPS];
    // CL consecutive planes are taken by the CL CTAs of a cluster (rank r: plane CL * ticket + r).  Ranks > 0 get
    // the values of their previous plane PUSHED into shared memory by the CTA of rank r - 1 (st.shared::cluster,
    // visible after ~0.1 us, polled in local shared memory); rank 0 fetches them from the working vector in L2 as
    // before (the plane before it belongs to an earlier ticket, i.e. to a cluster that already runs).  Launched
    // without clusters, CL = 1 and everything goes through L2.
    const unsigned int CL = cluster_nctarank(), crank = cluster_ctarank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // The CTA's last warp computes nothing: its lane w re-fills the record stages of block w (see refill).  With the
    // bulk copies issued by lane 0 of the block's solving warp -- a proxy fence, an mbarrier transaction, the copy and an
    // L2 prefetch, ~500 cycles of one lane -- every second solve started that much later (profiles/r02b_wave_*).
    const bool dma = warp >= a.g.W * PH;
    const int wl = dma ? 0 : warp / PH;                  // 32-line block of the plane
    const int ph = warp % PH;                            // this warp solves the block's steps with ls % PH == ph
    const Grid g = a.g;
    const int W = g.W;
    const int Sw = g.S;                                  // local steps of a block: nx + 62
    State st;
    st.lane = lane;
    st.blk = wl * BLK_DOUBLES + lane;
    st.ring = W * BLK_DOUBLES + 32 * wl + lane;          // column y + 1 is at [+1]; column y (the line to the left) at [+0]
    st.pol = tma::policy_evict_first();
    // records start finite: steps before the first record multiply whatever the stage holds
    for (int i = threadIdx.x; i < W * BLK_DOUBLES; i += blockDim.x) wave_sm[i] = 0.0;
    const int y = wl * 32 + lane;
    uint64_t *full = reinterpret_cast<uint64_t *>(wave_sm + wl * BLK_DOUBLES + O_FULL);
    for (;;) {
        // every CTA of the cluster has left its plane: nothing is pushed into this CTA's inboxes any more
        if (CL > 1) cluster_sync();
        else __syncthreads();
        if (threadIdx.x == 0 && crank == 0) {
            const long long t = (long long)atomicAdd(a.ticket, 1u);
            if (CL > 1) {
                for (unsigned int r = 0; r < CL; ++r)
                    st_cluster_u64(cluster_map(tma::smem_u32(&s_plane), r), (unsigned long long)(t * CL + r));
            } else {
                s_plane = t;
            }
        }
        for (int i = threadIdx.x; i < 2 * RROWS * RWC; i += blockDim.x) wave_sm[W * BLK_DOUBLES + i] = 0.0;   // both rings
        {
            const double arm = crank > 0 ? __longlong_as_double((long long)SENT) : 0.0;     // inbox rows start "not ready"
            if (!dma)
                for (int i = lane + 32 * ph; i < SROWS * 32; i += 32 * PH) wave_sm[wl * BLK_DOUBLES + O_RINGS + i] = arm;
        }
        if (threadIdx.x < MAX_WARPS) s_prog[threadIdx.x] = -1;
        if (lane == 0 && ph == 0 && !dma) {
            for (int i = 0; i < NST; ++i) tma::mbar_init(&full[i], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (CL > 1) cluster_sync();
        else __syncthreads();
        const long long z = s_plane;
        if (z - crank >= g.nz) break;          // the whole cluster is past the last plane
        if (z >= g.nz) {                       // only this rank is: keep the cluster's barriers
            if (lane == 0 && ph == 0 && !dma)
                for (int i = 0; i < NST; ++i) mbar_inval(&full[i]);
            continue;
        }
        const int ls0 = -2 * RING;               // first local step of a block (request / ring-fill machinery only)
        const int s_end = 64 * (W - 1) + Sw;
        if (dma) {
            // lane w < W: the record stream of block w; same steps and the same barriers as the computing warps
            st.lane = 0;
            st.blk = lane * BLK_DOUBLES;
            st.ls = ls0;
            st.rec_next = a.rec + (((z * W + lane) * Sw) + (ls0 + 1 + NST)) * REC_DOUBLES;
#pragma unroll 1
            for (int s = ls0; s < s_end; s += PH) {
                const int l = s - 64 * lane;
                const bool live = lane < W && l >= ls0 && l < Sw;
                __syncthreads();
                if (live) refill<0>(a, st, Sw);
                __syncthreads();
                if (live) {
                    refill<1>(a, st, Sw);
                    st.ls += PH;
                    st.rec_next += (size_t)PH * REC_DOUBLES;
                }
            }
            __syncthreads();
            continue;
        }
        const long long grp = z * W + wl;
        const long long line_p0 = (z * g.ny + y) * g.nx;
        const bool has_prev = z > 0;
        st.nx_eff = y < g.ny ? g.nx : 0;
        st.ls = ls0;
        st.xp = ls0 - 2 * lane;
        st.pre = 0.0;
        st.v_own = st.v_nb = 0.0;
        st.bb = st.bb_n = 0.0;
        st.dd = st.rcp = st.dd_n = 1.0;
        st.ieee = false;
#pragma unroll
        for (int i = 0; i < K - 2; ++i) st.pp[i] = 0.0;
        // the own line in the previous plane; requested at local step ls for ring row ls + 5 + GA
        st.pm = a.w + ((has_prev ? grp - W : 0) * Sw + (ls0 + 5 + GA)) * 32;
        st.pm_lo = -(5 + GA);
        st.pm_hi = has_prev ? Sw - (5 + GA) : -(1 << 30);
        st.in_push = crank > 0;
        st.out_push = crank + 1 < CL && z + 1 < g.nz;
        st.inbox = tma::smem_u32(wave_sm + wl * BLK_DOUBLES + O_RINGS + lane);
        st.prog = tma::smem_u32(&s_prog[wl]);
        st.push_to = st.out_push ? cluster_map(st.inbox, crank + 1) : 0u;
        st.prog_to = st.in_push ? cluster_map(st.prog, crank - 1) : 0u;
        st.w_out = a.w + (grp * Sw + ls0) * 32 + lane;
        st.wc_out = reinterpret_cast<unsigned long long *>(a.w_clean) + (grp * Sw + ls0) * 32 + lane;
        {
            // row of this lane at position xp: p = line_p0 + xp (lower), n - 1 - p (upper)
            const long long p_now = line_p0 + st.xp;
            const long long row_now = UPPER ? g.n - 1 - p_now : p_now;
            st.x_out = a.x + row_now;
            st.b_req = a.b + (UPPER ? row_now - BD : row_now + BD);
            st.d_req = a.D + (UPPER ? row_now - BD : row_now + BD);
        }
        st.rec_next = a.rec + (grp * Sw + (ls0 + 1 + NST)) * REC_DOUBLES;
        // global steps of the CTA: block w's local step is s - 64 w (its lane 0 is "lane 32 w" of the plane);
        // all warps walk the same steps and keep the same barriers
#ifdef BIS_PERF_DEBUG
        if (a.stamps && threadIdx.x == 0) a.stamps[2 * z] = bis_globaltimer();
#endif
#ifdef BIS_PERF_DEBUG
        long long dbg_acc[5] = {0, 0, 0, 0, 0};
        st.polls = 0;
#endif
#pragma unroll 1
        for (int s = ls0; s < s_end; s += PH) {
            const int l = s - 64 * wl;
            const bool live = l >= ls0 && l < Sw;    // warp-uniform; st.ls == l while live
#ifdef BIS_PERF_DEBUG
            const long long tc0 = clock64();
#endif
            if (live) {
                if (ph == 0) solve<0, UPPER>(a, st, Sw);
                else prepare<0, UPPER>(a, st, Sw);
            }
#ifdef BIS_PERF_DEBUG
            const long long tc1 = clock64();
#endif
            __syncthreads();
#ifdef BIS_PERF_DEBUG
            const long long tc2 = clock64();
#endif
            if (live) {
                if (ph == 0) {
                    prepare<1, UPPER>(a, st, Sw);
                } else {
                    solve<1, UPPER>(a, st, Sw);
                }
            }
#ifdef BIS_PERF_DEBUG
            const long long tc3 = clock64();
#endif
            __syncthreads();
#ifdef BIS_PERF_DEBUG
            if (live && l >= 64 && l + 64 < Sw) {
                const long long tc4 = clock64();
                // ph 0: phase A = solve, phase B = prepare ; ph 1: phase A = prepare, phase B = refill + solve
                dbg_acc[0] += tc1 - tc0;
                dbg_acc[1] += tc2 - tc1;
                dbg_acc[2] += tc3 - tc2;
                dbg_acc[3] += tc4 - tc3;
                dbg_acc[4] += 1;
            }
#endif
            if (live) advance_pair<UPPER>(st);
        }
        cp_async_wait<0>();
        __syncthreads();
        if (lane == 0 && ph == 0)
            for (int i = 0; i < NST; ++i) mbar_inval(&full[i]);
#ifdef BIS_PERF_DEBUG
        if (a.stamps && threadIdx.x == 0) a.stamps[2 * z + 1] = bis_globaltimer();
        if (a.stamps && lane == 0 && wl == (W > 1 ? 1 : 0)) {
            int slot = -1;
            const long long zs[5] = {1, 8, 33, g.nz / 2, g.nz / 2 + 4};
            for (int i = 0; i < 5; ++i)
                if (z == zs[i]) slot = i;
            if (slot >= 0) {
                for (int i = 0; i < 5; ++i) a.stamps[2 * g.nz + (slot * 2 + ph) * 6 + i] = (unsigned long long)dbg_acc[i];
                a.stamps[2 * g.nz + (slot * 2 + ph) * 6 + 5] = (unsigned long long)st.polls;
            }
        }
#endif
    }
}

}  // namespace wave
