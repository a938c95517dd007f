/* oracle/port/bis_oracle.c -- TEST INFRASTRUCTURE, never shipped, never timed
 * as the product.
 *
 * Plain-C, single-threaded restatement of the iteration-loop hot path of
 * DanecLacey/basic_iterative_solvers (citations are file:line under
 * /root/reference).  It is the checker the CUDA path is compared with in
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * Pinning: tests/test_oracle_*.py check every function below against
 *   (1) the reference's own known-answer tests (tests/test_kernels.cpp,
 *       tests/test_utilities.cpp, tests/test_solvers.cpp), and
 *   (2) the compiled, unmodified reference (oracle/_ref/libbis_ref.so) in
 *       the build container, whose outputs are committed as fixtures under
 *       tests/golden/ by tests/golden/make_golden.py.
 *
 * Floating-point contract (re-checked by objdump on the
 * -march=x86-64-v3 build of the reference used here): compiled with
 * -ffp-contract=off so that every rounding below is explicit:
 *   - triangular solves: unfused multiply then add, storage order, then one
 *     subtract and one divide (reference codegen: vmulsd/vaddsd/vsubsd/vdivsd)
 *   - sum_vectors / subtract_vectors: one fused multiply-add (vfmadd/vfnmadd)
 *   - elemwise_mult / elemwise_div / scale / normalize_x: separately rounded
 *   - ILU(0) row update: fused (vfnmadd231sd)
 *   - SpMV: unfused multiply then add in storage order (GCC's in-order simd
 *     reduction) -- bit-identical to the compiled reference at any thread count
 *     (threads only split rows)
 *   - dot, norm: left-to-right sum of separately rounded products; identical to
 *     the reference at ONE OpenMP thread, tolerance only at more threads (the
 *     cross-thread combine order is OpenMP's).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define O_MAX_ITERS 1000          /* CMakeLists.txt:20 */
#define O_TOL 1e-14               /* CMakeLists.txt:21 */
#define O_RES_CHECK_LEN 1         /* CMakeLists.txt:23 */
#define O_INIT_X_VAL 0.1          /* CMakeLists.txt:26 */
#define O_B_VAL 1.0               /* CMakeLists.txt:27 */
#define O_ILU0_PIVOT_TOL 1e-8     /* CMakeLists.txt:28 */
#define O_ILU0_PIVOT_REPL 1e-4    /* CMakeLists.txt:29 */

/* Solver::tolerance / Solver::max_iters are public members of the reference
 * (solver.hpp:29-30) initialised from the macros above; tests may lower them
 * (never above O_MAX_ITERS: the history arrays are sized by it). */
static double o_tol = O_TOL;
static int o_max_iters = O_MAX_ITERS;
void o_set_params(double tol, int max_iters) {
    o_tol = tol > 0.0 ? tol : O_TOL;
    o_max_iters = (max_iters > 0 && max_iters <= O_MAX_ITERS) ? max_iters : O_MAX_ITERS;
}

/* common.hpp:38-56 */
enum { P_NONE = 0, P_J = 1, P_GS = 2, P_BGS = 3, P_SGS = 4, P_2ST = 5, P_S2ST = 6, P_ILU0 = 7 };
enum { M_J = 0, M_GS = 1, M_SGS = 2, M_GM = 3, M_CG = 4, M_BI = 5 };

typedef struct {
    int n;
    int nnz;
    const int *rp;
    const int *col;
    const double *val;
} crs_t;

/* ---- kernels.hpp --------------------------------------------------------- */

/* kernels.hpp:22-42 native_spmv: y[r] = sum_k val[k] * x[col[k]].
 * Order and rounding of the compiled reference (g++ 13.3, -O3 -fopenmp,
 * x86-64-v3; established by single-row experiments against oracle/_ref, see
 * DESIGN.md "summation orders"): GCC may not reassociate the `omp simd
 * reduction(+)` without -ffast-math, so it vectorises the multiplies and adds
 * the separately rounded products IN STORAGE ORDER (fold-left reduction):
 * acc = acc + (val*x), two roundings per nonzero, no FMA. */
void o_spmv(int n_rows, const int *rp, const int *col, const double *val,
            const double *x, double *y) {
    for (int r = 0; r < n_rows; ++r) {
        double acc = 0.0;
        for (int k = rp[r]; k < rp[r + 1]; ++k) {
            double p = val[k] * x[col[k]];
            acc = acc + p;
        }
        y[r] = acc;
    }
}

/* kernels.hpp:54-76 native_sptrsv: rows ascending, x may alias b */
void o_sptrsv(int n, const int *rp, const int *col, const double *val,
              double *x, const double *D, const double *b) {
    for (int r = 0; r < n; ++r) {
        double s = 0.0;
        for (int k = rp[r]; k < rp[r + 1]; ++k) {
            double p = val[k] * x[col[k]];
            s = s + p;
        }
        x[r] = (b[r] - s) / D[r];
    }
}

/* kernels.hpp:88-107 native_bsptrsv: rows descending */
void o_bsptrsv(int n, const int *rp, const int *col, const double *val,
               double *x, const double *D, const double *b) {
    for (int r = n - 1; r >= 0; --r) {
        double s = 0.0;
        for (int k = rp[r]; k < rp[r + 1]; ++k) {
            double p = val[k] * x[col[k]];
            s = s + p;
        }
        x[r] = (b[r] - s) / D[r];
    }
}

/* kernels.hpp:119-126 */
void o_subtract_vectors(double *out, const double *a, const double *b, int n, double s) {
    for (int i = 0; i < n; ++i) out[i] = fma(-s, b[i], a[i]);
}
/* kernels.hpp:128-135 */
void o_sum_vectors(double *out, const double *a, const double *b, int n, double s) {
    for (int i = 0; i < n; ++i) out[i] = fma(s, b[i], a[i]);
}
/* kernels.hpp:137-144: (a*s)*b */
void o_elemwise_mult_vectors(double *out, const double *a, const double *b, int n, double s) {
    for (int i = 0; i < n; ++i) { double t = a[i] * s; out[i] = t * b[i]; }
}
/* kernels.hpp:146-153: a/(s*b) */
void o_elemwise_div_vectors(double *out, const double *a, const double *b, int n, double s) {
    for (int i = 0; i < n; ++i) { double t = s * b[i]; out[i] = a[i] / t; }
}
/* kernels.hpp:194-203 */
double o_euclidean_vec_norm(const double *v, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) { double p = v[i] * v[i]; acc = acc + p; }
    return sqrt(acc);
}
/* kernels.hpp:205-212 */
double o_dot(const double *a, const double *b, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) { double p = a[i] * b[i]; acc = acc + p; }
    return acc;
}
/* kernels.hpp:214-220 */
void o_scale(double *out, const double *v, double s, int n) {
    for (int i = 0; i < n; ++i) out[i] = v[i] * s;
}
/* kernels.hpp:236-241 */
void o_init_vector(double *v, double val, long n) {
    for (long i = 0; i < n; ++i) v[i] = val;
}
/* kernels.hpp:252-257 */
void o_copy_vector(double *out, const double *in, int n) {
    for (int i = 0; i < n; ++i) out[i] = in[i];
}
/* kernels.hpp:155-162 */
void o_compute_residual(const crs_t *A, const double *x, const double *b,
                        double *r, double *tmp) {
    o_spmv(A->n, A->rp, A->col, A->val, x, tmp);
    o_subtract_vectors(r, b, tmp, A->n, 1.0);
}

/* dense helpers, row-major: kernels.hpp:222-234, 243-250, 259-310 */
static void o_identity(double *m, int nr, int nc) {
    for (int i = 0; i < nr; ++i)
        for (int j = 0; j < nc; ++j) m[nc * i + j] = (i == j) ? 1.0 : 0.0;
}
static void o_copy_dense(double *dst, const double *src, int nr, int nc) {
    memcpy(dst, src, sizeof(double) * (size_t)nr * nc);
}
/* C = A(nra x nca) * B(nca x ncb); kernels.hpp:273-284 (the reference build
 * inlines this into least_squares and contracts `tmp += a*b` to vfmadd231sd) */
static void o_dgemm_t2(const double *A, const double *B, double *C, int nra, int nca, int ncb) {
    for (int i = 0; i < nra; ++i)
        for (int j = 0; j < ncb; ++j) {
            double t = 0.0;
            for (int k = 0; k < nca; ++k) t = fma(A[i * nca + k], B[k * ncb + j], t);
            C[i * ncb + j] = t;
        }
}
/* kernels.hpp:299-310: y[i] = sum_j (alpha*A_ij)*x_j, alpha = 1 */
static void o_dgemv(const double *A, const double *x, double *y, int nr, int nc) {
    for (int i = 0; i < nr; ++i) {
        y[i] = 0.0;
        for (int j = 0; j < nc; ++j) { double p = (1.0 * A[i * nc + j]) * x[j]; y[i] = y[i] + p; }
    }
}

/* kernels.hpp:312-333.  PRECOND_INNER_ITERS is a compile-time -D of the reference (0 in its default build,
 * CMakeLists.txt:25); here a run-time setting so that one library serves the fixtures of the 0 / 1 / 2
 * flavours (oracle/_ref/libbis_ref_in1.so, _in2.so).  work / tmp swap LOCALLY (std::swap on by-value
 * pointers), the caller's buffers keep their roles. */
static int g_precond_inner_iters = 0;
void o_set_precond_inner_iters(int k) { g_precond_inner_iters = k; }
static void o_two_stage_gs(const crs_t *strict, double *tmp, double *work, const double *D_inv,
                           const double *in, double *out, int n) {
    o_elemwise_mult_vectors(work, D_inv, in, n, 1.0);
    o_copy_vector(out, work, n);
    for (int inner = 1; inner <= g_precond_inner_iters; ++inner) {
        o_spmv(strict->n, strict->rp, strict->col, strict->val, work, tmp);
        o_elemwise_mult_vectors(tmp, D_inv, tmp, n, -1.0);
        double *t = work;
        work = tmp;
        tmp = t;
        o_sum_vectors(out, out, work, n, 1.0);
    }
}

/* kernels.hpp:336-414 (PRECOND_OUTER_ITERS = 1) */
void o_apply_preconditioner(int p, int n, const crs_t *Ls, const crs_t *Us,
                            const double *A_D, const double *A_D_inv,
                            const double *L_D, const double *U_D, double *out,
                            double *in, double *tmp, double *work) {
    switch (p) {
    case P_J:
        o_elemwise_div_vectors(out, in, A_D, n, 1.0);
        break;
    case P_GS:
        o_sptrsv(n, Ls->rp, Ls->col, Ls->val, out, A_D, in);
        break;
    case P_BGS:
        o_bsptrsv(n, Us->rp, Us->col, Us->val, out, A_D, in);
        break;
    case P_SGS:
        o_sptrsv(n, Ls->rp, Ls->col, Ls->val, tmp, A_D, in);
        o_elemwise_mult_vectors(tmp, tmp, A_D, n, 1.0);
        o_bsptrsv(n, Us->rp, Us->col, Us->val, out, A_D, tmp);
        break;
    case P_2ST:
        o_two_stage_gs(Ls, tmp, work, A_D_inv, in, out, n);
        break;
    case P_S2ST:
        o_two_stage_gs(Ls, tmp, work, A_D_inv, in, out, n);
        o_elemwise_mult_vectors(out, out, A_D, n, 1.0);
        o_two_stage_gs(Us, tmp, work, A_D_inv, out, out, n);
        break;
    case P_ILU0:
        o_sptrsv(n, Ls->rp, Ls->col, Ls->val, tmp, L_D, in);
        o_bsptrsv(n, Us->rp, Us->col, Us->val, out, U_D, tmp);
        break;
    default:
        o_copy_vector(out, in, n);
    }
}

/* ---- utilities/LU_factors.hpp ------------------------------------------- */

/* split_LU_new (LU_factors.hpp:122-309), strict parts only (the hot path never
 * touches L/U with diagonal), plus the diagonal that
 * peel_diag_crs_new (LU_factors.hpp:827-869) extracts from L and U.
 * Two calls: counts first (l_col == NULL), then fill. */
void o_split_strict(int n, const int *rp, const int *col, const double *val,
                    int *l_rp, int *l_col, double *l_val, int *u_rp, int *u_col,
                    double *u_val, double *A_D, double *A_D_inv) {
    int nl = 0, nu = 0;
    l_rp[0] = 0;
    u_rp[0] = 0;
    for (int i = 0; i < n; ++i) {
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            int c = col[k];
            if (c < i) {
                if (l_col) { l_col[nl] = c; l_val[nl] = val[k]; }
                ++nl;
            } else if (c > i) {
                if (u_col) { u_col[nu] = c; u_val[nu] = val[k]; }
                ++nu;
            } else if (A_D) {
                A_D[i] = val[k];
                if (A_D_inv) A_D_inv[i] = 1.0 / val[k];
            }
        }
        l_rp[i + 1] = nl;
        u_rp[i + 1] = nu;
    }
}

static int cmp_int(const void *a, const void *b) {
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* factor_ILU0_old (LU_factors.hpp:320-539): row-wise IKJ on A's pattern.
 * Outputs strict L (unit diagonal implied, L_D = 1), strict U, U_D.
 * l_rp/u_rp sized n+1; l_col/l_val sized nnz_lower(A); u_* sized nnz_upper(A):
 * ILU(0) keeps A's pattern.  Columns come out ascending (std::sort at :349,
 * :423; split_LU at :538 preserves order). */
void o_ilu0(int n, const int *rp, const int *col, const double *val, int *l_rp,
            int *l_col, double *l_val, double *L_D, int *u_rp, int *u_col,
            double *u_val, double *U_D) {
    double *w = (double *)calloc((size_t)n, sizeof(double));
    int maxlen = 0;
    for (int i = 0; i < n; ++i)
        if (rp[i + 1] - rp[i] > maxlen) maxlen = rp[i + 1] - rp[i];
    int *idx = (int *)malloc(sizeof(int) * (size_t)(maxlen + 1));
    int nl = 0, nu = 0;
    l_rp[0] = 0;
    u_rp[0] = 0;
    for (int i = 0; i < n; ++i) {
        int len = 0;
        for (int k = rp[i]; k < rp[i + 1]; ++k) {   /* :341-346 scatter */
            w[col[k]] = val[k];
            idx[len++] = col[k];
        }
        qsort(idx, (size_t)len, sizeof(int), cmp_int);   /* :349 */
        for (int q = 0; q < len; ++q) {                   /* :355-389 eliminate */
            int k = idx[q];
            if (k >= i) break;
            double pivot = U_D[k];                        /* U(k,k), :360-366 */
            if (fabs(pivot) < 1e-16) continue;            /* :370 */
            double f = w[k] / pivot;
            w[k] = f;
            for (int t = u_rp[k]; t < u_rp[k + 1]; ++t) { /* j > k entries of U row k */
                int j = u_col[t];
                if (w[j] != 0.0) w[j] = fma(-f, u_val[t], w[j]);   /* :384-386 */
            }
        }
        double ud = 0.0;
        for (int q = 0; q < len; ++q) {                   /* :398-406 gather */
            int j = idx[q];
            if (j < i) { l_col[nl] = j; l_val[nl] = w[j]; ++nl; }
            else if (j == i) ud = w[j];
            else { u_col[nu] = j; u_val[nu] = w[j]; ++nu; }
        }
        if (fabs(ud) < O_ILU0_PIVOT_TOL)                  /* :410-412 */
            ud = (ud >= 0 ? 1.0 : -1.0) * O_ILU0_PIVOT_REPL;
        U_D[i] = ud;
        L_D[i] = 1.0;                                     /* :508-509 */
        l_rp[i + 1] = nl;
        u_rp[i + 1] = nu;
        for (int q = 0; q < len; ++q) w[idx[q]] = 0.0;    /* :431-434 */
    }
    free(w);
    free(idx);
}

/* ---- methods/jacobi.hpp:27-40 normalize_x -------------------------------- */
void o_normalize_x(double *x_new, const double *x_old, const double *D,
                   const double *b, int n) {
    for (int i = 0; i < n; ++i) {
        double scaled = D[i] * x_old[i];
        double adj = x_new[i] - scaled;
        x_new[i] = (b[i] - adj) / D[i];
    }
}

/* ---- methods/gmres.hpp:55-121 least_squares ------------------------------ */
void o_gmres_least_squares(int k, int m, double *J, const double *H, double *H_tmp,
                           double *Q, double *Q_tmp, double *R) {
    o_identity(J, m + 1, m + 1);
    o_identity(H_tmp, m + 1, m);
    if (k == 0) o_copy_dense(H_tmp, H, m + 1, m);
    else o_dgemm_t2(Q, H, H_tmp, m + 1, m + 1, m);
    double a = H_tmp[k * m + k], b = H_tmp[(k + 1) * m + k];
    double den = sqrt(fma(a, a, b * b));   /* vmulsd + vfmadd231sd + vsqrtsd */
    double c = a / den, s = b / den;
    J[k * (m + 1) + k] = c;
    J[k * (m + 1) + k + 1] = s;
    J[(k + 1) * (m + 1) + k] = -1.0 * s;
    J[(k + 1) * (m + 1) + k + 1] = c;
    o_dgemm_t2(J, Q, Q_tmp, m + 1, m + 1, m + 1);
    o_copy_dense(Q, Q_tmp, m + 1, m + 1);
    o_dgemm_t2(Q, H, R, m + 1, m + 1, m);
}
/* methods/gmres.hpp:123-148 update_g */
double o_gmres_update_g(int k, int m, const double *Q, double *g, double *g_tmp, double beta) {
    o_init_vector(g_tmp, 0.0, m + 1);
    g_tmp[0] = beta;
    o_copy_vector(g, g_tmp, m + 1);
    o_dgemv(Q, g, g_tmp, m + 1, m + 1);
    o_copy_vector(g, g_tmp, m + 1);
    return fabs(g[k + 1]);
}

/* ---- the solver: solver.hpp + solver_harness.hpp + methods/{cg,...}.hpp --- */
typedef struct {
    int method, precond, m;           /* m = gmres restart length */
    int n;
    crs_t A, Ls, Us;
    double *A_D, *A_D_inv, *L_D, *U_D;
    double *x_star, *x_0, *b, *tmp, *work, *residual, *residual_0;
    /* per-method vectors */
    double *x_new, *x_old, *p_new, *p_old, *z_new, *z_old, *r_new, *r_old;
    double *v, *h, *s, *s_tmp, *y, *z;          /* bicgstab */
    double *x, *V, *Vy, *yk, *H, *H_tmp, *J, *Q, *Q_tmp, *w, *R, *g, *g_tmp;
    double rho_old, rho_new, beta;
    /* bookkeeping (solver.hpp:25-35) */
    double stopping, residual_norm;
    int iter_count, hist_count, restart_count, restarted, converged;
    double *hist;
} osolver;

static double *vnew(int n, double v) {
    double *p = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) p[i] = v;
    return p;
}

static void o_precond(osolver *S, double *out, double *in) {
    o_apply_preconditioner(S->precond, S->n, &S->Ls, &S->Us, S->A_D, S->A_D_inv,
                           S->L_D, S->U_D, out, in, S->tmp, S->work);
}

/* GMRESSolver::init_structs (gmres.hpp:241-272) */
static void gm_init_structs(osolver *S) {
    int n = S->n, m = S->m;
    o_init_vector(S->tmp, 0.0, n);          /* Solver::init_structs solver.hpp:112-120 */
    o_init_vector(S->work, 0.0, n);
    o_init_vector(S->residual, 0.0, n);
    o_init_vector(S->residual_0, 0.0, n);
    if (!S->restarted) {
        o_copy_vector(S->x, S->x_0, n);
        o_copy_vector(S->x_old, S->x_0, n);
    }
    o_init_vector(S->V, 0.0, (long)n * (m + 1));
    o_init_vector(S->Vy, 0.0, n);
    o_init_vector(S->w, 0.0, n);
    o_init_vector(S->yk, 0.0, m);
    o_init_vector(S->g, 0.0, m + 1);
    o_init_vector(S->g_tmp, 0.0, m + 1);
    o_init_vector(S->H, 0.0, (m + 1) * m);
    o_init_vector(S->H_tmp, 0.0, (m + 1) * m);
    o_identity(S->J, m + 1, m + 1);
    o_init_vector(S->R, 0.0, (m + 1) * m);
    o_identity(S->Q, m + 1, m + 1);
    o_identity(S->Q_tmp, m + 1, m + 1);
}

/* GMRESSolver::init_residual (gmres.hpp:274-324) */
static void gm_init_residual(osolver *S) {
    int n = S->n;
    o_compute_residual(&S->A, S->x, S->b, S->residual, S->tmp);
    if (!S->restarted) {
        S->residual_norm = o_euclidean_vec_norm(S->residual, n);
        S->hist[S->hist_count++] = S->residual_norm;
    }
    o_precond(S, S->residual, S->residual);
    double pn = o_euclidean_vec_norm(S->residual, n);
    S->beta = pn;
    S->g[0] = pn;
    S->g_tmp[0] = pn;
    o_scale(S->V, S->residual, 1.0 / pn, n);
    if (S->restarted) {
        S->residual_norm = pn;
        o_copy_vector(S->residual_0, S->residual, n);   /* Solver::init_residual solver.hpp:148-151 */
        S->hist[S->hist_count++] = S->residual_norm;
    }
}

/* GMRESSolver::get_explicit_x (gmres.hpp:326-375).  SURVEY.md F6: the
 * reference reads y[k] one past the end when k == m; de-facto that word is 0,
 * i.e. only k terms contribute.  Restated as the sum over j < k. */
static void gm_get_explicit_x(osolver *S) {
    int n = S->n, m = S->m;
    int k = S->iter_count - S->restart_count * m;
    double diag = 1.0;
    for (int r = k - 1; r >= 0; --r) {
        double sum = 0.0;
        for (int c = r; c < k; ++c) {
            if (r == c) diag = S->R[r * m + c];
            else sum = fma(S->R[r * m + c], S->yk[c], sum);   /* vfmadd231sd */
        }
        S->yk[r] = (S->g[r] - sum) / diag;
    }
    /* dgemm_transpose1 (kernels.hpp:259-271) with n_cols_A = k+1; term k is
     * V[k]*y[k] with y[k] == 0 for k < m (init_vector) and defined as 0 for
     * k == m.  Adding +0.0*finite never changes the sum unless V[k] is
     * non-finite; restated without it. */
    for (int i = 0; i < n; ++i) {
        double t = 0.0;
        for (int j = 0; j < k; ++j) { double p = S->V[(size_t)j * n + i] * S->yk[j]; t = t + p; }
        S->Vy[i] = t;
    }
    for (int i = 0; i < n; ++i) S->x[i] = S->x_old[i] + S->Vy[i];
}

static void swapd(double **a, double **b) { double *t = *a; *a = *b; *b = t; }

static void o_init_residual(osolver *S) {
    int n = S->n;
    switch (S->method) {
    case M_J:   /* jacobi.hpp:79-83 */
    case M_GS:  /* gauss_seidel.hpp:76-80 */
    case M_SGS:
        o_compute_residual(&S->A, S->method == M_J ? S->x_old : S->x, S->b, S->residual, S->tmp);
        S->residual_norm = o_euclidean_vec_norm(S->residual, n);
        break;
    case M_CG:  /* cg.hpp:100-120 */
        o_compute_residual(&S->A, S->x_old, S->b, S->residual, S->tmp);
        o_precond(S, S->z_old, S->residual);
        o_copy_vector(S->p_old, S->z_old, n);
        o_copy_vector(S->r_old, S->residual, n);
        S->residual_norm = o_euclidean_vec_norm(S->residual, n);
        break;
    case M_BI:  /* bicgstab.hpp:146-169 */
        o_compute_residual(&S->A, S->x_old, S->b, S->residual, S->tmp);
        o_copy_vector(S->r_old, S->residual, n);
        S->residual_norm = o_euclidean_vec_norm(S->residual, n);
        o_precond(S, S->residual, S->residual);
        o_copy_vector(S->p_old, S->residual, n);
        S->rho_old = o_dot(S->r_old, S->residual, n);
        break;
    case M_GM:
        gm_init_residual(S);
        return;
    }
    o_copy_vector(S->residual_0, S->residual, n);   /* solver.hpp:148-151 */
    S->hist[S->hist_count++] = S->residual_norm;
}

static void o_iterate(osolver *S) {
    int n = S->n;
    switch (S->method) {
    case M_J:   /* jacobi.hpp:43-52 */
        o_spmv(n, S->A.rp, S->A.col, S->A.val, S->x_old, S->x_new);
        o_normalize_x(S->x_new, S->x_old, S->A_D, S->b, n);
        break;
    case M_GS:
    case M_SGS: /* gauss_seidel.hpp:26-52, 121-124 */
        o_spmv(n, S->Us.rp, S->Us.col, S->Us.val, S->x, S->tmp);
        o_subtract_vectors(S->tmp, S->b, S->tmp, n, 1.0);
        o_sptrsv(n, S->Ls.rp, S->Ls.col, S->Ls.val, S->x, S->A_D, S->tmp);
        if (S->method == M_SGS) {
            o_spmv(n, S->Ls.rp, S->Ls.col, S->Ls.val, S->x, S->tmp);
            o_subtract_vectors(S->tmp, S->b, S->tmp, n, 1.0);
            o_bsptrsv(n, S->Us.rp, S->Us.col, S->Us.val, S->x, S->A_D, S->tmp);
        }
        break;
    case M_CG: { /* cg.hpp:6-54 */
        o_spmv(n, S->A.rp, S->A.col, S->A.val, S->p_old, S->tmp);
        double rz = o_dot(S->r_old, S->z_old, n);
        double alpha = rz / o_dot(S->tmp, S->p_old, n);
        o_sum_vectors(S->x_new, S->x_old, S->p_old, n, alpha);
        o_subtract_vectors(S->r_new, S->r_old, S->tmp, n, alpha);
        o_precond(S, S->z_new, S->r_new);
        double beta = o_dot(S->r_new, S->z_new, n) / rz;
        o_sum_vectors(S->p_new, S->z_new, S->p_old, n, beta);
        break;
    }
    case M_BI: { /* bicgstab.hpp:8-83 */
        o_precond(S, S->y, S->p_old);
        o_spmv(n, S->A.rp, S->A.col, S->A.val, S->y, S->v);
        double alpha = S->rho_old / o_dot(S->residual_0, S->v, n);
        o_subtract_vectors(S->s, S->r_old, S->v, n, alpha);
        o_precond(S, S->s_tmp, S->s);
        o_spmv(n, S->A.rp, S->A.col, S->A.val, S->s_tmp, S->z);
        double zs = o_dot(S->z, S->s, n);
        double omega = zs / o_dot(S->z, S->z, n);
        o_sum_vectors(S->h, S->x_old, S->y, n, alpha);
        o_sum_vectors(S->x_new, S->h, S->s_tmp, n, omega);
        o_subtract_vectors(S->r_new, S->s, S->z, n, omega);
        S->rho_new = o_dot(S->residual_0, S->r_new, n);
        double beta = (S->rho_new / S->rho_old) * (alpha / omega);
        o_subtract_vectors(S->tmp, S->p_old, S->v, n, omega);
        o_sum_vectors(S->p_new, S->r_new, S->tmp, n, beta);
        swapd(&S->residual, &S->r_new);   /* bicgstab.hpp:177 */
        break;
    }
    case M_GM: { /* gmres.hpp:150-196 */
        int m = S->m;
        int k = S->iter_count - S->restart_count * m;
        o_spmv(n, S->A.rp, S->A.col, S->A.val, S->V + (size_t)k * n, S->w);
        o_precond(S, S->w, S->w);
        for (int j = 0; j <= k; ++j) {   /* orthogonalize_V gmres.hpp:6-53 */
            double hjk = o_dot(S->w, S->V + (size_t)j * n, n);
            S->H[k + j * m] = hjk;
            o_subtract_vectors(S->w, S->w, S->V + (size_t)j * n, n, hjk);
        }
        double hn = o_euclidean_vec_norm(S->w, n);
        S->H[(k + 1) * m + k] = hn;
        o_scale(S->V + (size_t)(k + 1) * n, S->w, 1.0 / hn, n);
        o_gmres_least_squares(k, m, S->J, S->H, S->H_tmp, S->Q, S->Q_tmp, S->R);
        S->residual_norm = o_gmres_update_g(k, m, S->Q, S->g, S->g_tmp, S->beta);
        break;
    }
    }
}

static void o_record_residual(osolver *S) {
    int n = S->n;
    switch (S->method) {
    case M_J:   /* jacobi.hpp:102-107 */
        o_compute_residual(&S->A, S->x_new, S->b, S->residual, S->tmp);
        S->residual_norm = o_euclidean_vec_norm(S->residual, n);
        break;
    case M_GS:
    case M_SGS: /* gauss_seidel.hpp:99-104 */
        o_compute_residual(&S->A, S->x, S->b, S->residual, S->tmp);
        S->residual_norm = o_euclidean_vec_norm(S->residual, n);
        break;
    case M_CG:  /* cg.hpp:162-166 */
        S->residual_norm = o_euclidean_vec_norm(S->r_new, n);
        break;
    case M_BI:  /* bicgstab.hpp:220-223 */
        S->residual_norm = o_euclidean_vec_norm(S->residual, n);
        break;
    default:
        break;
    }
    S->hist[S->hist_count++] = S->residual_norm;   /* solver.hpp:162-164 */
}

static void o_exchange(osolver *S) {
    switch (S->method) {
    case M_J:
        swapd(&S->x_old, &S->x_new);
        break;
    case M_CG:  /* cg.hpp:129-133 */
        swapd(&S->p_old, &S->p_new);
        swapd(&S->z_old, &S->z_new);
        swapd(&S->r_old, &S->r_new);
        swapd(&S->x_old, &S->x_new);
        break;
    case M_BI: { /* bicgstab.hpp:181-185 */
        swapd(&S->p_old, &S->p_new);
        swapd(&S->r_old, &S->residual);
        swapd(&S->x_old, &S->x_new);
        double t = S->rho_old; S->rho_old = S->rho_new; S->rho_new = t;
        break;
    }
    default:
        break;
    }
}

/* gmres.hpp:388-415 */
static void o_check_restart(osolver *S) {
    if (S->method != M_GM) return;
    int conv = S->residual_norm < S->stopping;
    int over = S->iter_count > o_max_iters;
    int cycle = (S->iter_count % S->m == 0) && (S->iter_count != 0);
    if (!conv && !over && cycle) {
        S->restarted = 1;
        gm_get_explicit_x(S);
        o_copy_vector(S->x_old, S->x, S->n);
        gm_init_structs(S);
        gm_init_residual(S);
        ++S->restart_count;
    }
}

/* solver.hpp:177-191 */
static int o_stop(const osolver *S) {
    int conv = fabs(S->residual_norm) < S->stopping;
    int over = S->iter_count >= (o_max_iters - S->restart_count);
    int div = fabs(S->residual_norm) > DBL_MAX || isnan(S->residual_norm);
    return conv || over || div;
}

/* Whole solve: preprocessing.hpp:26-100 (no -scale), solver_harness.hpp:7-61,
 * save_x_star (solver.hpp:153-159 + per-method overrides).
 * b / x0 may be NULL (B_VAL / INIT_X_VAL).  history needs 2*MAX_ITERS doubles.
 * out_int: iter_count, hist_count, converged, restart_count.
 * out_dbl: stopping_criteria, final true residual norm. */
int o_solve(int n, const int *rp, const int *col, const double *val, int method,
            int precond, int restart_len, const double *b, const double *x0,
            double *history, double *x_star, int *out_int, double *out_dbl) {
    osolver S;
    memset(&S, 0, sizeof S);
    S.method = method; S.precond = precond; S.m = restart_len; S.n = n;
    S.hist = history;
    S.residual_norm = DBL_MAX;
    S.A.n = n; S.A.nnz = rp[n]; S.A.rp = rp; S.A.col = col; S.A.val = val;

    /* allocate_structs / init_structs (solver.hpp:82-120) */
    S.x_star = vnew(n, 0.0); S.x_0 = vnew(n, O_INIT_X_VAL); S.b = vnew(n, O_B_VAL);
    S.tmp = vnew(n, 0.0); S.work = vnew(n, 0.0); S.residual = vnew(n, 0.0);
    S.residual_0 = vnew(n, 0.0); S.A_D = vnew(n, 1.0); S.A_D_inv = vnew(n, 0.0);
    S.L_D = vnew(n, 1.0); S.U_D = vnew(n, 1.0);
    if (b) o_copy_vector(S.b, b, n);
    if (x0) o_copy_vector(S.x_0, x0, n);

    /* factor_LU (LU_factors.hpp:900-934) */
    int *l_rp = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    int *u_rp = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    o_split_strict(n, rp, col, val, l_rp, NULL, NULL, u_rp, NULL, NULL, NULL, NULL);
    int nl = l_rp[n], nu = u_rp[n];
    int *l_col = (int *)malloc(sizeof(int) * (size_t)(nl + 1));
    int *u_col = (int *)malloc(sizeof(int) * (size_t)(nu + 1));
    double *l_val = (double *)malloc(sizeof(double) * (size_t)(nl + 1));
    double *u_val = (double *)malloc(sizeof(double) * (size_t)(nu + 1));
    o_split_strict(n, rp, col, val, l_rp, l_col, l_val, u_rp, u_col, u_val, S.A_D, S.A_D_inv);
    if (precond == P_ILU0)
        o_ilu0(n, rp, col, val, l_rp, l_col, l_val, S.L_D, u_rp, u_col, u_val, S.U_D);
    S.Ls.n = n; S.Ls.nnz = nl; S.Ls.rp = l_rp; S.Ls.col = l_col; S.Ls.val = l_val;
    S.Us.n = n; S.Us.nnz = nu; S.Us.rp = u_rp; S.Us.col = u_col; S.Us.val = u_val;

    int m = restart_len;
    switch (method) {
    case M_J:
        S.x_new = vnew(n, 0.0); S.x_old = vnew(n, 0.0); o_copy_vector(S.x_old, S.x_0, n);
        break;
    case M_GS: case M_SGS:
        S.x = vnew(n, 0.0); o_copy_vector(S.x, S.x_0, n);
        break;
    case M_CG:
        S.x_new = vnew(n, 0.0); S.x_old = vnew(n, 0.0); o_copy_vector(S.x_old, S.x_0, n);
        S.p_new = vnew(n, 0.0); S.p_old = vnew(n, 0.0); S.r_new = vnew(n, 0.0);
        S.r_old = vnew(n, 0.0); S.z_new = vnew(n, 0.0); S.z_old = vnew(n, 0.0);
        break;
    case M_BI:
        S.x_new = vnew(n, 0.0); S.x_old = vnew(n, 0.0); o_copy_vector(S.x_old, S.x_0, n);
        S.p_new = vnew(n, 0.0); S.p_old = vnew(n, 0.0); S.r_new = vnew(n, 0.0);
        S.r_old = vnew(n, 0.0); S.v = vnew(n, 0.0); S.h = vnew(n, 0.0); S.s = vnew(n, 0.0);
        S.s_tmp = vnew(n, 0.0); S.y = vnew(n, 0.0); S.z = vnew(n, 0.0);
        break;
    case M_GM:
        S.x = vnew(n, 0.0); S.x_old = vnew(n, 0.0);
        S.V = vnew(n * (m + 1), 0.0); S.Vy = vnew(n, 0.0); S.yk = vnew(m + 1, 0.0);
        S.H = vnew((m + 1) * m, 0.0); S.H_tmp = vnew((m + 1) * m, 0.0);
        S.J = vnew((m + 1) * (m + 1), 0.0); S.Q = vnew((m + 1) * (m + 1), 0.0);
        S.Q_tmp = vnew((m + 1) * (m + 1), 0.0); S.w = vnew(n, 0.0);
        S.R = vnew((m + 1) * m, 0.0); S.g = vnew(m + 1, 0.0); S.g_tmp = vnew(m + 1, 0.0);
        gm_init_structs(&S);
        break;
    default:
        return 1;
    }

    o_init_residual(&S);
    S.stopping = o_tol * S.residual_norm;   /* solver.hpp:173-175 */

    /* solver_harness.hpp:15-51 */
    do {
        o_iterate(&S);
        ++S.iter_count;
        if (S.iter_count % O_RES_CHECK_LEN == 0) o_record_residual(&S);
        o_exchange(&S);
        o_check_restart(&S);
    } while (!o_stop(&S));
    S.converged = S.residual_norm < S.stopping;

    /* save_x_star */
    double **xfinal = NULL;
    switch (method) {
    case M_GM: gm_get_explicit_x(&S); xfinal = &S.x; break;   /* gmres.hpp:377-381 */
    case M_GS: case M_SGS: xfinal = &S.x; break;
    default: xfinal = &S.x_old; break;
    }
    swapd(xfinal, &S.x_star);
    o_compute_residual(&S.A, S.x_star, S.b, S.residual, S.tmp);
    S.residual_norm = o_euclidean_vec_norm(S.residual, n);
    if (S.hist_count + 1 < 2 * O_MAX_ITERS) S.hist[S.hist_count + 1] = S.residual_norm;

    if (x_star) o_copy_vector(x_star, S.x_star, n);
    out_int[0] = S.iter_count; out_int[1] = S.hist_count;
    out_int[2] = S.converged; out_int[3] = S.restart_count;
    out_dbl[0] = S.stopping; out_dbl[1] = S.residual_norm;

    double *all[] = {S.x_star, S.x_0, S.b, S.tmp, S.work, S.residual, S.residual_0, S.A_D,
                     S.A_D_inv, S.L_D, S.U_D, S.x_new, S.x_old, S.p_new, S.p_old, S.z_new,
                     S.z_old, S.r_new, S.r_old, S.v, S.h, S.s, S.s_tmp, S.y, S.z, S.x, S.V,
                     S.Vy, S.yk, S.H, S.H_tmp, S.J, S.Q, S.Q_tmp, S.w, S.R, S.g, S.g_tmp};
    for (size_t i = 0; i < sizeof all / sizeof all[0]; ++i) free(all[i]);
    free(l_rp); free(u_rp); free(l_col); free(u_col); free(l_val); free(u_val);
    return 0;
}
