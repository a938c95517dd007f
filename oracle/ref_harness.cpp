// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Thin extern "C" shim around the UNMODIFIED reference sources, which are
// #included from where they lie under /root/reference (never copied into this
// repo).  Built by oracle/Makefile into oracle/_ref/libbis_ref.so.  It exists
// so that tests/ and bench.py's cpu_baseline / --impl reference legs can run
// the reference's own kernels.hpp / methods/*.hpp / solver_harness.hpp on
// arrays handed over from Python (ctypes), and so that the plain-C
// restatement in oracle/port/ can be pinned against the real thing.
//
// The reference's non-inline free functions (methods/*.hpp, preprocessing.hpp,
// solver_harness.hpp) may only be included from one TU: this is that TU.
//
// ILU(0): the stock no-SMAX build routes factor_ILU0 to an empty stub
// (utilities/LU_factors.hpp:765-779), so `ilu0_old != 0` runs the
// reference's working factor_ILU0_old (LU_factors.hpp:320-539) in the place
// where factor_LU (LU_factors.hpp:900-934) would have called factor_ILU0.

#include "common.hpp"
#include "methods/bicgstab.hpp"
#include "methods/cg.hpp"
#include "methods/gauss_seidel.hpp"
#include "methods/gmres.hpp"
#include "methods/jacobi.hpp"
#include "postprocessing.hpp"
#include "preprocessing.hpp"
#include "solver_harness.hpp"
#include "utilities/utilities.hpp"

#include <cstring>
#include <omp.h>
#include <sstream>

namespace {

std::unique_ptr<MatrixCRS> make_crs(int n_rows, int n_cols, int nnz,
                                    const int *rp, const int *col,
                                    const double *val) {
    auto A = std::make_unique<MatrixCRS>(n_rows, n_cols, nnz);
    std::memcpy(A->row_ptr, rp, sizeof(int) * (n_rows + 1));
    if (nnz > 0) {
        std::memcpy(A->col, col, sizeof(int) * nnz);
        std::memcpy(A->val, val, sizeof(double) * nnz);
    }
    return A;
}

// A non-owning view: MatrixCRS's destructor delete[]s its arrays, so hand the
// pointers over and take them back before the object dies.
struct CrsView {
    MatrixCRS m;
    CrsView(int n_rows, int n_cols, int nnz, const int *rp, const int *col,
            const double *val) {
        m.n_rows = n_rows;
        m.n_cols = n_cols;
        m.nnz = nnz;
        m.row_ptr = const_cast<int *>(rp);
        m.col = const_cast<int *>(col);
        m.val = const_cast<double *>(val);
    }
    ~CrsView() {
        m.row_ptr = nullptr;
        m.col = nullptr;
        m.val = nullptr;
    }
};

Solver *make_solver(const Args *a) {
    // main.cpp:22-44
    switch (a->method) {
    case SolverType::Jacobi:
        return new JacobiSolver(a);
    case SolverType::GaussSeidel:
        return new GaussSeidelSolver(a);
    case SolverType::SymmetricGaussSeidel:
        return new SymmetricGaussSeidelSolver(a);
    case SolverType::ConjugateGradient:
        return new ConjugateGradientSolver(a);
    case SolverType::GMRES:
        return new GMRESSolver(a);
    case SolverType::BiCGSTAB:
        return new BiCGSTABSolver(a);
    }
    return nullptr;
}

struct QuietStdout {
    std::streambuf *old;
    std::ostringstream sink;
    bool on;
    explicit QuietStdout(bool quiet) : old(nullptr), on(quiet) {
        if (on)
            old = std::cout.rdbuf(sink.rdbuf());
    }
    ~QuietStdout() {
        if (on)
            std::cout.rdbuf(old);
    }
};

} // namespace

extern "C" {

int ref_omp_max_threads() { return omp_get_max_threads(); }
void ref_omp_set_threads(int n) { omp_set_num_threads(n); }

// ---- kernels.hpp ---------------------------------------------------------
void ref_spmv(int n_rows, int n_cols, int nnz, const int *rp, const int *col,
              const double *val, const double *x, double *y) {
    CrsView A(n_rows, n_cols, nnz, rp, col, val);
    native_spmv(&A.m, x, y);
}
void ref_sptrsv(int n, int nnz, const int *rp, const int *col,
                const double *val, double *x, const double *D,
                const double *b) {
    CrsView L(n, n, nnz, rp, col, val);
    native_sptrsv(&L.m, x, D, b);
}
void ref_bsptrsv(int n, int nnz, const int *rp, const int *col,
                 const double *val, double *x, const double *D,
                 const double *b) {
    CrsView U(n, n, nnz, rp, col, val);
    native_bsptrsv(&U.m, x, D, b);
}
void ref_subtract_vectors(double *r, const double *a, const double *b, int N,
                          double s) {
    subtract_vectors(r, a, b, N, s);
}
void ref_sum_vectors(double *r, const double *a, const double *b, int N,
                     double s) {
    sum_vectors(r, a, b, N, s);
}
void ref_elemwise_mult_vectors(double *r, const double *a, const double *b,
                               int N, double s) {
    elemwise_mult_vectors(r, a, b, N, s);
}
void ref_elemwise_div_vectors(double *r, const double *a, const double *b,
                              int N, double s) {
    elemwise_div_vectors(r, a, b, N, s);
}
void ref_scale(double *r, const double *v, double s, int N) {
    scale(r, v, s, N);
}
void ref_copy_vector(double *o, const double *i, int N) {
    copy_vector(o, i, N);
}
double ref_dot(const double *a, const double *b, int N) { return dot(a, b, N); }
double ref_euclidean_vec_norm(const double *v, int N) {
    return euclidean_vec_norm(v, N);
}
void ref_normalize_x(double *x_new, const double *x_old, const double *D,
                     const double *b, int n) {
    normalize_x(x_new, x_old, D, b, n);
}
void ref_compute_residual(int n, int nnz, const int *rp, const int *col,
                          const double *val, const double *x, const double *b,
                          double *r, double *tmp) {
    CrsView A(n, n, nnz, rp, col, val);
    compute_residual(&A.m, x, b, r, tmp);
}
// precond: PrecondType as int (common.hpp:38-47)
void ref_apply_preconditioner(int precond, int N, int nnz_l, const int *l_rp,
                              const int *l_col, const double *l_val, int nnz_u,
                              const int *u_rp, const int *u_col,
                              const double *u_val, double *A_D,
                              double *A_D_inv, double *L_D, double *U_D,
                              double *out, double *in, double *tmp,
                              double *work) {
    CrsView L(N, N, nnz_l, l_rp, l_col, l_val);
    CrsView U(N, N, nnz_u, u_rp, u_col, u_val);
    apply_preconditioner(static_cast<PrecondType>(precond), N, &L.m, &U.m, A_D,
                         A_D_inv, L_D, U_D, out, in, tmp, work);
}

// ---- GMRES small dense pieces (methods/gmres.hpp:55-148) -----------------
void ref_gmres_least_squares(int N, int k, int m, double *J, double *H,
                             double *H_tmp, double *Q, double *Q_tmp,
                             double *R) {
    Timers t;
    init_timers(&t);
    least_squares(&t, N, k, m, J, H, H_tmp, Q, Q_tmp, R);
}
double ref_gmres_update_g(int N, int k, int m, double *Q, double *g,
                          double *g_tmp, double beta) {
    Timers t;
    init_timers(&t);
    double rn = 0.0;
    update_g(&t, N, k, m, Q, g, g_tmp, rn, beta);
    return rn;
}

// ---- factorisation: LU_factors.hpp ---------------------------------------
// Two-phase: ref_factor_begin() runs factor_LU (and factor_ILU0_old when
// asked) and parks the result; the caller reads the sizes, allocates, and
// ref_factor_fetch() copies out and frees.
struct FactorResult {
    std::unique_ptr<MatrixCRS> A, L, Ls, U, Us;
    std::vector<double> A_D, A_D_inv, L_D, U_D;
};
static FactorResult *g_factor = nullptr;

int ref_factor_begin(int n, int nnz, const int *rp, const int *col,
                     const double *val, int precond, int ilu0_old,
                     int *nnz_l_strict, int *nnz_u_strict) {
    delete g_factor;
    g_factor = new FactorResult;
    FactorResult &f = *g_factor;
    f.A = make_crs(n, n, nnz, rp, col, val);
    f.L = std::make_unique<MatrixCRS>();
    f.Ls = std::make_unique<MatrixCRS>();
    f.U = std::make_unique<MatrixCRS>();
    f.Us = std::make_unique<MatrixCRS>();
    f.A_D.assign(n, 1.0);
    f.A_D_inv.assign(n, 0.0);
    f.L_D.assign(n, 1.0);
    f.U_D.assign(n, 1.0);
    Timers t;
    init_timers(&t);
    PrecondType p = static_cast<PrecondType>(precond);
    if (p == PrecondType::ILU0 && ilu0_old) {
        // factor_LU (LU_factors.hpp:900-934) with factor_ILU0 ->
        // factor_ILU0_old
        factor_LU(&t, f.A.get(), f.A_D.data(), f.A_D_inv.data(), f.L.get(),
                  f.Ls.get(), f.L_D.data(), f.U.get(), f.Us.get(),
                  f.U_D.data(), PrecondType::None);
        factor_ILU0_old(&t, f.A.get(), f.L.get(), f.Ls.get(), f.L_D.data(),
                        f.U.get(), f.Us.get(), f.U_D.data());
        peel_diag_crs(f.U.get(), f.U_D.data());
    } else {
        factor_LU(&t, f.A.get(), f.A_D.data(), f.A_D_inv.data(), f.L.get(),
                  f.Ls.get(), f.L_D.data(), f.U.get(), f.Us.get(),
                  f.U_D.data(), p);
    }
    *nnz_l_strict = f.Ls->nnz;
    *nnz_u_strict = f.Us->nnz;
    return 0;
}

void ref_factor_fetch(int *l_rp, int *l_col, double *l_val, int *u_rp,
                      int *u_col, double *u_val, double *A_D, double *A_D_inv,
                      double *L_D, double *U_D) {
    FactorResult &f = *g_factor;
    int n = f.A->n_rows;
    std::memcpy(l_rp, f.Ls->row_ptr, sizeof(int) * (n + 1));
    std::memcpy(u_rp, f.Us->row_ptr, sizeof(int) * (n + 1));
    if (f.Ls->nnz) {
        std::memcpy(l_col, f.Ls->col, sizeof(int) * f.Ls->nnz);
        std::memcpy(l_val, f.Ls->val, sizeof(double) * f.Ls->nnz);
    }
    if (f.Us->nnz) {
        std::memcpy(u_col, f.Us->col, sizeof(int) * f.Us->nnz);
        std::memcpy(u_val, f.Us->val, sizeof(double) * f.Us->nnz);
    }
    std::memcpy(A_D, f.A_D.data(), sizeof(double) * n);
    std::memcpy(A_D_inv, f.A_D_inv.data(), sizeof(double) * n);
    std::memcpy(L_D, f.L_D.data(), sizeof(double) * n);
    std::memcpy(U_D, f.U_D.data(), sizeof(double) * n);
    delete g_factor;
    g_factor = nullptr;
}

// ---- whole solves --------------------------------------------------------
// method: SolverType as int (common.hpp:49-56); precond: PrecondType as int.
// b / x0: N-vectors or NULL (-> the reference's B_VAL / INIT_X_VAL fill,
// solver.hpp:98-108).  use_ref_preprocessing != 0 calls the reference's own
// preprocessing() verbatim (requires b == x0 == NULL and !ilu0_old);
// otherwise the same call sequence (preprocessing.hpp:26-100, no-SMAX branch)
// is issued here so that b/x0 can be set and factor_ILU0_old selected.
// history must hold 2*MAX_ITERS doubles (solver.hpp:64).
// out_int[0]=iter_count (before postprocessing's GMRES adjustment),
// out_int[1]=collected_residual_norms_count, out_int[2]=convergence_flag,
// out_int[3]=gmres_restart_count.
// out_dbl[0]=stopping_criteria, out_dbl[1]=final residual_norm (true residual
// from save_x_star), out_dbl[2]=iterate_time [s], out_dbl[3]=spmv_time [s],
// out_dbl[4]=precond_time [s], out_dbl[5]=solve_time [s].
// Bounded runs for bench.py's CPU-baseline leg: caps the public `max_iters`
// member of the reference's Solver (solver.hpp:29); 0 restores MAX_ITERS.
static int g_ref_max_iters = 0;
void ref_set_max_iters(int max_iters) { g_ref_max_iters = max_iters; }

int ref_solve(int n, int nnz, const int *rp, const int *col, const double *val,
              int method, int precond, int restart_len, int num_scale,
              int ilu0_old, int use_ref_preprocessing, const double *b,
              const double *x0, int quiet, double *history, double *iter_time,
              double *x_star, int *out_int, double *out_dbl) {
    Args args;
    args.matrix_file_name = "in-memory";
    args.method = static_cast<SolverType>(method);
    args.preconditioner = static_cast<PrecondType>(precond);
    args.restart_length = restart_len;
    args.num_scale = num_scale != 0;

    Timers timers;
    init_timers(&timers);
    QuietStdout q(quiet != 0);

    Solver *solver = make_solver(&args);
    if (!solver)
        return 1;
    if (g_ref_max_iters > 0 && g_ref_max_iters <= MAX_ITERS)
        solver->max_iters = g_ref_max_iters;
    std::unique_ptr<MatrixCRS> A = make_crs(n, n, nnz, rp, col, val);

    if (use_ref_preprocessing) {
        if (b || x0 || ilu0_old) {
            delete solver;
            return 2;
        }
        TIME(timers.preprocessing, preprocessing(&args, solver, &timers, A))
    } else {
        // preprocessing.hpp:26-100 restated as calls (no-SMAX branch)
        solver->allocate_structs(A->n_cols);
        if (b)
            copy_vector(solver->b, b, n);
        if (x0)
            copy_vector(solver->x_0, x0, n);
        solver->init_structs(A->n_cols);
        solver->A = std::move(A);
        if (solver->num_scale) {
            extract_scale(solver->A.get(), solver->A_D_scale);
            scale_mat(solver->A.get(), solver->A_D_scale);
            scale_vec(solver->x_0, solver->A_D_scale, n);
            scale_vec(solver->b, solver->A_D_scale, n);
        }
        solver->L = std::make_unique<MatrixCRS>();
        solver->L_strict = std::make_unique<MatrixCRS>();
        solver->U = std::make_unique<MatrixCRS>();
        solver->U_strict = std::make_unique<MatrixCRS>();
        if (solver->preconditioner == PrecondType::ILU0 && ilu0_old) {
            factor_LU(&timers, solver->A.get(), solver->A_D, solver->A_D_inv,
                      solver->L.get(), solver->L_strict.get(), solver->L_D,
                      solver->U.get(), solver->U_strict.get(), solver->U_D,
                      PrecondType::None);
            factor_ILU0_old(&timers, solver->A.get(), solver->L.get(),
                            solver->L_strict.get(), solver->L_D,
                            solver->U.get(), solver->U_strict.get(),
                            solver->U_D);
            peel_diag_crs(solver->U.get(), solver->U_D);
        } else {
            factor_LU(&timers, solver->A.get(), solver->A_D, solver->A_D_inv,
                      solver->L.get(), solver->L_strict.get(), solver->L_D,
                      solver->U.get(), solver->U_strict.get(), solver->U_D,
                      solver->preconditioner);
        }
        solver->init_residual();
        solver->init_stopping_criteria();
    }

    TIME(timers.solve, solve(&args, solver, &timers))

    int cnt = solver->collected_residual_norms_count;
    // save_x_star parks the final true residual at [count+1] (solver.hpp:158)
    int ncopy = cnt + 2;
    if (ncopy > 2 * MAX_ITERS)
        ncopy = 2 * MAX_ITERS;
    for (int i = 0; i < ncopy; ++i) {
        history[i] = solver->collected_residual_norms[i];
        if (iter_time)
            iter_time[i] = solver->time_per_iteration[i];
    }
    if (x_star)
        std::memcpy(x_star, solver->x_star, sizeof(double) * n);
    out_int[0] = solver->iter_count;
    out_int[1] = cnt;
    out_int[2] = solver->convergence_flag ? 1 : 0;
    out_int[3] = solver->gmres_restart_count;
    out_dbl[0] = solver->stopping_criteria;
    out_dbl[1] = solver->residual_norm;
    out_dbl[2] = (double)timers.iterate_time->get_wtime();
    out_dbl[3] = (double)timers.spmv_time->get_wtime();
    out_dbl[4] = (double)timers.precond_time->get_wtime();
    out_dbl[5] = (double)timers.solve_time->get_wtime();
    if (!quiet)
        postprocessing(&args, solver, &timers);
    delete solver;
    return 0;
}

// Read a MatrixMarket file with the reference's reader and hand back CRS
// (sparse_matrix.hpp:225-350, utilities/utilities.hpp:326-367).
static std::unique_ptr<MatrixCRS> g_mtx;
int ref_read_mtx_begin(const char *path, int *n, int *nnz) {
    try {
        auto coo = std::make_unique<MatrixCOO>();
        coo->read_from_mtx(path);
        g_mtx = std::make_unique<MatrixCRS>();
        convert_coo_to_crs(coo.get(), g_mtx.get());
    } catch (const std::exception &e) {
        fprintf(stderr, "ref_read_mtx: %s\n", e.what());
        return 1;
    }
    *n = g_mtx->n_rows;
    *nnz = g_mtx->nnz;
    return 0;
}
void ref_read_mtx_fetch(int *rp, int *col, double *val) {
    std::memcpy(rp, g_mtx->row_ptr, sizeof(int) * (g_mtx->n_rows + 1));
    std::memcpy(col, g_mtx->col, sizeof(int) * g_mtx->nnz);
    std::memcpy(val, g_mtx->val, sizeof(double) * g_mtx->nnz);
    g_mtx.reset();
}

} // extern "C"
