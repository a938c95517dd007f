// host/host_capi.cpp -- C entry points of the HOST library (libbis_host.so):
// the reference-shaped Solver/harness stack driven from plain C (ctypes in the
// tests and bench.py).  Device work goes through the C-ABI of libbis_b200.so.
#include "run.hpp"

#include <cstring>

static thread_local std::string g_host_err;

extern "C" {

const char *bis_host_last_error(void) { return g_host_err.c_str(); }

// parse_cli round trip (CPU-only test of the host logic): fills
// out[0]=method, out[1]=precond, out[2]=restart_length, out[3]=num_scale.
int bis_host_parse_cli(int argc, char **argv, int *out) {
    try {
        Args a;
        parse_cli(&a, argc, argv);
        out[0] = static_cast<int>(a.method);
        out[1] = static_cast<int>(a.preconditioner);
        out[2] = a.restart_length;
        out[3] = a.num_scale ? 1 : 0;
        return 0;
    } catch (const std::exception &e) {
        g_host_err = e.what();
        return 1;
    }
}

// Host preprocessing (CPU-only testable): strict split + diagonals (+ ILU(0)).
// Call with l_col == NULL to obtain the sizes first.
int bis_host_factor(int n, const int *rp, const int *col, const double *val, int precond, int *l_rp,
                    int *l_col, double *l_val, int *u_rp, int *u_col, double *u_val, double *A_D,
                    double *A_D_inv, double *L_D, double *U_D) {
    try {
        MatrixCRS A(n, n, rp[n]);
        std::memcpy(A.row_ptr, rp, sizeof(int) * (n + 1));
        std::memcpy(A.col, col, sizeof(int) * rp[n]);
        std::memcpy(A.val, val, sizeof(double) * rp[n]);
        MatrixCRS L, U;
        std::vector<double> d(n, 1.0), di(n, 0.0), ld(n, 1.0), ud(n, 1.0);
        factor_LU(&A, d.data(), di.data(), &L, ld.data(), &U, ud.data(), static_cast<PrecondType>(precond));
        std::memcpy(l_rp, L.row_ptr, sizeof(int) * (n + 1));
        std::memcpy(u_rp, U.row_ptr, sizeof(int) * (n + 1));
        if (l_col) {
            std::memcpy(l_col, L.col, sizeof(int) * L.nnz);
            std::memcpy(l_val, L.val, sizeof(double) * L.nnz);
            std::memcpy(u_col, U.col, sizeof(int) * U.nnz);
            std::memcpy(u_val, U.val, sizeof(double) * U.nnz);
            std::memcpy(A_D, d.data(), sizeof(double) * n);
            std::memcpy(A_D_inv, di.data(), sizeof(double) * n);
            std::memcpy(L_D, ld.data(), sizeof(double) * n);
            std::memcpy(U_D, ud.data(), sizeof(double) * n);
        }
        return 0;
    } catch (const std::exception &e) {
        g_host_err = e.what();
        return 1;
    }
}

// Host matrix generators / reader (CPU-only testable).  kind 0: .mtx path in
// `name`; otherwise a generator name.  Two-phase: sizes, then fill.
static std::unique_ptr<MatrixCRS> g_mat;
int bis_host_matrix_begin(const char *name, int *n, int *nnz) {
    try {
        Args a;
        a.matrix_file_name = name;
        std::unique_ptr<DeviceCRS> none;
        obtain_matrix(&a, nullptr, true, g_mat, none);
        *n = g_mat->n_rows;
        *nnz = g_mat->nnz;
        return 0;
    } catch (const std::exception &e) {
        g_host_err = e.what();
        return 1;
    }
}
int bis_host_matrix_fetch(int *rp, int *col, double *val) {
    if (!g_mat) return 1;
    std::memcpy(rp, g_mat->row_ptr, sizeof(int) * (g_mat->n_rows + 1));
    std::memcpy(col, g_mat->col, sizeof(int) * g_mat->nnz);
    std::memcpy(val, g_mat->val, sizeof(double) * g_mat->nnz);
    g_mat.reset();
    return 0;
}

// GMRES host pieces (CPU-only testable against the oracle).
void bis_host_gmres_least_squares(int k, int m, double *J, double *H, double *H_tmp, double *Q,
                                  double *Q_tmp, double *R) {
    least_squares(k, m, J, H, H_tmp, Q, Q_tmp, R);
}
double bis_host_gmres_update_g(int k, int m, double *Q, double *g, double *g_tmp, double beta) {
    double rn = 0.0;
    update_g(k, m, Q, g, g_tmp, rn, beta);
    return rn;
}

// A whole solve on the device.  Matrix: host CRS arrays (rp != NULL) or a
// generator / file name.  b / x0: host vectors or NULL (B_VAL / INIT_X_VAL).
// max_iters <= 0 keeps MAX_ITERS; tol <= 0 keeps TOL.
// history: 2*MAX_ITERS doubles; iter_time likewise (may be NULL); x_star: n
// doubles (may be NULL).
// out_int: iter_count, history count, converged, restart_count, kernel launches.
// out_dbl: stopping criteria, final true residual, solve seconds, preprocessing
// seconds, mean seconds per iteration (harness table, first iteration excluded).
int bis_host_solve(bis_context *dev, const char *matrix_name, int n, const int *rp, const int *col,
                   const double *val, int method, int precond, int restart_len, const double *b,
                   const double *x0, int max_iters, double tol, int quiet, double *history,
                   double *iter_time, double *x_star, int *out_int, double *out_dbl) {
    try {
        Args args;
        args.matrix_file_name = matrix_name ? matrix_name : "in-memory";
        args.method = static_cast<SolverType>(method);
        args.preconditioner = static_cast<PrecondType>(precond);
        args.restart_length = restart_len;
        args.quiet = quiet != 0;
        Timers timers;
        std::unique_ptr<Solver> solver = make_solver(&args, dev);
        if (max_iters > 0) {
            if (max_iters > MAX_ITERS) bis_fatal("max_iters exceeds MAX_ITERS");
            solver->max_iters = max_iters;
        }
        if (tol > 0) solver->tolerance = tol;
        std::unique_ptr<MatrixCRS> A;
        std::unique_ptr<DeviceCRS> dA;
        if (rp) {
            A = std::make_unique<MatrixCRS>(n, n, rp[n]);
            std::memcpy(A->row_ptr, rp, sizeof(int) * (n + 1));
            std::memcpy(A->col, col, sizeof(int) * rp[n]);
            std::memcpy(A->val, val, sizeof(double) * rp[n]);
        } else {
            obtain_matrix(&args, dev, solver->needs_triangular_factors(), A, dA);
        }
        int64_t info0[8];
        BIS_OK(bis_context_info(dev, info0));
        TIME(timers.preprocessing, preprocessing(&args, solver.get(), &timers, A, std::move(dA), b, x0))
        TIME(timers.solve, solve(&args, solver.get(), &timers))
        int64_t info1[8];
        BIS_OK(bis_context_info(dev, info1));
        const int cnt = solver->collected_residual_norms_count;
        const int ncopy = std::min(cnt + 2, 2 * MAX_ITERS);
        for (int i = 0; i < ncopy; ++i) {
            history[i] = solver->collected_residual_norms[i];
            if (iter_time) iter_time[i] = solver->time_per_iteration[i];
        }
        if (x_star) BIS_OK(bis_vector_download(dev, x_star, solver->x_star, solver->N));
        out_int[0] = solver->iter_count;
        out_int[1] = cnt;
        out_int[2] = solver->convergence_flag ? 1 : 0;
        out_int[3] = solver->gmres_restart_count;
        out_int[4] = (int)(info1[3] - info0[3]);
        out_dbl[0] = solver->stopping_criteria;
        out_dbl[1] = solver->residual_norm;
        out_dbl[2] = timers.solve_time.get_wtime();
        out_dbl[3] = timers.preprocessing_time.get_wtime();
        double acc = 0.0;
        int m = 0;
        for (int i = 3; i <= cnt; ++i) {   // time_per_iteration[i+1] pairs with history[i]; skip the first
            acc += solver->time_per_iteration[i];
            ++m;
        }
        out_dbl[4] = m ? acc / m : 0.0;
        if (!quiet) postprocessing(&args, solver.get(), &timers);
        return 0;
    } catch (const std::exception &e) {
        g_host_err = e.what();
        return 1;
    }
}

} // extern "C"
