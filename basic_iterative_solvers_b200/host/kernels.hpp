// host/kernels.hpp -- the reference's free-function kernel API (kernels.hpp)
// re-pointed at the C-ABI.  Same names, same argument meaning; `double *`
// arguments are DEVICE addresses; the `Interface *smax` slot of the reference
// (kernels.hpp:44-46) is the `bis_context *`.  A failing C-ABI call is fatal
// (there is no CPU fallback), mirroring the reference's exit-on-error.
#pragma once

#include "common.hpp"
#include "sparse_matrix.hpp"

using Interface = bis_context;

inline void bis_ok(int rc, const char *what) {
    if (rc != 0) bis_fatal(std::string(what) + ": " + bis_last_error());
}
#define BIS_OK(call) bis_ok((call), #call)

// ---- allocation helpers: new double[N] / delete[] of the reference -----------
inline double *dev_new(Interface *dev, int64_t n) {
    double *p = nullptr;
    BIS_OK(bis_vector_alloc(dev, n, &p));
    return p;
}
inline void dev_delete(Interface *dev, double *&p) {
    if (p) bis_vector_free(dev, p);
    p = nullptr;
}

// ---- kernels.hpp:44-117 ---------------------------------------------------------
inline void spmv(Interface *dev, const DeviceCRS *A, const double *x, double *y) {
    BIS_OK(bis_spmv(dev, A->handle, x, y));
}
inline void sptrsv(Interface *dev, const DeviceCRS *L, double *x, const double *D, const double *b) {
    BIS_OK(bis_sptrsv(dev, L->handle, x, D, b));
}
inline void bsptrsv(Interface *dev, const DeviceCRS *U, double *x, const double *D, const double *b) {
    BIS_OK(bis_bsptrsv(dev, U->handle, x, D, b));
}
// ---- kernels.hpp:119-153 --------------------------------------------------------
inline void subtract_vectors(Interface *dev, double *r, const double *a, const double *b, int64_t N,
                             double scale = 1.0) {
    BIS_OK(bis_subtract_vectors(dev, r, a, b, N, scale));
}
inline void sum_vectors(Interface *dev, double *r, const double *a, const double *b, int64_t N,
                        double scale = 1.0) {
    BIS_OK(bis_sum_vectors(dev, r, a, b, N, scale));
}
inline void elemwise_mult_vectors(Interface *dev, double *r, const double *a, const double *b,
                                  int64_t N, double scale = 1.0) {
    BIS_OK(bis_elemwise_mult_vectors(dev, r, a, b, N, scale));
}
inline void elemwise_div_vectors(Interface *dev, double *r, const double *a, const double *b,
                                 int64_t N, double scale = 1.0) {
    BIS_OK(bis_elemwise_div_vectors(dev, r, a, b, N, scale));
}
// ---- kernels.hpp:155-257 --------------------------------------------------------
inline void compute_residual(Interface *dev, const DeviceCRS *A, const double *x, const double *b,
                             double *residual, double *tmp) {
    BIS_OK(bis_compute_residual(dev, A->handle, x, b, residual, tmp));
}
inline double euclidean_vec_norm(Interface *dev, const double *v, int64_t N) {
    double r = 0.0;
    BIS_OK(bis_euclidean_vec_norm(dev, v, N, &r));
    return r;
}
inline double dot(Interface *dev, const double *a, const double *b, int64_t N) {
    double r = 0.0;
    BIS_OK(bis_dot(dev, a, b, N, &r));
    return r;
}
inline void scale(Interface *dev, double *r, const double *v, double s, int64_t N) {
    BIS_OK(bis_scale(dev, r, v, s, N));
}
inline void init_vector(Interface *dev, double *v, double val, int64_t N) {
    BIS_OK(bis_init_vector(dev, v, val, N));
}
inline void copy_vector(Interface *dev, double *out, const double *in, int64_t N) {
    BIS_OK(bis_copy_vector(dev, out, in, N));
}
// ---- kernels.hpp:336-414 --------------------------------------------------------
inline void apply_preconditioner(Interface *dev, const PrecondType preconditioner, int64_t N,
                                 const DeviceCRS *L_strict, const DeviceCRS *U_strict, double *A_D,
                                 double *A_D_inv, double *L_D, double *U_D, double *output,
                                 double *input, double *tmp, double *work) {
    BIS_OK(bis_apply_preconditioner(dev, static_cast<int>(preconditioner), N,
                                    L_strict ? L_strict->handle : nullptr,
                                    U_strict ? U_strict->handle : nullptr, A_D, A_D_inv, L_D, U_D,
                                    output, input, tmp, work));
}
// value of a device scalar slot (one host<->device synchronisation)
inline double scalar(Interface *dev, int slot) {
    double v = 0.0;
    BIS_OK(bis_scalar_get(dev, slot, 1, &v));
    return v;
}

// ---- tiny dense helpers of GMRES, host side, restated literally ------------------
// (kernels.hpp:222-250, 273-310; rounding as the reference build emits it, see
// oracle/port/bis_oracle.c)
inline void init_dense_identity_matrix(double *mat, int n_rows, int n_cols) {
    for (int i = 0; i < n_rows; ++i)
        for (int j = 0; j < n_cols; ++j) mat[n_cols * i + j] = (i == j) ? 1.0 : 0.0;
}
inline void copy_dense_matrix(double *dst, const double *src, int n_rows, int n_cols) {
    for (int i = 0; i < n_rows * n_cols; ++i) dst[i] = src[i];
}
inline void dgemm_transpose2(const double *A, const double *B, double *C, int n_rows_A, int n_cols_A,
                             int n_cols_B) {
    for (int i = 0; i < n_rows_A; ++i)
        for (int j = 0; j < n_cols_B; ++j) {
            double t = 0.0;
            for (int k = 0; k < n_cols_A; ++k) t = std::fma(A[i * n_cols_A + k], B[k * n_cols_B + j], t);
            C[i * n_cols_B + j] = t;
        }
}
inline void dgemv(const double *A, const double *x, double *y, int n_rows_A, int n_cols_A) {
    for (int i = 0; i < n_rows_A; ++i) {
        double acc = 0.0;   // host code is built with -ffp-contract=off: mul, then add
        for (int j = 0; j < n_cols_A; ++j) acc = acc + A[i * n_cols_A + j] * x[j];
        y[i] = acc;
    }
}
