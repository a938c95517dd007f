// host/lu_factors.hpp -- HOST restatement of the factor step: strict L/U split + diagonal (split_LU_new / peel_diag_crs_new,
// reference utilities/LU_factors.hpp:122-309, 827-869) and ILU(0)
// (factor_ILU0_old, LU_factors.hpp:320-539 -- the working pure-C++ routine;
// the stock dispatcher picks an SMAX-only stub).
//
// Only what the native kernels consume is produced:
// L_strict, U_strict, A_D, A_D_inv, L_D, U_D.  The solver stack no longer calls this: preprocessing()
// splits / factors on the device (csrc/bis_factor.cu).  It stays as the CPU-testable statement of the
// rules (bis_host_factor; pinned bit for bit to the compiled reference's factors by
// tests/test_host_cpu.py) that the device factorisation is checked against on the GPU.
#pragma once

#include "common.hpp"
#include "sparse_matrix.hpp"

// Strictly lower / strictly upper copies of A, entries in A's within-row order,
// and the diagonal D (+ 1/D).  Fatal like SanityChecker::no_diag / zero_diag
// (common.hpp:388-396) when a row has no or a (near-)zero diagonal.
inline void split_LU(const MatrixCRS *A, MatrixCRS *L_strict, MatrixCRS *U_strict, double *D,
                     double *D_inv) {
    const int n = A->n_rows;
    std::vector<int> lrp(n + 1, 0), urp(n + 1, 0);
    for (int i = 0; i < n; ++i) {
        int nl = 0, nu = 0;
        for (int k = A->row_ptr[i]; k < A->row_ptr[i + 1]; ++k) {
            if (A->col[k] < i) ++nl;
            else if (A->col[k] > i) ++nu;
        }
        lrp[i + 1] = lrp[i] + nl;
        urp[i + 1] = urp[i] + nu;
    }
    auto reset = [&](MatrixCRS *M, const std::vector<int> &rp) {
        delete[] M->row_ptr;
        delete[] M->col;
        delete[] M->val;
        M->n_rows = M->n_cols = n;
        M->nnz = rp[n];
        M->row_ptr = new int[n + 1];
        M->col = new int[M->nnz > 0 ? M->nnz : 1];
        M->val = new double[M->nnz > 0 ? M->nnz : 1];
        std::copy(rp.begin(), rp.end(), M->row_ptr);
    };
    reset(L_strict, lrp);
    reset(U_strict, urp);
    for (int i = 0; i < n; ++i) {
        int pl = lrp[i], pu = urp[i];
        bool have_diag = false;
        for (int k = A->row_ptr[i]; k < A->row_ptr[i + 1]; ++k) {
            const int c = A->col[k];
            const double v = A->val[k];
            if (c < i) {
                L_strict->col[pl] = c;
                L_strict->val[pl++] = v;
            } else if (c > i) {
                U_strict->col[pu] = c;
                U_strict->val[pu++] = v;
            } else {
                have_diag = true;
                if (std::abs(v) < 1e-16)
                    bis_fatal("Zero detected on diagonal at row index " + std::to_string(i));
                if (D) D[i] = v;
                if (D_inv) D_inv[i] = 1.0 / v;
            }
        }
        if (!have_diag) bis_fatal("No diagonal to extract at row index " + std::to_string(i));
    }
}

// ILU(0), row-wise IKJ on A's pattern.  Writes the strict factors over
// L_strict/U_strict (same pattern as A's strict parts but columns ascending),
// L_D = 1, U_D = diag(U).  Rounding follows the reference build: the row update
// is one fused multiply-add (vfnmadd231sd), the L factor one division.
inline void factor_ILU0(const MatrixCRS *A, MatrixCRS *L_strict, double *L_D, MatrixCRS *U_strict,
                        double *U_D) {
    const int n = A->n_rows;
    std::vector<double> w(n, 0.0);
    std::vector<int> idx;
    int nl = 0, nu = 0;
    L_strict->row_ptr[0] = U_strict->row_ptr[0] = 0;
    for (int i = 0; i < n; ++i) {
        idx.clear();
        for (int k = A->row_ptr[i]; k < A->row_ptr[i + 1]; ++k) {
            w[A->col[k]] = A->val[k];
            idx.push_back(A->col[k]);
        }
        std::sort(idx.begin(), idx.end());
        for (int k : idx) {
            if (k >= i) break;
            const double pivot = U_D[k];
            if (std::abs(pivot) < 1e-16) continue;   // unusable pivot: skip this elimination
            const double factor = w[k] / pivot;
            w[k] = factor;
            for (int t = U_strict->row_ptr[k]; t < U_strict->row_ptr[k + 1]; ++t) {
                const int j = U_strict->col[t];
                if (w[j] != 0.0) w[j] = std::fma(-factor, U_strict->val[t], w[j]);   // only on A's pattern
            }
        }
        double u_diag = 0.0;
        for (int j : idx) {
            if (j < i) {
                L_strict->col[nl] = j;
                L_strict->val[nl++] = w[j];
            } else if (j == i) {
                u_diag = w[j];
            } else {
                U_strict->col[nu] = j;
                U_strict->val[nu++] = w[j];
            }
        }
        if (std::abs(u_diag) < ILU0_PIVOT_TOLERANCE)
            u_diag = (u_diag >= 0 ? 1.0 : -1.0) * ILU0_PIVOT_REPLACEMENT;
        U_D[i] = u_diag;
        L_D[i] = 1.0;
        L_strict->row_ptr[i + 1] = nl;
        U_strict->row_ptr[i + 1] = nu;
        for (int j : idx) w[j] = 0.0;
    }
}

// factor_LU (LU_factors.hpp:900-934) on host arrays.
inline void factor_LU(const MatrixCRS *A, double *A_D, double *A_D_inv, MatrixCRS *L_strict,
                      double *L_D, MatrixCRS *U_strict, double *U_D, PrecondType preconditioner) {
    split_LU(A, L_strict, U_strict, A_D, A_D_inv);
    if (preconditioner == PrecondType::ILU0) factor_ILU0(A, L_strict, L_D, U_strict, U_D);
}
