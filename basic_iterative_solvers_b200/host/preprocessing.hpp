// host/preprocessing.hpp -- preprocessing() of the reference
// (preprocessing.hpp:26-100): allocate/init the solver vectors, hand A over,
// factor on the host, then -- new in the build -- upload A, L_strict, U_strict
// and the diagonals ONCE; they stay device-resident for the whole solve.
//
// Deviations, both stated in DESIGN.md: (1) the triangular copies are only
// built when the method or preconditioner uses them (the reference always
// builds all four, preprocessing.hpp:70-81; at HPCG-512 they would not fit);
// (2) -scale (preprocessing.hpp:39-50) is outside the hot-path scope and is
// rejected.
#pragma once

#include "common.hpp"
#include "lu_factors.hpp"
#include "solver.hpp"

// A: host CRS (consumed), or nullptr when dA was generated on the device.
// b_host / x0_host: optional right-hand side and initial guess (host arrays of
// the local length) replacing the B_VAL / INIT_X_VAL fills, set where the
// reference's tests set them: before init_structs (tests/test_solvers.cpp:79-84).
inline void preprocessing(Args *cli_args, Solver *solver, Timers *timers,
                          std::unique_ptr<MatrixCRS> &A, std::unique_ptr<DeviceCRS> dA = nullptr,
                          const double *b_host = nullptr, const double *x0_host = nullptr) {
    if (cli_args->num_scale) bis_fatal("-scale 1 is not supported by the device path (out of hot-path scope)");
    Interface *dev = solver->dev;

    timers->preprocessing_upload_time.start();
    if (A) {
        solver->dA = upload_crs(dev, A.get());
    } else {
        if (!dA) bis_fatal("preprocessing: no matrix");
        solver->dA = std::move(dA);
    }
    timers->preprocessing_upload_time.stop();
    const int64_t n = solver->dA->n_rows;
    solver->N_global = solver->dA->n_rows_global;

    timers->preprocessing_init_time.start();
    solver->allocate_structs(n);
    if (b_host) BIS_OK(bis_vector_upload(dev, solver->b, b_host, n));
    if (x0_host) BIS_OK(bis_vector_upload(dev, solver->x_0, x0_host, n));
    solver->init_structs(n);
    timers->preprocessing_init_time.stop();
    solver->A = std::move(A);

    timers->preprocessing_factor_time.start();
    if (solver->needs_triangular_factors()) {
        if (!solver->A) bis_fatal("this method/preconditioner needs L/U factors: the matrix must be host-resident");
        MatrixCRS L_strict, U_strict;
        const MatrixCRS *Ah = solver->A.get();
        std::vector<double> A_D(n, 1.0), A_D_inv(n, 0.0), L_D(n, 1.0), U_D(n, 1.0);
        factor_LU(Ah, A_D.data(), A_D_inv.data(), &L_strict, L_D.data(), &U_strict, U_D.data(),
                  solver->preconditioner);
        timers->preprocessing_upload_time.start();
        solver->dL_strict = upload_triangular(dev, &L_strict, false);
        solver->dU_strict = upload_triangular(dev, &U_strict, true);
        BIS_OK(bis_vector_upload(dev, solver->A_D, A_D.data(), n));
        BIS_OK(bis_vector_upload(dev, solver->A_D_inv, A_D_inv.data(), n));
        BIS_OK(bis_vector_upload(dev, solver->L_D, L_D.data(), n));
        BIS_OK(bis_vector_upload(dev, solver->U_D, U_D.data(), n));
        timers->preprocessing_upload_time.stop();
    } else {
        // peel_diag_crs on the device: A_D and 1/A_D (LU_factors.hpp:827-869)
        BIS_OK(bis_matrix_extract_diagonal(dev, solver->dA->handle, solver->A_D, solver->A_D_inv));
    }
    solver->A.reset();   // the host copy is not needed during the solve
    timers->preprocessing_factor_time.stop();

    timers->preprocessing_init_time.start();
    solver->init_residual();
    solver->init_stopping_criteria();
    timers->preprocessing_init_time.stop();
}
