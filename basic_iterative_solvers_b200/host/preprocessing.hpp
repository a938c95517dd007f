// host/preprocessing.hpp -- preprocessing() of the reference
// (preprocessing.hpp:26-100): allocate/init the solver vectors, hand A over,
// upload A ONCE (or adopt a device-generated A), then -- new in the build -- split / ILU(0)-factor
// and level-analyse it ON THE DEVICE (csrc/bis_factor.cu); everything stays device-resident for the
// whole solve.
//
// Deviations, both stated in DESIGN.md: (1) the triangular copies are only
// built when the method or preconditioner uses them (the reference always
// builds all four, preprocessing.hpp:70-81; at HPCG-512 they would not fit);
// (2) the factor step works on the device-resident matrix.
#pragma once

#include "common.hpp"
#include "solver.hpp"

// A: host CRS (consumed), or nullptr when dA was generated on the device.
// b_host / x0_host: optional right-hand side and initial guess (host arrays of
// the local length) replacing the B_VAL / INIT_X_VAL fills, set where the
// reference's tests set them: before init_structs (tests/test_solvers.cpp:79-84).
inline void preprocessing(Args *cli_args, Solver *solver, Timers *timers,
                          std::unique_ptr<MatrixCRS> &A, std::unique_ptr<DeviceCRS> dA = nullptr,
                          const double *b_host = nullptr, const double *x0_host = nullptr) {
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

    if (solver->num_scale) {
        // preprocessing.hpp:39-50: solve A'x' = b' with A' = D^-1/2 A D^-1/2, b' = D^-1/2 b, x_0' = D^-1/2 x_0
        BIS_OK(bis_matrix_scale_symmetric(dev, solver->dA->handle, solver->A_D_scale));
        BIS_OK(bis_elemwise_mult_vectors(dev, solver->x_0, solver->A_D_scale, solver->x_0, n, 1.0));
        BIS_OK(bis_elemwise_mult_vectors(dev, solver->b, solver->A_D_scale, solver->b, n, 1.0));
    }

    // Permutation seam (preprocessing.hpp:52-65, permute_mat of utilities/smax_helpers.hpp:44-80): with the context
    // option "perm_mode" = 1 (the reference's PERM_MODE = C) A, b and x_0 are permuted by a multicolouring of A's
    // graph (2: by BFS levels, 3: reverse Cuthill-McKee, 4: Cuthill-McKee) before factoring -- a LABELLED mode: iteration counts differ from the unpermuted solve, and x_star stays
    // in the permuted numbering as in the reference.  Like there, the solver's own copy of x_0 (init_structs, above)
    // is not touched.
    int perm_mode = 0;
    if (bis_context_get_option(dev, "perm_mode", &perm_mode) == 0 && perm_mode != 0) {
        int *d_perm = nullptr, *d_inv = nullptr;
        BIS_OK(bis_index_alloc(dev, n, &d_perm));
        BIS_OK(bis_index_alloc(dev, n, &d_inv));
        if (perm_mode == 1) BIS_OK(bis_matrix_colouring_permutation(dev, solver->dA->handle, d_perm, d_inv, &solver->n_colours));
        else BIS_OK(bis_matrix_bfs_permutation(dev, solver->dA->handle, perm_mode, d_perm, d_inv, &solver->n_colours));   // 2 BFS, 3 RCM, 4 CM
        bis_matrix *pa = nullptr;
        BIS_OK(bis_matrix_permute_symmetric(dev, solver->dA->handle, d_perm, d_inv, &pa));
        solver->dA = adopt_device_matrix(dev, pa);
        BIS_OK(bis_vector_permute(dev, solver->tmp, solver->x_0, d_perm, n));
        copy_vector(dev, solver->x_0, solver->tmp, n);
        BIS_OK(bis_vector_permute(dev, solver->tmp, solver->b, d_perm, n));
        copy_vector(dev, solver->b, solver->tmp, n);
        init_vector(dev, solver->tmp, 0.0, n);
        BIS_OK(bis_context_synchronize(dev));
        bis_index_free(dev, d_perm);
        bis_index_free(dev, d_inv);
    }

    timers->preprocessing_factor_time.start();
    // peel_diag_crs on the device: A_D and 1/A_D (LU_factors.hpp:827-869); fatal on a missing or zero diagonal
    BIS_OK(bis_matrix_extract_diagonal(dev, solver->dA->handle, solver->A_D, solver->A_D_inv));
    if (solver->needs_triangular_factors()) {
        // factor_LU (LU_factors.hpp:900-934) on the device-resident matrix: strict split, or ILU(0)
        // (factor_ILU0_old) when it is the preconditioner; level analysis included.  L_D = U_D = 1 unless ILU(0).
        bis_matrix *l = nullptr, *u = nullptr;
        // A Krylov method that uses the factors only inside gs / bgs / sgs / ilu0 preconditioner solves never multiplies
        // by them: they may give up their natural-order CRS once the level-ordered copy exists (HPCG-512 -p sgs then
        // fits one GPU).  Gauss-Seidel SWEEPS (b - T x) and the two-stage preconditioners keep it.
        const bool krylov = solver->method == SolverType::ConjugateGradient || solver->method == SolverType::GMRES ||
                            solver->method == SolverType::BiCGSTAB;
        const bool solves_only = solver->preconditioner == PrecondType::GaussSeidel ||
                                 solver->preconditioner == PrecondType::BackwardsGaussSeidel ||
                                 solver->preconditioner == PrecondType::SymmetricGaussSeidel ||
                                 solver->preconditioner == PrecondType::ILU0;
        int keep_before = 1;
        BIS_OK(bis_context_get_option(dev, "factor_keep_crs", &keep_before));
        struct RestoreOption {      // also when factoring throws: the context outlives a failed solve (Python host API)
            Interface *dev;
            int value;
            ~RestoreOption() { bis_context_set_option(dev, "factor_keep_crs", value); }
        } restore_keep{dev, keep_before};
        if (krylov && solves_only) BIS_OK(bis_context_set_option(dev, "factor_keep_crs", 0));
        if (solver->preconditioner == PrecondType::ILU0)
            BIS_OK(bis_matrix_ilu0(dev, solver->dA->handle, ILU0_PIVOT_TOLERANCE, ILU0_PIVOT_REPLACEMENT, &l, &u,
                                   solver->L_D, solver->U_D));
        else
            BIS_OK(bis_matrix_split_triangular(dev, solver->dA->handle, &l, &u));
        solver->dL_strict = adopt_device_matrix(dev, l);
        solver->dU_strict = adopt_device_matrix(dev, u);
    }
    solver->A.reset();   // the host copy is not needed during the solve
    timers->preprocessing_factor_time.stop();

    timers->preprocessing_init_time.start();
    solver->init_residual();
    solver->init_stopping_criteria();
    timers->preprocessing_init_time.stop();
}
