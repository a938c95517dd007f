// host/run.hpp -- run(): build the Solver for the chosen method, obtain A,
// preprocess, solve, postprocess (reference main.cpp:17-70).  Shared by the CLI
// (main.cpp) and the C entry point used by the tests (host_capi.cpp).
#pragma once

#include "methods/bicgstab.hpp"
#include "methods/cg.hpp"
#include "methods/gauss_seidel.hpp"
#include "methods/gmres.hpp"
#include "methods/jacobi.hpp"
#include "postprocessing.hpp"
#include "preprocessing.hpp"
#include "solver_harness.hpp"
#include "utilities.hpp"

inline std::unique_ptr<Solver> make_solver(const Args *cli_args, Interface *dev) {
    switch (cli_args->method) {
    case SolverType::Jacobi: return std::make_unique<JacobiSolver>(cli_args, dev);
    case SolverType::GaussSeidel: return std::make_unique<GaussSeidelSolver>(cli_args, dev);
    case SolverType::SymmetricGaussSeidel: return std::make_unique<SymmetricGaussSeidelSolver>(cli_args, dev);
    case SolverType::ConjugateGradient: return std::make_unique<ConjugateGradientSolver>(cli_args, dev);
    case SolverType::GMRES: return std::make_unique<GMRESSolver>(cli_args, dev);
    case SolverType::BiCGSTAB: return std::make_unique<BiCGSTABSolver>(cli_args, dev);
    }
    bis_fatal("Error: Unknown or unsupported solver type.");
}

// Matrix by name: a MatrixMarket file (read and converted on the host, uploaded by preprocessing) or
// a built-in generator, which runs on the device.  Splitting, ILU(0) and the level analysis happen on
// the device too (csrc/bis_factor.cu), so no method needs a host copy of a generated matrix.
// on_host: build a generated matrix in host memory instead (the CPU-only tests and the reference leg
// of the benchmarks feed the same CRS to the oracle).
inline void obtain_matrix(const Args *cli_args, Interface *dev, std::unique_ptr<MatrixCRS> &A,
                          std::unique_ptr<DeviceCRS> &dA, bool on_host = false) {
    const MatrixSpec spec = parse_matrix_spec(cli_args->matrix_file_name);
    if (spec.kind == MatrixSpec::File) {
        MatrixCOO coo;
        if (on_host || !dev) {
            coo.read_from_mtx(spec.path);
            A = std::make_unique<MatrixCRS>();
            convert_coo_to_crs(&coo, A.get());
            return;
        }
        // the entries go to the device in file order; the stable sort by row and the CRS conversion happen there
        coo.read_from_mtx(spec.path, /*sort_by_row=*/false);
        dA = std::make_unique<DeviceCRS>();
        dA->dev = dev;
        BIS_OK(bis_matrix_upload_coo(dev, coo.n_rows, coo.n_cols, coo.nnz, coo.I.data(), coo.J.data(), coo.values.data(), 0,
                                     &dA->handle));
        dA->refresh_info();
        return;
    }
    if (on_host) {
        A = spec.kind == MatrixSpec::Hpcg ? generate_hpcg(spec.nx, spec.ny, spec.nz)
                                          : generate_anderson(spec.nx, spec.ny, spec.nz, spec.ranpot, spec.t,
                                                              spec.seed, spec.periodic);
        return;
    }
    dA = std::make_unique<DeviceCRS>();
    dA->dev = dev;
    if (spec.kind == MatrixSpec::Hpcg)
        BIS_OK(bis_matrix_generate_hpcg(dev, spec.nx, spec.ny, spec.nz, &dA->handle));
    else
        BIS_OK(bis_matrix_generate_anderson(dev, spec.nx, spec.ny, spec.nz, spec.ranpot, spec.t, spec.seed,
                                            spec.periodic ? 1 : 0, &dA->handle));
    dA->refresh_info();
}

inline void run(Args *cli_args, Timers *timers, Interface *dev) {
    std::unique_ptr<Solver> solver = make_solver(cli_args, dev);
    std::unique_ptr<MatrixCRS> A;
    std::unique_ptr<DeviceCRS> dA;
    obtain_matrix(cli_args, dev, A, dA);
    TIME(timers->preprocessing, preprocessing(cli_args, solver.get(), timers, A, std::move(dA)))
    TIME(timers->solve, solve(cli_args, solver.get(), timers))
    TIME(timers->postprocessing, postprocessing(cli_args, solver.get(), timers))
}
