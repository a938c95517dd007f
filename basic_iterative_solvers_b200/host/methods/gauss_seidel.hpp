// host/methods/gauss_seidel.hpp -- GaussSeidelSolver / SymmetricGaussSeidelSolver
// (reference methods/gauss_seidel.hpp:54-142).  A forward sweep is
// tmp = b - U_strict x (one fused kernel for :30-34) followed by the
// level-scheduled solve x = (D+L)^-1 tmp (:36); the backward sweep mirrors it
// (:40-52).  Row semantics and summation order are the reference's.
#pragma once

#include "../solver.hpp"

inline void gs_separate_iteration(Interface *dev, const DeviceCRS *U, const DeviceCRS *L, double *tmp,
                                  const double *D, const double *b, double *x) {
    BIS_OK(bis_spmv_sub(dev, U->handle, x, b, tmp));   // tmp <- b - U*x
    sptrsv(dev, L, x, D, tmp);                         // x <- (D+L)^{-1} tmp
}
inline void bgs_separate_iteration(Interface *dev, const DeviceCRS *U, const DeviceCRS *L, double *tmp,
                                   const double *D, const double *b, double *x) {
    BIS_OK(bis_spmv_sub(dev, L->handle, x, b, tmp));   // tmp <- b - L*x
    bsptrsv(dev, U, x, D, tmp);                        // x <- (D+U)^{-1} tmp
}

class GaussSeidelSolver : public Solver {
  public:
    double *x = nullptr;

    GaussSeidelSolver(const Args *cli_args, Interface *device) : Solver(cli_args, device) {}

    void allocate_structs(const int64_t n) override {
        Solver::allocate_structs(n);
        x = dev_new(dev, n);
    }
    void init_structs(const int64_t n) override {
        Solver::init_structs(n);
        copy_vector(dev, x, x_0, n);
    }
    void init_residual() override {
        BIS_OK(bis_spmv_residual(dev, dA->handle, x, b, residual, tmp, S_RR));
        residual_norm = std::sqrt(scalar(dev, S_RR));
        Solver::init_residual();
    }
    void iterate(Timers *) override {
        graphed(0, [&] { gs_separate_iteration(dev, dU_strict.get(), dL_strict.get(), tmp, A_D, b, x); });
    }
    void exchange() override {}
    void save_x_star() override {
        std::swap(x, x_star);
        Solver::save_x_star();
    }
    void record_residual_norm() override {
        BIS_OK(bis_spmv_residual(dev, dA->handle, x, b, residual, tmp, S_RR));
        residual_norm = std::sqrt(scalar(dev, S_RR));
        Solver::record_residual_norm();
    }
    ~GaussSeidelSolver() override { dev_delete(dev, x); }
};

class SymmetricGaussSeidelSolver : public GaussSeidelSolver {
  public:
    SymmetricGaussSeidelSolver(const Args *cli_args, Interface *device)
        : GaussSeidelSolver(cli_args, device) {}
    void iterate(Timers *) override {
        graphed(0, [&] {
            gs_separate_iteration(dev, dU_strict.get(), dL_strict.get(), tmp, A_D, b, x);
            bgs_separate_iteration(dev, dU_strict.get(), dL_strict.get(), tmp, A_D, b, x);
        });
    }
};
