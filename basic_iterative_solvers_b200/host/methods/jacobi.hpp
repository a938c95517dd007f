// host/methods/jacobi.hpp -- JacobiSolver (reference methods/jacobi.hpp:54-122)
// with the sweep running on the device: SpMV and normalize_x
// (jacobi.hpp:27-52) are ONE kernel (bis_spmv_jacobi), the per-iteration
// residual sampling (jacobi.hpp:102-107) is one fused SpMV + subtract + norm.
#pragma once

#include "../solver.hpp"

inline void jacobi_separate_iteration(Interface *dev, const DeviceCRS *A, const double *D,
                                      const double *b, double *x_new, const double *x_old) {
    BIS_OK(bis_spmv_jacobi(dev, A->handle, D, b, x_old, x_new));
}

class JacobiSolver : public Solver {
  public:
    double *x_new = nullptr;
    double *x_old = nullptr;

    JacobiSolver(const Args *cli_args, Interface *device) : Solver(cli_args, device) {}

    void allocate_structs(const int64_t n) override {
        Solver::allocate_structs(n);
        x_new = dev_new(dev, n);
        x_old = dev_new(dev, n);
    }
    void init_structs(const int64_t n) override {
        Solver::init_structs(n);
        init_vector(dev, x_new, 0.0, n);
        copy_vector(dev, x_old, x_0, n);
    }
    void init_residual() override {
        BIS_OK(bis_spmv_residual(dev, dA->handle, x_old, b, residual, tmp, S_RR));
        residual_norm = std::sqrt(scalar(dev, S_RR));
        Solver::init_residual();
    }
    void iterate(Timers *) override {
        graphed(exchange_count & 1, [&] { jacobi_separate_iteration(dev, dA.get(), A_D, b, x_new, x_old); });
    }
    void exchange() override {
        ++exchange_count;
        std::swap(x_old, x_new);
    }
    void save_x_star() override {
        std::swap(x_old, x_star);
        Solver::save_x_star();
    }
    bool can_run_ahead() const override { return true; }
    void enqueue_residual_norm() override {
        BIS_OK(bis_spmv_residual(dev, dA->handle, x_new, b, residual, tmp, S_RR));
    }
    void record_residual_norm() override {
        BIS_OK(bis_spmv_residual(dev, dA->handle, x_new, b, residual, tmp, S_RR));
        residual_norm = std::sqrt(scalar(dev, S_RR));
        Solver::record_residual_norm();
    }
    ~JacobiSolver() override {
        dev_delete(dev, x_new);
        dev_delete(dev, x_old);
    }
};
