// host/methods/jacobi.hpp -- JacobiSolver (reference methods/jacobi.hpp:54-122)
// with the sweep running on the device: SpMV and normalize_x
// (jacobi.hpp:27-52) are ONE kernel, and the per-iteration residual sampling
// (jacobi.hpp:102-107) rides on the NEXT sweep's kernel: iteration k+1 starts
// from x_k and forms A x_k anyway, so ||b - A x_k|| comes out of the same
// product (bis_spmv_jacobi_residual) -- one SpMV per iteration instead of two.
// The harness enqueues iteration k+1 before it waits for the norm of
// iteration k (run-ahead), so that norm is simply read after that launch;
// only when no further iteration is enqueued (last pass) does a separate
// residual kernel run.  Same values, same decisions as the reference's order.
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
    // exchange_count at the time the last sweep was enqueued: that launch left ||b - A x_{exchange_count}||^2 in S_RR
    long long norm_of_exchange = -1;
    void iterate(Timers *) override {
        graphed(exchange_count & 1, [&] {
            BIS_OK(bis_spmv_jacobi_residual(dev, dA->handle, A_D, b, x_old, x_new, residual, S_RR));
        });
        norm_of_exchange = (long long)exchange_count;
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
    // run-ahead sampling: nothing to enqueue after the sweep -- the norm of the iterate just produced is formed by the
    // next sweep (or, if the harness enqueues none, by a residual kernel in read_norm_end)
    void enqueue_residual_norm() override {}
    void read_norm_begin() override {}
    double read_norm_end() override {
        // by now exchange() has made the new iterate x_old
        if (norm_of_exchange != (long long)exchange_count)
            BIS_OK(bis_spmv_residual(dev, dA->handle, x_old, b, residual, tmp, S_RR));
        double rr = 0.0;
        BIS_OK(bis_scalar_read_begin(dev, S_RR, 1));
        BIS_OK(bis_scalar_read_end(dev, S_RR, 1, &rr));
        return rr;
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
