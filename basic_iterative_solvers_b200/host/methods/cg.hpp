// host/methods/cg.hpp -- ConjugateGradientSolver (reference methods/cg.hpp).
// One iteration (cg.hpp:6-54) is enqueued without a host round trip:
//   tmp = A p_old ; (tmp,p_old)                         bis_spmv_dot      :16,:23
//   alpha ; r_new ; (r_new,r_new) ; z_new ;             bis_cg_update      :23-41,:47
//   (r_new,z_new)   [None / Jacobi fused; other M^-1 via apply_preconditioner]
//   beta ; p_new ; x_new = x_old + alpha p_old          bis_cg_direction_x :27-28, :47-52
// (10 vector passes per iteration with the Jacobi preconditioner instead of the 18 of the
// reference's separate kernels)
// alpha and beta are formed on the device from the scalar slots.  (r_old,z_old)
// (:19) is carried over from the previous iteration's (r_new,z_new): same
// numbers, same deterministic reduction, so the value is bit-identical to
// recomputing it.  The only host<->device sync is the residual norm the
// harness samples every iteration (solver_harness.hpp:24).
#pragma once

#include "../solver.hpp"

inline void cg_separate_iteration(Interface *dev, const PrecondType preconditioner, const int64_t N,
                                  const DeviceCRS *A, const DeviceCRS *L, const DeviceCRS *U,
                                  double *A_D, double *A_D_inv, double *L_D, double *U_D,
                                  double *x_new, double *x_old, double *tmp, double *work,
                                  double *p_new, double *p_old, double *r_new, double *r_old,
                                  double *z_new, double *z_old, const int s_rz, const int s_rz_new) {
    (void)z_old;
    BIS_OK(bis_spmv_dot(dev, A->handle, p_old, tmp, p_old, S_PAP, -1));
    // the x update (cg.hpp:27-28) rides with the direction update below: one pass over p_old less
    BIS_OK(bis_cg_update(dev, static_cast<int>(preconditioner), N, nullptr, x_old, p_old, r_new, r_old,
                         tmp, z_new, A_D, s_rz, S_PAP, S_RR, s_rz_new));
    if (preconditioner != PrecondType::None && preconditioner != PrecondType::Jacobi) {
        apply_preconditioner(dev, preconditioner, N, L, U, A_D, A_D_inv, L_D, U_D, z_new, r_new, tmp, work);
        BIS_OK(bis_dot_to_slot(dev, r_new, z_new, N, s_rz_new));
    }
    BIS_OK(bis_cg_direction_x(dev, N, p_new, z_new, p_old, x_new, x_old, s_rz_new, s_rz, S_PAP));
}

class ConjugateGradientSolver : public Solver {
  public:
    double *x_new = nullptr, *x_old = nullptr, *p_old = nullptr, *p_new = nullptr;
    double *z_old = nullptr, *z_new = nullptr, *residual_old = nullptr, *residual_new = nullptr;
    // device scalar slots of (r_old,z_old) and (r_new,z_new): swapped in exchange() like the vectors
    int s_rz = S_RZ, s_rz_new = S_RZ_NEW;

    ConjugateGradientSolver(const Args *cli_args, Interface *device) : Solver(cli_args, device) {}

    void allocate_structs(const int64_t n) override {
        Solver::allocate_structs(n);
        for (double **p : {&x_new, &x_old, &p_new, &p_old, &residual_new, &residual_old, &z_new, &z_old})
            *p = dev_new(dev, n);
    }
    void init_structs(const int64_t n) override {
        Solver::init_structs(n);
        for (double *p : {x_new, p_new, p_old, residual_new, residual_old, z_new, z_old})
            init_vector(dev, p, 0.0, n);
        copy_vector(dev, x_old, x_0, n);
    }
    // cg.hpp:100-120
    void init_residual() override {
        BIS_OK(bis_spmv_residual(dev, dA->handle, x_old, b, residual, tmp, S_RR));
        precondition(z_old, residual);
        copy_vector(dev, p_old, z_old, N);
        copy_vector(dev, residual_old, residual, N);
        BIS_OK(bis_dot_to_slot(dev, residual_old, z_old, N, s_rz));   // first (r_old, z_old)
        residual_norm = std::sqrt(scalar(dev, S_RR));
        Solver::init_residual();
    }
    void iterate(Timers *) override {
        // the vector pointers and the two (r,z) slots swap in exchange(): the arguments have period 2
        graphed(exchange_count & 1, [&] {
            cg_separate_iteration(dev, preconditioner, N, dA.get(), dL_strict.get(), dU_strict.get(), A_D,
                                  A_D_inv, L_D, U_D, x_new, x_old, tmp, work, p_new, p_old, residual_new,
                                  residual_old, z_new, z_old, s_rz, s_rz_new);
        });
    }
    void exchange() override {
        ++exchange_count;
        std::swap(p_old, p_new);
        std::swap(z_old, z_new);
        std::swap(residual_old, residual_new);
        std::swap(x_old, x_new);
        std::swap(s_rz, s_rz_new);   // (r_old,z_old) of the next iteration is this one's (r_new,z_new)
    }
    void save_x_star() override {
        std::swap(x_old, x_star);
        Solver::save_x_star();
    }
    bool can_run_ahead() const override { return true; }   // x_new / x_old are double-buffered
    // cg.hpp:162-166: ||residual_new||_2, already reduced by bis_cg_update
    void record_residual_norm() override {
        residual_norm = std::sqrt(scalar(dev, S_RR));
        Solver::record_residual_norm();
    }
    ~ConjugateGradientSolver() override {
        for (double **p : {&x_new, &x_old, &p_new, &p_old, &residual_new, &residual_old, &z_new, &z_old})
            dev_delete(dev, *p);
    }
};
