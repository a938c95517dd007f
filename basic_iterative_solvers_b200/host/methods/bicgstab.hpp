// host/methods/bicgstab.hpp -- BiCGSTABSolver (reference methods/bicgstab.hpp).
// One iteration (bicgstab.hpp:8-83) as five device launches for the None /
// Jacobi preconditioners (two more per apply_preconditioner otherwise):
//   v = A y ; (r0,v)                                   bis_spmv_dot    :30,:34
//   alpha ; s = r_old - alpha v ; s_tmp = M^-1 s       bis_bicgstab_s  :34-45
//   z = A s_tmp ; (z,s) ; (z,z)                        bis_spmv_dot    :48,:51
//   omega ; h ; x_new ; r_new ; (r0,r_new) ; (r,r)     bis_bicgstab_xr :51-68
//   beta ; tmp ; p_new ; y_next = M^-1 p_new           bis_bicgstab_p  :70-78, next :24-27
// h and tmp are dead temporaries in the reference (written, never read again)
// and are not materialised.  `t`, `t_tmp` are unused in the reference and not
// allocated.  The residual_0 / residual / residual_old rotation of
// bicgstab.hpp:146-185 is kept as is.
#pragma once

#include "../solver.hpp"

class BiCGSTABSolver : public Solver {
  public:
    double *x_new = nullptr, *x_old = nullptr, *p_old = nullptr, *p_new = nullptr;
    double *v = nullptr, *s = nullptr, *s_tmp = nullptr, *z = nullptr, *y = nullptr;
    double *residual_old = nullptr, *residual_new = nullptr;
    double rho_old = 0.0, rho_new = 0.0;   // host mirrors (diagnostics only)

    int s_rho_old = S_RHO_OLD, s_rho_new = S_RHO_NEW;   // device scalar slots, swapped in exchange()

    BiCGSTABSolver(const Args *cli_args, Interface *device) : Solver(cli_args, device) {}

    void allocate_structs(const int64_t n) override {
        Solver::allocate_structs(n);
        for (double **p : {&x_new, &x_old, &p_new, &p_old, &residual_new, &residual_old, &v, &s, &s_tmp, &y, &z})
            *p = dev_new(dev, n);
    }
    void init_structs(const int64_t n) override {
        Solver::init_structs(n);
        for (double *p : {x_new, p_new, p_old, residual_new, residual_old, v, s, s_tmp})
            init_vector(dev, p, 0.0, n);
        copy_vector(dev, x_old, x_0, n);
    }
    // bicgstab.hpp:146-169
    void init_residual() override {
        BIS_OK(bis_spmv_residual(dev, dA->handle, x_old, b, residual, tmp, S_RR));
        copy_vector(dev, residual_old, residual, N);
        residual_norm = std::sqrt(scalar(dev, S_RR));
        precondition(residual, residual);   // in place: residual_0 is the PRECONDITIONED r0
        copy_vector(dev, p_old, residual, N);
        BIS_OK(bis_dot_to_slot(dev, residual_old, residual, N, s_rho_old));
        Solver::init_residual();
        if (fused_precond()) precondition(y, p_old);   // y of the first iteration (:24-27)
    }
    void iterate(Timers *) override {
        // residual / residual_new / residual_old rotate with period 3 (the swap below and the one in
        // exchange(), bicgstab.hpp:177,183), everything else with period 2: the arguments repeat every 6
        graphed(exchange_count % 6, [&] { enqueue_iterate(); });
        std::swap(residual, residual_new);   // bicgstab.hpp:177
    }
    void enqueue_iterate() {
        const int pc = static_cast<int>(preconditioner);
        if (!fused_precond()) precondition(y, p_old);
        BIS_OK(bis_spmv_dot(dev, dA->handle, y, v, residual_0, S_R0V, -1));
        BIS_OK(bis_bicgstab_s(dev, pc, N, s, s_tmp, residual_old, v, A_D, s_rho_old, S_R0V));
        if (!fused_precond()) precondition(s_tmp, s);
        BIS_OK(bis_spmv_dot(dev, dA->handle, s_tmp, z, s, S_ZS, S_ZZ));
        BIS_OK(bis_bicgstab_xr(dev, N, nullptr, x_new, x_old, y, s_tmp, residual_new, s, z, residual_0,
                               s_rho_old, S_R0V, S_ZS, S_ZZ, s_rho_new, S_RR_BI));
        BIS_OK(bis_bicgstab_p(dev, pc, N, nullptr, p_new, p_old, v, residual_new,
                              fused_precond() ? y : nullptr, A_D, s_rho_new, s_rho_old, S_R0V, S_ZS, S_ZZ));
    }
    void exchange() override {
        ++exchange_count;
        std::swap(p_old, p_new);
        std::swap(residual_old, residual);   // bicgstab.hpp:183
        std::swap(x_old, x_new);
        std::swap(s_rho_old, s_rho_new);   // std::swap(rho_old, rho_new), bicgstab.hpp:185: the slots trade places
    }
    void save_x_star() override {
        std::swap(x_old, x_star);
        Solver::save_x_star();
    }
    bool can_run_ahead() const override { return true; }
    int norm_slot() const override { return S_RR_BI; }
    // bicgstab.hpp:220-223: ||residual||_2 == (r_new,r_new) reduced in bis_bicgstab_xr
    void record_residual_norm() override {
        residual_norm = std::sqrt(scalar(dev, S_RR_BI));
        Solver::record_residual_norm();
    }
    ~BiCGSTABSolver() override {
        for (double **p : {&x_new, &x_old, &p_new, &p_old, &residual_new, &residual_old, &v, &s, &s_tmp, &y, &z})
            dev_delete(dev, *p);
    }
};
