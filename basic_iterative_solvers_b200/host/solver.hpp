// host/solver.hpp -- the Solver base class of the reference (solver.hpp:9-193)
// with device-resident state: every `double *` vector member is a DEVICE
// address, the matrices are device CRS handles, and the residual history stays
// on the host.  Virtual interface, member names, stopping logic
// (solver.hpp:173-191) and history bookkeeping (solver.hpp:148-171) are the
// reference's.
#pragma once

#include <cstdlib>

#include "common.hpp"
#include "kernels.hpp"
#include "sparse_matrix.hpp"

#include <cfloat>

// device scalar slots shared by the methods (include/bis_b200.h "device scalars")
enum Slot : int {
    S_RR = 0,        // (r,r) of the newest residual
    S_RZ_NEW = 1,    // adjacent to S_RR: one allreduce of 2
    S_RZ = 2,
    S_PAP = 3,
    S_R0V = 4,
    S_ZS = 5,
    S_ZZ = 6,        // adjacent to S_ZS
    S_RHO_NEW = 7,
    S_RR_BI = 8,     // adjacent to S_RHO_NEW
    S_RHO_OLD = 9,
    S_BETA2 = 10,
    S_H = 16         // S_H + j, j = 0..k: GMRES Hessenberg column h_jk; S_H + k + 1: ||w||^2
};

class Solver {
  public:
    SolverType method;
    PrecondType preconditioner = PrecondType::None;
    Interface *dev = nullptr;   // the device context (the reference's `smax` slot)

    // host copy of A (only kept while preprocessing needs it) and device mirrors
    std::unique_ptr<MatrixCRS> A;
    std::unique_ptr<DeviceCRS> dA, dL_strict, dU_strict;
    int64_t N = 0;          // local rows
    int64_t N_global = 0;

    double stopping_criteria = 0.0;
    int iter_count = 0;
    int collected_residual_norms_count = 0;
    double residual_norm = DBL_MAX;
    int max_iters = MAX_ITERS;
    double tolerance = TOL;
    int residual_check_len = RES_CHECK_LEN;
    int gmres_restart_len = 0;
    int gmres_restart_count = 0;
    bool num_scale = false;

    // common vectors [dev]
    double *x_star = nullptr, *x_0 = nullptr, *b = nullptr, *tmp = nullptr, *work = nullptr;
    double *residual = nullptr, *residual_0 = nullptr;
    double *A_D = nullptr, *A_D_inv = nullptr, *L_D = nullptr, *U_D = nullptr, *A_D_scale = nullptr;

    // bookkeeping [host]
    double *collected_residual_norms = nullptr;
    double *time_per_iteration = nullptr;

    bool convergence_flag = false;
    bool gmres_restarted = false;

    Solver(const Args *cli_args, Interface *device)
        : method(cli_args->method), preconditioner(cli_args->preconditioner), dev(device),
          gmres_restart_len(cli_args->restart_length), num_scale(cli_args->num_scale) {
        collected_residual_norms = new double[max_iters * 2]();
        time_per_iteration = new double[max_iters * 2]();
        int v = 0;
        if (dev && bis_context_get_option(dev, "graph", &v) == 0) graphs_on = v != 0;
        if (dev && bis_context_get_option(dev, "precond_inner_iters", &v) == 0) precond_inner_iters = v;
        // RES_CHECK_LEN is a compile-time constant of the reference (CMakeLists.txt:232-243); BIS_RES_CHECK_LEN overrides it
        // at run time (measurements: how much of an iteration is the read-back of the residual norm)
        if (const char *e = std::getenv("BIS_RES_CHECK_LEN"))
            if (std::atoi(e) > 0) residual_check_len = std::atoi(e);
    }

    // ---- CUDA graphs (new in the build) ------------------------------------------------------------
    // The device part of an iteration is a fixed sequence of launches: its arguments repeat with the
    // period of the method's pointer exchange, its scalars (alpha, beta, omega, h_jk) live on the device.
    // graphed(key, enqueue): the first time a key comes up the calls are issued as usual (that also
    // builds whatever a kernel builds lazily), the second time they are recorded into a CUDA graph and the
    // graph is launched, from then on it is replayed -- one submission per iteration instead of 3..30.
    // `enqueue` must only enqueue (no host-side state change, no readback).  Values are bit-identical.
    struct GraphSlot {
        bis_graph *graph = nullptr;
        int seen = 0;
        bool failed = false;
    };
    std::vector<GraphSlot> graph_slots;
    bool graphs_on = false;

    template <class F> void graphed(const int key, F &&enqueue) {
        if (!graphs_on || !dev || key < 0) {
            enqueue();
            return;
        }
        if ((int)graph_slots.size() <= key) graph_slots.resize((size_t)key + 1);
        GraphSlot &slot = graph_slots[(size_t)key];
        if (slot.graph) {
            BIS_OK(bis_graph_launch(dev, slot.graph));
            return;
        }
        if (slot.failed || slot.seen++ < 1 || bis_graph_begin(dev) != 0) {
            enqueue();
            return;
        }
        enqueue();
        if (bis_graph_end(dev, &slot.graph) != 0 || !slot.graph) {
            slot.failed = true;     // e.g. a call that is not a pure enqueue: issue the sequence as usual
            slot.graph = nullptr;
            enqueue();
            return;
        }
        BIS_OK(bis_graph_launch(dev, slot.graph));
    }
    void drop_graphs() {
        for (GraphSlot &slot : graph_slots)
            if (slot.graph) bis_graph_free(dev, slot.graph);
        graph_slots.clear();
    }
    int n_colours = 0;        // colours (perm_mode 1) or BFS levels (perm_mode 2..4) of the permutation (0: unpermuted)
    int exchange_count = 0;   // pointer exchanges so far: the key of the iteration's graph

    virtual void iterate(Timers *) = 0;
    virtual void exchange() = 0;

    int precond_inner_iters = PRECOND_INNER_ITERS;   // the reference's -DPRECOND_INNER_ITERS; here the context option
    bool needs_triangular_factors() const {
        const bool two_stage = preconditioner == PrecondType::TwoStageGS || preconditioner == PrecondType::SymmetricTwoStageGS;
        return (two_stage && precond_inner_iters > 0) ||   // their inner sweeps multiply by the strict factors
               method == SolverType::GaussSeidel || method == SolverType::SymmetricGaussSeidel ||
               preconditioner == PrecondType::GaussSeidel ||
               preconditioner == PrecondType::BackwardsGaussSeidel ||
               preconditioner == PrecondType::SymmetricGaussSeidel ||
               preconditioner == PrecondType::ILU0;
    }

    // solver.hpp:82-110: x_star = 0, x_0 = INIT_X_VAL, b = B_VAL, diagonals = 1
    virtual void allocate_structs(const int64_t n) {
        N = n;
        x_star = dev_new(dev, n);
        x_0 = dev_new(dev, n);
        b = dev_new(dev, n);
        tmp = dev_new(dev, n);
        work = dev_new(dev, n);
        residual = dev_new(dev, n);
        residual_0 = dev_new(dev, n);
        A_D = dev_new(dev, n);
        A_D_inv = dev_new(dev, n);
        L_D = dev_new(dev, n);
        U_D = dev_new(dev, n);
        A_D_scale = dev_new(dev, n);
        if (!gmres_restarted) {
            init_vector(dev, x_star, 0.0, n);
            init_vector(dev, x_0, INIT_X_VAL, n);
            init_vector(dev, b, B_VAL, n);
            init_vector(dev, A_D, 1.0, n);
            init_vector(dev, A_D_inv, 0.0, n);
            init_vector(dev, L_D, 1.0, n);
            init_vector(dev, U_D, 1.0, n);
            init_vector(dev, A_D_scale, 0.0, n);
        }
    }

    // solver.hpp:112-120
    virtual void init_structs(const int64_t n) {
        init_vector(dev, tmp, 0.0, n);
        init_vector(dev, work, 0.0, n);
        init_vector(dev, residual, 0.0, n);
        init_vector(dev, residual_0, 0.0, n);
    }

    virtual void check_restart(Timers *) {}
    virtual void get_explicit_x() {}

    virtual ~Solver() {
        drop_graphs();
        for (double **p : {&x_star, &x_0, &b, &tmp, &work, &residual, &residual_0, &A_D, &A_D_inv, &L_D, &U_D, &A_D_scale})
            dev_delete(dev, *p);
        delete[] collected_residual_norms;
        delete[] time_per_iteration;
    }

    // solver.hpp:148-151
    virtual void init_residual() {
        copy_vector(dev, residual_0, residual, N);
        collected_residual_norms[collected_residual_norms_count++] = residual_norm;
    }

    // solver.hpp:153-159: true residual of x_star, parked at [count + 1]
    virtual void save_x_star() {
        BIS_OK(bis_spmv_residual(dev, dA->handle, x_star, b, residual, tmp, S_RR));
        residual_norm = std::sqrt(scalar(dev, S_RR));
        if (collected_residual_norms_count + 1 < 2 * max_iters)
            collected_residual_norms[collected_residual_norms_count + 1] = residual_norm;
    }

    virtual void record_residual_norm() {
        collected_residual_norms[collected_residual_norms_count++] = residual_norm;
    }

    // ---- run-ahead (new in the build) ------------------------------------------------------------
    // Methods whose iterate() never touches the vector that would become x_star (double-buffered
    // x_new / x_old) let the harness enqueue iteration k+1 BEHIND the residual-norm readback of
    // iteration k and only then wait for the value: the device works while the host decides.  When
    // the decision is "stop", the iteration that ran ahead has only written the buffers a further
    // iteration would have overwritten anyway.  norm_slot(): the device scalar holding ||r||^2.
    virtual bool can_run_ahead() const { return false; }
    virtual int norm_slot() const { return S_RR; }
    virtual void enqueue_residual_norm() {}   // kernels that produce the norm, if iterate() did not

    // the split read of ||r||^2 (virtual so that the CPU tests can drive the harness without a device)
    virtual void read_norm_begin() { BIS_OK(bis_scalar_read_begin(dev, norm_slot(), 1)); }
    virtual double read_norm_end() {
        double rr = 0.0;
        BIS_OK(bis_scalar_read_end(dev, norm_slot(), 1, &rr));
        return rr;
    }

    void sample_residual_begin() {
        if (iter_count % residual_check_len == 0) {
            enqueue_residual_norm();
            read_norm_begin();
        }
    }
    void sample_residual_end(Stopwatch *per_iteration_time) {
        if (iter_count % residual_check_len == 0) {
            const double rr = read_norm_end();
            residual_norm = std::sqrt(rr);
            Solver::record_residual_norm();
            time_per_iteration[collected_residual_norms_count] = per_iteration_time->check();
        }
    }

    void sample_residual(Stopwatch *per_iteration_time) {
        if (iter_count % residual_check_len == 0) {
            record_residual_norm();
            time_per_iteration[collected_residual_norms_count] = per_iteration_time->check();
        }
    }

    void init_stopping_criteria() { stopping_criteria = tolerance * residual_norm; }

    bool check_stopping_criteria() {
        bool norm_convergence = std::abs(residual_norm) < stopping_criteria;
        bool over_max_iters = iter_count >= (max_iters - gmres_restart_count);
        bool divergence = std::abs(residual_norm) > DBL_MAX || std::isnan(residual_norm);
        return norm_convergence || over_max_iters || divergence;
    }

  protected:
    void precondition(double *out, double *in) {
        apply_preconditioner(dev, preconditioner, N, dL_strict.get(), dU_strict.get(), A_D, A_D_inv,
                             L_D, U_D, out, in, tmp, work);
    }
    bool fused_precond() const {
        return preconditioner == PrecondType::None || preconditioner == PrecondType::Jacobi;
    }
};
