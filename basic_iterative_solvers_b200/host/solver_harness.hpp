// host/solver_harness.hpp -- solve(): the iteration loop of the reference
// (solver_harness.hpp:7-61), unchanged in structure.  iterate() only enqueues
// device work; sample_residual() is where the host waits for the residual norm
// (RES_CHECK_LEN = 1: once per iteration), so per_iteration_time is a true
// time per iteration.  For methods that allow it the next iteration is enqueued
// before the host waits for that norm (harness_step).
#pragma once

#include "common.hpp"
#include "solver.hpp"

// One pass of the reference's loop body (solver_harness.hpp:17-50).  `ahead`: iterate() of this pass
// has already been enqueued by the previous pass.  With run-ahead the order of the reference
// (iterate, sample, exchange, check_restart) becomes iterate, enqueue the norm readback, exchange,
// enqueue the NEXT iterate (if `may_run_ahead`), wait for the norm: same values, same decisions.
inline void harness_step(Solver *solver, Timers *timers, bool &ahead, bool may_run_ahead) {
    timers->per_iteration_time.start();
    if (!ahead) TIME(timers->iterate, solver->iterate(timers))
    ahead = false;
    ++solver->iter_count;
    if (solver->can_run_ahead()) {
        TIME(timers->sample, solver->sample_residual_begin())
        TIME(timers->exchange, solver->exchange())
        if (may_run_ahead && solver->iter_count < solver->max_iters - solver->gmres_restart_count) {
            TIME(timers->iterate, solver->iterate(timers))
            ahead = true;
        }
        TIME(timers->sample, solver->sample_residual_end(&timers->per_iteration_time))
    } else {
        TIME(timers->sample, solver->sample_residual(&timers->per_iteration_time))
        TIME(timers->exchange, solver->exchange())
    }
    TIME(timers->restart, solver->check_restart(timers))
}

inline void solve(Args *cli_args, Solver *solver, Timers *timers) {
    const double initial_res = solver->collected_residual_norms[0];
    bool res_3 = false, res_6 = false;
    bool ahead = false;
    do {
        harness_step(solver, timers, ahead, true);
        if (solver->residual_norm / initial_res < 1e-3 && !res_3) {
            if (!cli_args->quiet) std::cout << "res3 => iter_count: " << solver->iter_count << std::endl;
            res_3 = true;
        }
        if (solver->residual_norm / initial_res < 1e-6 && !res_6) {
            if (!cli_args->quiet) std::cout << "res6 => iter_count: " << solver->iter_count << std::endl;
            res_6 = true;
        }
    } while (!solver->check_stopping_criteria());

    if (solver->residual_norm < solver->stopping_criteria) solver->convergence_flag = true;
    TIME(timers->save_x_star, solver->save_x_star())
    // a watchdog trip inside a triangular solve surfaces here as a fatal error
    if (solver->dev) BIS_OK(bis_context_synchronize(solver->dev));
}
