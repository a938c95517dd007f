// host/solver_harness.hpp -- solve(): the iteration loop of the reference
// (solver_harness.hpp:7-61), unchanged in structure.  iterate() only enqueues
// device work; sample_residual() is where the host waits for the residual norm
// (RES_CHECK_LEN = 1: once per iteration), so per_iteration_time is a true
// time per iteration.
#pragma once

#include "common.hpp"
#include "solver.hpp"

inline void solve(Args *cli_args, Solver *solver, Timers *timers) {
    const double initial_res = solver->collected_residual_norms[0];
    bool res_3 = false, res_6 = false;
    do {
        timers->per_iteration_time.start();
        TIME(timers->iterate, solver->iterate(timers))
        ++solver->iter_count;
        TIME(timers->sample, solver->sample_residual(&timers->per_iteration_time))
        if (solver->residual_norm / initial_res < 1e-3 && !res_3) {
            if (!cli_args->quiet) std::cout << "res3 => iter_count: " << solver->iter_count << std::endl;
            res_3 = true;
        }
        if (solver->residual_norm / initial_res < 1e-6 && !res_6) {
            if (!cli_args->quiet) std::cout << "res6 => iter_count: " << solver->iter_count << std::endl;
            res_6 = true;
        }
        TIME(timers->exchange, solver->exchange())
        TIME(timers->restart, solver->check_restart(timers))
    } while (!solver->check_stopping_criteria());

    if (solver->residual_norm < solver->stopping_criteria) solver->convergence_flag = true;
    TIME(timers->save_x_star, solver->save_x_star())
    // a watchdog trip inside a triangular solve surfaces here as a fatal error
    BIS_OK(bis_context_synchronize(solver->dev));
}
