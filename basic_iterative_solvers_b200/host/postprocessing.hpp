// host/postprocessing.hpp -- the parity read-out: residual history and summary
// in the reference's stdout format (postprocessing.hpp:8-77).
#pragma once

#include "common.hpp"
#include "solver.hpp"

inline void print_residuals(const double *norms, const double *time_per_iteration, int count,
                            int res_check_len) {
    std::cout << std::scientific << std::setprecision(16);
    std::cout << std::endl
              << std::string(15, ' ') << "Residual Norms" << std::string(27, ' ') << "Time for iteration"
              << std::endl;
    std::cout << "+------------------------------------------+" << std::string(8, ' ')
              << "+-------------------------+" << std::endl;
    for (int i = 0; i < count; ++i) {
        std::cout << "||A*x_" << i * res_check_len << " - b||_2 = " << norms[i];
        if (i > 0) std::cout << std::right << std::setw(30) << time_per_iteration[i + 1] << "[s]";
        std::cout << std::endl;
    }
}

inline void summary_output(Args *, Solver *solver) {
    print_residuals(solver->collected_residual_norms, solver->time_per_iteration,
                    solver->collected_residual_norms_count, solver->residual_check_len);
    if (solver->method == SolverType::GMRES) solver->iter_count += solver->gmres_restart_count;
    std::cout << "\nSolver: " << to_string(solver->method);
    if (solver->method == SolverType::GMRES) std::cout << "(" << solver->gmres_restart_len << ")";
    if (solver->preconditioner != PrecondType::None)
        std::cout << " with preconditioner: " << to_string(solver->preconditioner);
    if (solver->convergence_flag)
        std::cout << " converged in: " << solver->iter_count << " iterations." << std::endl;
    else
        std::cout << " did not converge after " << solver->iter_count << " iterations." << std::endl;
    std::cout << "With the stopping criteria \"tol * ||Ax_0 - b||_2\" is: " << solver->stopping_criteria
              << std::endl;
    std::cout << "The residual of the final iteration is: ||A*x_star - b||_2 = " << std::scientific
              << solver->collected_residual_norms[solver->collected_residual_norms_count - 1] << ".\n";
}

inline void print_timers(Args *, Timers *t) {
    auto row = [](const char *name, double s) {
        std::cout << std::left << std::setw(34) << name << std::right << std::fixed << std::setprecision(6)
                  << std::setw(14) << s << " [s]" << std::endl;
    };
    std::cout << "\n+---------------- timers (host wall clock) ----------------+" << std::endl;
    row("Total", t->total_time.get_wtime());
    row("| Preprocessing", t->preprocessing_time.get_wtime());
    row("| | init", t->preprocessing_init_time.get_wtime());
    row("| | factor (host)", t->preprocessing_factor_time.get_wtime());
    row("| | upload (host->device)", t->preprocessing_upload_time.get_wtime());
    row("| Solve", t->solve_time.get_wtime());
    row("| | iterate (enqueue)", t->iterate_time.get_wtime());
    row("| | sample (waits for device)", t->sample_time.get_wtime());
    row("| | restart", t->restart_time.get_wtime());
    row("| | save x*", t->save_x_star_time.get_wtime());
    std::cout << "+-----------------------------------------------------------+" << std::endl;
}

inline void postprocessing(Args *cli_args, Solver *solver, Timers *) { summary_output(cli_args, solver); }
