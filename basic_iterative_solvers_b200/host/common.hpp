// host/common.hpp -- enums, CLI arguments, timers and the fatal-error
// convention of the host side.  Mirrors the reference's common.hpp
// (PrecondType/SolverType :38-56, Args :105-111, Stopwatch/Timers :206-354)
// for the names the hot path touches.
#pragma once

#include "bis_b200.h"

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

// compile-time solver parameters (reference CMakeLists.txt:20-29, 232-243)
#ifndef MAX_ITERS
#define MAX_ITERS 1000
#endif
#ifndef TOL
#define TOL 1e-14
#endif
#ifndef RES_CHECK_LEN
#define RES_CHECK_LEN 1
#endif
#ifndef INIT_X_VAL
#define INIT_X_VAL 0.1
#endif
#ifndef B_VAL
#define B_VAL 1.0
#endif
#ifndef PRECOND_INNER_ITERS
#define PRECOND_INNER_ITERS 0   // inner sweeps of -p 2st / s2st (kernels.hpp:321); a context option overrides it
#endif
#ifndef ILU0_PIVOT_TOLERANCE
#define ILU0_PIVOT_TOLERANCE 1e-8
#endif
#ifndef ILU0_PIVOT_REPLACEMENT
#define ILU0_PIVOT_REPLACEMENT 1e-4
#endif

// values match the C-ABI's BIS_PRECOND_* and the reference's enum order
enum class PrecondType {
    None = BIS_PRECOND_NONE,
    Jacobi = BIS_PRECOND_JACOBI,
    GaussSeidel = BIS_PRECOND_GS,
    BackwardsGaussSeidel = BIS_PRECOND_BGS,
    SymmetricGaussSeidel = BIS_PRECOND_SGS,
    TwoStageGS = BIS_PRECOND_2ST,
    SymmetricTwoStageGS = BIS_PRECOND_S2ST,
    ILU0 = BIS_PRECOND_ILU0
};

enum class SolverType { Jacobi, GaussSeidel, SymmetricGaussSeidel, GMRES, ConjugateGradient, BiCGSTAB };

inline std::string to_string(PrecondType t) {
    static const char *names[] = {"none", "jacobi", "gauss-seidel", "backwards-gauss-seidel",
                                  "symmetric-gauss-seidel", "two-stage gauss-seidel",
                                  "symmetric two-stage gauss-seidel", "incomplete LU(0)"};
    int i = static_cast<int>(t);
    return (i >= 0 && i < 8) ? names[i] : "unknown";
}
inline std::string to_string(SolverType t) {
    static const char *names[] = {"jacobi", "gauss-seidel", "symmetric-gauss-seidel", "gmres",
                                  "conjugate-gradient", "bicgstab"};
    int i = static_cast<int>(t);
    return (i >= 0 && i < 6) ? names[i] : "unknown";
}

struct Args {
    std::string matrix_file_name{};
    SolverType method{};
    PrecondType preconditioner{};
    int restart_length = 10;
    bool num_scale = false;
    // additions of the build (not in the reference CLI)
    int device = 0;
    bool quiet = false;
};

// The reference prints to stderr and exit(EXIT_FAILURE)s on fatal errors
// (common.hpp:382-396, utilities.hpp:14-20).  The host library throws this
// instead; main.cpp turns it back into the same stderr + exit behaviour and the
// C entry point (host_capi.cpp) into an error code.
struct FatalError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
[[noreturn]] inline void bis_fatal(const std::string &msg) { throw FatalError(msg); }

// Wall-clock stopwatch.  Device work is asynchronous: only timers whose
// region ends in a host<->device synchronisation are meaningful
// (per_iteration, iterate+sample, solve, preprocessing, total).
class Stopwatch {
    using clk = std::chrono::steady_clock;
    clk::time_point t0{};
    double acc = 0.0;

  public:
    void start() { t0 = clk::now(); }
    void stop() { acc += std::chrono::duration<double>(clk::now() - t0).count(); }
    double check() const { return std::chrono::duration<double>(clk::now() - t0).count(); }
    double get_wtime() const { return acc; }
};

struct Timers {
    Stopwatch total_time, preprocessing_time, preprocessing_init_time, preprocessing_factor_time,
        preprocessing_upload_time, solve_time, per_iteration_time, iterate_time, sample_time,
        exchange_time, restart_time, save_x_star_time, postprocessing_time;
};

#define TIME(timer_name, routine)                                                                  \
    do {                                                                                           \
        timer_name##_time.start();                                                                 \
        routine;                                                                                   \
        timer_name##_time.stop();                                                                  \
    } while (0);

#ifdef DEBUG_MODE
#define IF_DEBUG_MODE(stmt) stmt;
#else
#define IF_DEBUG_MODE(stmt)
#endif
