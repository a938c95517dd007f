// host/utilities.hpp -- command-line parsing with the reference's grammar
// (utilities/utilities.hpp:12-108):
//   <matrix> <-j|-gs|-sgs|-cg|-gm|-bi> [-p j|gs|bgs|sgs|2st|s2st|ilu0] [-scale 0|1] [-rl N]
// plus `-dev N` (CUDA device) added by the build.  Unknown arguments only warn,
// as in the reference (utilities.hpp:103-106).  `-p ilut` is advertised by the
// reference's README but not implemented there: rejected here
// with the reference's own message.
#pragma once

#include "common.hpp"

inline void parse_cli(Args *cli_args, int argc, char *argv[], bool bench_mode = false) {
    if ((argc < 2 && bench_mode) || (argc < 3 && !bench_mode))
        bis_fatal(std::string("ERROR: parse_cli: Not enough arguments given. A call should contain:\n") +
                  argv[0] + " <matrix> <method> [extra_features]");
    cli_args->matrix_file_name = argv[1];
    int first = 2;
    if (!bench_mode) {
        ++first;
        static const std::map<std::string, SolverType> methods = {
            {"-j", SolverType::Jacobi},           {"-gs", SolverType::GaussSeidel},
            {"-sgs", SolverType::SymmetricGaussSeidel}, {"-cg", SolverType::ConjugateGradient},
            {"-gm", SolverType::GMRES},           {"-bi", SolverType::BiCGSTAB}};
        auto it = methods.find(argv[2]);
        if (it == methods.end())
            bis_fatal("ERROR: parse_cli: Please choose an available solver:"
                      "\n-j (Jacobi)\n-gs (Gauss-Seidel)\n-sgs (Symmetric Gauss-Seidel)"
                      "\n-gm ([Preconditioned] GMRES)\n-cg ([Preconditioned] Conjugate Gradient)"
                      "\n-bi ([Preconditioned] BiCGSTAB)");
        cli_args->method = it->second;
    }
    static const std::map<std::string, PrecondType> preconds = {
        {"j", PrecondType::Jacobi},          {"gs", PrecondType::GaussSeidel},
        {"bgs", PrecondType::BackwardsGaussSeidel}, {"sgs", PrecondType::SymmetricGaussSeidel},
        {"2st", PrecondType::TwoStageGS},    {"s2st", PrecondType::SymmetricTwoStageGS},
        {"ilu0", PrecondType::ILU0}};
    for (int i = first; i < argc; ++i) {
        const std::string arg = argv[i];
        auto need_value = [&]() {
            if (i + 1 >= argc)
                bis_fatal(std::string("ERROR: parse_cli: Not enough arguments given. Some extra features"
                                      " need additional arguments. Example:\n") + argv[0] + " <matrix> <method> -p gs");
        };
        if (arg == "-p") {
            need_value();
            auto it = preconds.find(argv[++i]);
            if (it == preconds.end())
                bis_fatal("ERROR: assign_cli_inputs: Please choose an available preconditioner type: "
                          "\n-p j (Jacobi)\n-p gs (Gauss-Seidel)\n-p bgs (Backwards Gauss-Seidel)"
                          "\n-p sgs (Symmetric Gauss-Seidel)\n-p 2st (2 Stage Gauss-Seidel)"
                          "\n-p s2st (Symmetric 2 Stage Gauss-Seidel)\n-p ilu0 (Incomplete LU with 0 fill-in)");
            cli_args->preconditioner = it->second;
        } else if (arg == "-scale") {
            need_value();
            cli_args->num_scale = (bool)std::atoi(argv[++i]);
        } else if (arg == "-rl") {
            need_value();
            cli_args->restart_length = std::atoi(argv[++i]);
        } else if (arg == "-dev") {
            need_value();
            cli_args->device = std::atoi(argv[++i]);
        } else if (arg == "-q") {
            cli_args->quiet = true;
        } else {
            std::cout << "ERROR: assign_cli_inputs: Arguement \"" << arg << "\" not recongnized." << std::endl;
        }
    }
}
