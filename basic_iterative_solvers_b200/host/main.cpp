// host/main.cpp -- the command-line driver (reference main.cpp:72-99).
//   bis <matrix> <-j|-gs|-sgs|-cg|-gm|-bi> [-p j|gs|bgs|sgs|2st|s2st|ilu0] [-rl N] [-dev N]
// <matrix> is a MatrixMarket file, HPCG-<n>, HPCG-<nx>-<ny>-<nz> or
// Anderson,Lx=..,Ly=..,Lz=..,ranpot=..  Fatal errors print to stderr and exit
// with EXIT_FAILURE like the reference.
#include "run.hpp"

int main(int argc, char *argv[]) {
    Timers timers;
    Args cli_args;
    bis_context *dev = nullptr;
    int rc = EXIT_SUCCESS;
    try {
        parse_cli(&cli_args, argc, argv);
        if (bis_context_create(cli_args.device, &dev) != 0)
            bis_fatal(std::string("cannot create the device context: ") + bis_last_error());
        TIME(timers.total, run(&cli_args, &timers, dev))
        print_timers(&cli_args, &timers);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        rc = EXIT_FAILURE;
    }
    if (dev) bis_context_destroy(dev);
    return rc;
}
