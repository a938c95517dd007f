"""Labelled permutation modes (perm_mode = 1 multicolouring, 2 BFS levels, 3 reverse Cuthill-McKee) next to the natural
ordering: iterations to convergence, ms per iteration and time to solution on one B200:  python tools/perm_compare.py [n ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi, host  # noqa: E402

sizes = [int(v) for v in sys.argv[1:]] or [128, 256]
with capi.Context(0) as ctx:
    for n in sizes:
        for method, pre in (("cg", "sgs"), ("sgs", "none")):
            row = {"matrix": f"HPCG-{n}", "method": method, "precond": pre}
            for mode in (0, 1, 2, 3):
                ctx.set_option("perm_mode", mode)
                host.solve(ctx, method, pre, matrix_name=f"HPCG-{n}", want_x=False, max_iters=5)          # warm-up
                r = host.solve(ctx, method, pre, matrix_name=f"HPCG-{n}", want_x=False)
                tag = ("natural", "coloured", "bfs", "rcm")[mode]
                row[tag] = {"iterations": r.iter_count, "converged": r.converged, "solve_s": r.solve_time,
                            "ms_per_iter": 1e3 * r.solve_time / max(r.iter_count, 1), "preprocessing_s": r.preprocessing_time,
                            "final_true_residual_rel": r.final_true_residual / r.history[0]}
            ctx.set_option("perm_mode", 0)
            row["time_to_solution_ratio"] = row["natural"]["solve_s"] / row["coloured"]["solve_s"]
            print(json.dumps(row), flush=True)
