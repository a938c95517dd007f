"""CLI read-out fixture: the stdout of the UNMODIFIED reference executable (oracle/_ref/bis_ref, built from
/root/reference/main.cpp by oracle/Makefile) on a small matrix, for tests/test_cli_gpu.py to diff the
`||A*x_k - b||_2 = ...` lines (postprocessing.hpp:8-30) of lib/bis against.

    python tests/golden/make_cli_fixture.py

Matrix: 2-D 5-point finite-difference Laplacian on a 12 x 12 grid with a varying diagonal (4 + 0.01 * (row % 7)),
written here in Matrix Market "coordinate real general" form (not a copy of any reference data file).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(ROOT, "oracle", "_ref", "bis_ref")
N = 12

# no GMRES case: the reference executable reads y[restart_len] one past the end in get_explicit_x (gmres.hpp:358,
# SURVEY F6); as a stand-alone binary that word is heap garbage and `-gm -p j` never converges on this matrix
CASES = [("cg_sgs", ["-cg", "-p", "sgs"]), ("bi_j", ["-bi", "-p", "j"]), ("sgs", ["-sgs"]), ("j", ["-j"])]


def write_mtx(path):
    ent = []
    for y in range(N):
        for x in range(N):
            r = y * N + x
            for dy, dx in ((-1, 0), (0, -1), (0, 0), (0, 1), (1, 0)):
                yy, xx = y + dy, x + dx
                if 0 <= yy < N and 0 <= xx < N:
                    v = (4.0 + 0.01 * (r % 7)) if (dy == 0 and dx == 0) else -1.0
                    ent.append((r + 1, yy * N + xx + 1, v))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{N * N} {N * N} {len(ent)}\n")
        for r, c, v in ent:
            f.write(f"{r} {c} {v!r}\n")


def main():
    assert os.path.exists(BIN), "make -C oracle ref"
    mtx = os.path.join(OUT, "cli_fdm2d12.mtx")
    write_mtx(mtx)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for key, flags in CASES:
        if flags is None:
            continue
        out = subprocess.run([BIN, mtx] + flags, capture_output=True, text=True, env=env, timeout=120)
        assert out.returncode == 0, out.stderr
        with open(os.path.join(OUT, f"cli_fdm2d12_{key}.txt"), "w") as f:
            f.write(out.stdout)
        print(key, len(out.stdout.splitlines()), "lines")


if __name__ == "__main__":
    sys.exit(main())
