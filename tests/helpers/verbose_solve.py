"""Prints what QP_SOLVE prints with verbose = 1 for a few DogBot QPs: `ours` through libqpswift_b200.so,
`ref` through the compiled reference.  Run as a subprocess by tests/test_boundary_gpu.py (the libraries print
with C printf, which pytest's capture does not see per call)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from apf_quadruped_b200 import dogbot  # noqa: E402


def main():
    which, kind, count, maxit = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    batch = dogbot.generate(kind, count, start=77)
    pat = batch.pattern
    for i in range(count):
        args = (pat, batch.Ppr[i], batch.Apr[i], batch.Gpr[i], batch.c[i], batch.h[i], batch.b[i])
        if which == "ours":
            from apf_quadruped_b200 import qpswift
            qpswift.solve_single(*args, verbose=1, maxit=maxit)
        else:
            from oracle import reference
            reference.solve_sparse(*args, verbose=1, maxit=maxit)
        sys.stdout.flush()


if __name__ == "__main__":
    main()
