"""Regenerates the golden fixtures from the REAL reference (oracle/_ref, needs /root/reference).

    python tests/golden/make_golden.py

Fixtures (small, committed):
  plans.npz      -- P, Pinv, Parent, Lnz, Lp, Li, Kjc, Kir of the four DogBot patterns, as computed
                    by reference QP_SETUP (amd_l_order + LDL_symbolic + LDL_numeric's Li)
  amd_cases.npz  -- 160 random sparsity patterns (symmetric, triangular, unsymmetric, with dense
                    rows, unsorted) and the permutation reference amd_l_order returns for each
  solves.npz     -- reference x, y, z, s, flag, iterations, fval and the kkt_initialize point for
                    instances 0..23 of every pattern (inputs are regenerated from the seed by
                    apf_quadruped_b200.dogbot, so only outputs are stored)
The tests never trust these blindly where the live oracle is present: they re-derive and compare.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from apf_quadruped_b200 import dogbot  # noqa: E402
from oracle import reference as R  # noqa: E402

KINDS = ["stance", "crawl", "trot", "mpc"]
NSOLVE = 24


def amd_cases(count=160, seed=20261018):
    rng = np.random.default_rng(seed)
    ptr, idx, perm, stat, sizes = [], [], [], [], []
    for t in range(count):
        n = int(rng.integers(1, 90))
        A = sp.random(n, n, density=float(rng.choice([0.02, 0.06, 0.15, 0.4])), random_state=int(rng.integers(1 << 31)), format="csc")
        mode = t % 4
        if mode == 0:
            A = (A + A.T).tocsc()
        elif mode == 1:
            A = sp.triu(A).tocsc()
        elif mode == 3:
            A = (A + A.T).tolil()
            for r in rng.integers(0, n, size=2):
                A[r, :] = 1
                A[:, r] = 1
            A = A.tocsc()
        A.sort_indices()
        Ap, Ai = A.indptr.astype(np.int64), A.indices.astype(np.int64)
        if t % 5 == 2:  # unsorted columns -> the AMD_preprocess path
            Ai = Ai.copy()
            for j in range(n):
                Ai[Ap[j]:Ap[j + 1]] = Ai[Ap[j]:Ap[j + 1]][::-1]
        st, P = R.amd_order(n, Ap, Ai)
        ptr.append(Ap); idx.append(Ai); perm.append(P); stat.append(st); sizes.append(n)
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.int64)
    return {"n": np.array(sizes), "status": np.array(stat), "Ap": cat(ptr), "Ai": cat(idx), "P": cat(perm),
            "Ap_off": np.cumsum([0] + [len(a) for a in ptr]), "Ai_off": np.cumsum([0] + [len(a) for a in idx]),
            "P_off": np.cumsum([0] + [len(a) for a in perm])}


def main():
    assert R.available("reference"), "build oracle/_ref first (make -C oracle ref)"
    plans, solves = {}, {}
    for kind in KINDS:
        b = dogbot.generate(kind, NSOLVE)
        sols = [R.solve_sparse(b.pattern, b.Ppr[i], b.Apr[i], b.Gpr[i], b.c[i], b.h[i], b.b[i], kind="reference") for i in range(NSOLVE)]
        for k, v in sols[0].plan.items():
            plans[f"{kind}_{k}"] = np.asarray(v)
        for k in ("x", "y", "z", "s"):
            solves[f"{kind}_{k}"] = np.stack([getattr(s, k) for s in sols])
            solves[f"{kind}_init_{k}"] = np.stack([s.init[k] for s in sols])
        solves[f"{kind}_flag"] = np.array([s.flag for s in sols])
        solves[f"{kind}_iters"] = np.array([s.iters for s in sols])
        solves[f"{kind}_fval"] = np.array([s.fval for s in sols])
        print(kind, "iters", solves[f"{kind}_iters"].tolist())
    np.savez_compressed(os.path.join(HERE, "plans.npz"), **plans)
    np.savez_compressed(os.path.join(HERE, "solves.npz"), **solves)
    np.savez_compressed(os.path.join(HERE, "amd_cases.npz"), **amd_cases())
    for f in ("plans.npz", "solves.npz", "amd_cases.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
