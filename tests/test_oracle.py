"""The oracle itself is pinned: port == reference binary == committed golden vectors (CPU only)."""
import os

import numpy as np
import pytest

from apf_quadruped_b200 import dogbot

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KINDS = ["stance", "crawl", "trot", "mpc"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("which", ["reference", "port"])
def test_oracle_reproduces_golden_solves(oracle, which, kind):
    if not oracle.available(which):
        pytest.skip(f"oracle '{which}' not built")
    g = np.load(os.path.join(GOLD, "solves.npz"))
    count = len(g[f"{kind}_flag"])
    b = dogbot.generate(kind, count)
    out = oracle.solve_batch(b, kind=which)
    assert np.array_equal(out["flag"], g[f"{kind}_flag"])
    assert np.array_equal(out["iters"], g[f"{kind}_iters"])
    assert np.array_equal(out["fval"], g[f"{kind}_fval"])
    for k in ("x", "y", "z", "s"):
        assert np.array_equal(out[k], g[f"{kind}_{k}"]), k


@pytest.mark.parametrize("kind", ["stance", "trot"])
def test_port_equals_reference_everywhere(oracle, kind):
    if not (oracle.available("reference") and oracle.available("port")):
        pytest.skip("needs both oracles")
    b = dogbot.generate(kind, 40, start=3000)
    for kw in ({}, {"reltol": 1e-2, "abstol": 1e-2}, {"maxit": 4}, {"sigma_d": 0.1}):
        for i in range(b.count):
            args = (b.pattern, b.Ppr[i], b.Apr[i], b.Gpr[i], b.c[i], b.h[i], b.b[i])
            r = oracle.solve_sparse(*args, kind="reference", want_factor=True, **kw)
            p = oracle.solve_sparse(*args, kind="port", want_factor=True, **kw)
            assert (r.flag, r.iters, r.fval) == (p.flag, p.iters, p.fval)
            for k in ("x", "y", "z", "s"):
                assert np.array_equal(getattr(r, k), getattr(p, k))
                assert np.array_equal(r.init[k], p.init[k])
            assert np.array_equal(r.init["Lx"], p.init["Lx"]) and np.array_equal(r.init["D"], p.init["D"])
            for k in ("Kjc", "Kir", "Parent", "Lnz", "Lp", "Li"):
                assert np.array_equal(r.plan[k], p.plan[k])
            for k in ("n_rx", "n_ry", "n_rz", "n_mu", "alpha_p", "alpha_d", "sigma", "mu"):
                assert r.stats[k] == p.stats[k], k


def test_golden_solutions_are_kkt_points():
    """Independent sanity check of the fixtures: the stored optimal points satisfy the QP's KKT
    conditions to the solver tolerance (computed here with numpy, no solver involved)."""
    g = np.load(os.path.join(GOLD, "solves.npz"))
    for kind in KINDS:
        b = dogbot.generate(kind, 6)
        for i in range(6):
            if g[f"{kind}_flag"][i] != 0:
                continue
            P, A, G = b.dense(i)
            x, y, z, s = (g[f"{kind}_{k}"][i] for k in ("x", "y", "z", "s"))
            assert np.linalg.norm(P @ x + b.c[i] + G.T @ z + A.T @ y) < 1e-6
            assert np.linalg.norm(A @ x - b.b[i]) < 1e-6
            assert np.linalg.norm(G @ x + s - b.h[i]) < 1e-6
            assert np.all(s > 0) and np.all(z > 0) and s @ z / len(s) < 1e-6


def test_reference_dense_path_is_the_sparse_path_on_dogbot_patterns(oracle):
    """QP_SETUP_dense drops exact zeros (Auxilary.c:1154-1273); the DogBot blocks have none inside their
    patterns, so the reference's dense and sparse entries see the same QP -- the premise of the batched dense
    front-end's parity test (tests/test_dense_gpu.py)."""
    if not oracle.available("reference"):
        pytest.skip("needs the compiled reference")
    for kind in ("trot", "stance"):
        b = dogbot.generate(kind, 3, start=50)
        pat = b.pattern
        for row_major in (False, True):
            P, A, G = b.dense_batch(row_major=row_major)
            for i in range(b.count):
                d = oracle.solve_dense(pat.n, pat.m, pat.p, P[i], A[i], G[i], b.c[i], b.h[i], b.b[i], ordering=20 if row_major else 30)
                s = oracle.solve_sparse(pat, b.Ppr[i], b.Apr[i], b.Gpr[i], b.c[i], b.h[i], b.b[i], kind="reference")
                assert np.array_equal(d.plan["ccs"]["G"][1], pat.Gir) and np.array_equal(d.plan["ccs"]["G"][2], b.Gpr[i])
                assert np.array_equal(d.plan["P"], s.plan["P"])
                assert (d.flag, d.iters) == (s.flag, s.iters)
                for k in ("x", "y", "z", "s"):
                    assert np.array_equal(getattr(d, k), getattr(s, k))
