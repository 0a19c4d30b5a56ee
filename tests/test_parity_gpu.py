"""GPU parity: the CUDA path through the C ABI versus the CPU oracle, bit for bit.

The bar (BASELINE.json north_star): identical exit flags and iteration counts, solutions within
1e-8 relative.  Because the kernels reproduce the reference's operation order without FMA
contraction we assert the stronger property -- bit-identical x, y, z, s -- and keep the 1e-8 check
as the documented floor.
"""
import numpy as np
import pytest

from apf_quadruped_b200 import dogbot

pytestmark = pytest.mark.gpu

KINDS = ["stance", "crawl", "trot", "mpc"]
REL_TOL = 1e-8  # north_star tolerance: ||a-b||_inf / max(1, ||b||_inf)


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b))))) if a.size else 0.0


def check_against_oracle(out, ref, exact=True):
    assert np.array_equal(out["flag"], ref["flag"])
    assert np.array_equal(out["iters"], ref["iters"])
    for k in ("x", "y", "z", "s"):
        assert rel_err(out[k], ref[k]) <= REL_TOL, k
        if exact:
            assert np.array_equal(out[k], ref[k]), f"{k} not bit-identical"


@pytest.mark.parametrize("kind", KINDS)
def test_batch_matches_oracle(qps, oracle, kind):
    count = 96 if kind != "mpc" else 48
    batch = dogbot.generate(kind, count)
    solver = qps.BatchSolver(batch.pattern, count)
    out = solver.solve(batch)
    ref = oracle.solve_batch(batch)
    check_against_oracle(out, ref)
    fval, stat = solver.read_stats(count)
    assert np.array_equal(fval, ref["fval"])
    solver.close()


@pytest.mark.parametrize("kind", ["stance", "trot"])
def test_controller_tolerances(qps, oracle, kind):
    """reltol = abstol = 1e-2 as the controller sets them (main.cpp:1651-1652)."""
    batch = dogbot.generate(kind, 64, start=5000)
    solver = qps.BatchSolver(batch.pattern, 64)
    out = solver.solve(batch, qps.default_settings(reltol=1e-2, abstol=1e-2))
    ref = oracle.solve_batch(batch, reltol=1e-2, abstol=1e-2)
    check_against_oracle(out, ref)
    solver.close()


def test_initial_point_and_maxit(qps, oracle):
    """maxit = 0 returns the kkt_initialize point with QP_MAXIT; small maxit stops early."""
    batch = dogbot.generate("stance", 8, start=77)
    solver = qps.BatchSolver(batch.pattern, 8)
    out = solver.solve(batch, qps.default_settings(maxit=0))
    assert np.all(out["flag"] == qps.QP_MAXIT) and np.all(out["iters"] == 0)
    for i in range(8):
        ref = oracle.solve_sparse(batch.pattern, batch.Ppr[i], batch.Apr[i], batch.Gpr[i], batch.c[i], batch.h[i],
                                  batch.b[i], maxit=0)
        for k in ("x", "y", "z", "s"):
            assert np.array_equal(out[k][i], ref.init[k])
    out = solver.solve(batch, qps.default_settings(maxit=3))
    ref = oracle.solve_batch(batch, maxit=3)
    check_against_oracle(out, ref)
    assert np.all(out["flag"] == qps.QP_MAXIT)
    solver.close()


def test_no_equalities(qps, oracle):
    """p = 0: A, b, y absent (qpSWIFT.c:98-104)."""
    rng = np.random.default_rng(3)
    n, m = 12, 20
    from apf_quadruped_b200.dogbot import Pattern, Batch, _ccs_from_mask
    Pm = np.eye(n, dtype=bool)
    Pm[0:4, 0:4] = True
    Gm = rng.random((m, n)) < 0.3
    Gm[np.arange(m), rng.integers(0, n, m)] = True
    Pjc, Pir, _ = _ccs_from_mask(Pm)
    Gjc, Gir, _ = _ccs_from_mask(Gm)
    z = np.zeros(n + 1, dtype=np.int64)
    pat = Pattern("noeq", n, m, 0, Pjc, Pir, z, np.zeros(0, dtype=np.int64), Gjc, Gir)
    B = 16
    M = rng.normal(size=(B, n, n))
    Pd = np.einsum("bij,bkj->bik", M, M) * Pm + 2 * np.eye(n)
    rr, cc = np.nonzero(Pm.T)[::-1]
    Ppr = Pd[:, rr, cc]
    gr, gc = np.nonzero(Gm.T)[::-1]
    Gpr = rng.normal(size=(B, len(gr)))
    batch = Batch(pat, np.ascontiguousarray(Ppr), np.zeros((B, 0)), Gpr, rng.normal(size=(B, n)),
                  rng.uniform(0.5, 2.0, size=(B, m)), np.zeros((B, 0)))
    solver = qps.BatchSolver(pat, B)
    out = solver.solve(batch)
    ref = oracle.solve_batch(batch)
    check_against_oracle(out, ref)
    solver.close()


def test_user_permutation_and_sigma_d(qps, oracle):
    batch = dogbot.generate("crawl", 16, start=900)
    pat = batch.pattern
    perm = np.random.default_rng(5).permutation(pat.N)
    solver = qps.BatchSolver(pat, 16, sigma_d=0.05, permut=perm)
    out = solver.solve(batch)
    ref = oracle.solve_batch(batch, sigma_d=0.05, permut=perm)
    check_against_oracle(out, ref)
    solver.close()


def test_single_problem_api(qps, oracle):
    """QP_SETUP / QP_SOLVE / QP_CLEANUP exactly as the controller calls them, batch of one."""
    batch = dogbot.generate("stance", 4, start=31)
    pat = batch.pattern
    for i in range(4):
        got = qps.solve_single(pat, batch.Ppr[i], batch.Apr[i], batch.Gpr[i], batch.c[i], batch.h[i], batch.b[i])
        ref = oracle.solve_sparse(pat, batch.Ppr[i], batch.Apr[i], batch.Gpr[i], batch.c[i], batch.h[i], batch.b[i])
        assert got["flag"] == ref.flag and got["iters"] == ref.iters
        for k in ("x", "y", "z", "s"):
            assert np.array_equal(got[k], getattr(ref, k))
        assert got["fval"] == ref.fval
        assert got["sigma"] == ref.stats["sigma"]
        assert np.array_equal(got["P"], ref.plan["P"])


def test_ldl_kernels(qps, oracle):
    """LDL_numeric and the triangular solves on their own, against the oracle's factor."""
    torch = pytest.importorskip("torch")
    batch = dogbot.generate("stance", 32, start=11)
    pat = batch.pattern
    solver = qps.BatchSolver(pat, 32)
    N, lnz = solver.plan.info("N"), solver.plan.info("lnz")
    vals = np.concatenate([batch.Ppr, batch.Apr, batch.Gpr], axis=1)
    dev = torch.device("cuda:0")
    dvals = torch.from_numpy(vals).to(dev)
    zdiag = torch.full((32, pat.m), -1.0, dtype=torch.float64, device=dev)
    Lx = torch.empty((32, lnz), dtype=torch.float64, device=dev)
    Dd = torch.empty((32, N), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    solver.ldl_numeric_device(32, dvals.data_ptr(), zdiag.data_ptr(), Lx.data_ptr(), Dd.data_ptr(), st)
    rhs = torch.from_numpy(np.concatenate([-batch.c, batch.b, batch.h], axis=1)).to(dev)
    sol = torch.empty_like(rhs)
    solver.ldl_solve_device(32, Lx.data_ptr(), Dd.data_ptr(), rhs.data_ptr(), sol.data_ptr(), st)
    torch.cuda.synchronize()
    for i in range(32):
        ref = oracle.solve_sparse(pat, batch.Ppr[i], batch.Apr[i], batch.Gpr[i], batch.c[i], batch.h[i], batch.b[i],
                                  maxit=0, want_factor=True)
        assert np.array_equal(Lx[i].cpu().numpy(), ref.init["Lx"])
        assert np.array_equal(Dd[i].cpu().numpy(), ref.init["D"])
        assert np.array_equal(sol[i, :pat.n].cpu().numpy(), ref.init["x"])
        assert np.array_equal(sol[i, pat.n:pat.n + pat.p].cpu().numpy(), ref.init["y"])
    solver.close()
