"""SURVEY 8 f2 on the host: the numpy restatement of the controller's QP formulation (apf_quadruped_b200/wbc.py) against
the generator that every other test uses, and -- through the compiled reference -- as input of QP_SETUP_dense."""
import numpy as np
import pytest

from apf_quadruped_b200 import dogbot, wbc


@pytest.mark.parametrize("kind", ["stance", "trot", "crawl"])
def test_restatement_has_the_patterns_structure(kind):
    """Every structural entry of the pattern is filled by a formula and nothing lands outside the pattern."""
    S = wbc.random_states(16)
    Q, A, G, c, h, b = wbc.formulate_reference(kind, S)
    pat = dogbot.pattern(kind)
    _, Qm, Am, Gm = dogbot._PATTERN_CACHE[kind]
    assert Q.shape[1:] == (pat.n, pat.n) and A.shape[1:] == (pat.p, pat.n) and G.shape[1:] == (pat.m, pat.n)
    for M, mask in ((Q, Qm), (A, Am), (G, Gm)):
        assert not np.any(M[:, ~mask]), "a value outside the kind's sparsity pattern"
        assert np.all(np.any(M != 0, axis=0)[mask]), "a structural entry that no formula fills"
    rows = wbc.value_rows(kind, S)
    assert rows.Ppr.shape == (16, pat.nnzP) and rows.Apr.shape == (16, pat.nnzA) and rows.Gpr.shape == (16, pat.nnzG)


@pytest.mark.parametrize("kind", ["stance", "trot", "crawl"])
def test_restatement_agrees_with_the_generator(kind):
    """dogbot._wbc builds the same matrices with BLAS products; the restatement fixes the operation order.  Same numbers
    up to the rounding of the 6-term dot products, identical wherever no product is involved."""
    rng = np.random.Generator(np.random.Philox(key=[7, 7]))
    ph = dogbot._Physics(rng, 32)
    q = 0.5 * (dogbot.QMIN + dogbot.QMAX) + rng.uniform(-0.3, 0.3, size=(32, 12))
    dq = rng.uniform(-0.5, 0.5, size=(32, 12))
    k = 2.0 / dogbot.DT_LIMITS ** 2
    ph.ddqmin = k * (dogbot.QMIN - q - dogbot.DT_LIMITS * dq)
    ph.ddqmax = k * (dogbot.QMAX - q - dogbot.DT_LIMITS * dq)
    S = wbc.pack_states(ph.Jc, ph.Jj, ph.M, ph.bias, ph.Jdqd, ph.W, q, dq, ph.vdot_sw)
    _, (Q0, _), (A0, _), (G0, _), c0, h0, b0 = dogbot._wbc(kind, ph)
    Q, A, G, c, h, b = wbc.formulate_reference(kind, S)
    assert np.array_equal(A, A0) and np.array_equal(G, G0) and np.array_equal(b, b0) and np.array_equal(h, h0)
    assert np.allclose(Q, Q0, rtol=1e-14, atol=0) and np.allclose(c, c0, rtol=1e-13, atol=1e-12)


@pytest.mark.parametrize("kind", ["stance", "trot", "crawl"])
def test_reference_solves_the_restated_qps(kind):
    """The dense matrices of the restatement go through the compiled reference's QP_SETUP_dense + QP_SOLVE, and the
    explicit-pattern path (QP_SETUP with the plan's CCS) gives the same bits: the restatement produces exactly the
    pattern the plan was built for (no accidental zero inside it for these draws)."""
    from oracle import reference as R
    if not R.available("reference"):
        pytest.skip("oracle/_ref not built")
    S = wbc.random_states(6, start=100)
    Q, A, G, c, h, b = wbc.formulate_reference(kind, S)
    rows = wbc.value_rows(kind, S)
    pat = rows.pattern
    for i in range(6):
        dense = R.solve_dense(pat.n, pat.m, pat.p, np.asfortranarray(Q[i]).ravel(order="F"), np.asfortranarray(A[i]).ravel(order="F"),
                              np.asfortranarray(G[i]).ravel(order="F"), c[i], h[i], b[i])
        sparse = R.solve_sparse(pat, rows.Ppr[i], rows.Apr[i], rows.Gpr[i], rows.c[i], rows.h[i], rows.b[i])
        assert dense.flag == sparse.flag == 0 and dense.iters == sparse.iters
        for k in ("x", "y", "z", "s"):
            assert np.array_equal(getattr(dense, k), getattr(sparse, k))
    tau = wbc.torques_reference(kind, S, np.stack([R.solve_sparse(pat, rows.Ppr[i], rows.Apr[i], rows.Gpr[i], rows.c[i], rows.h[i], rows.b[i]).x for i in range(6)]))
    assert tau.shape == (6, 12) and np.all(np.isfinite(tau))


@pytest.mark.parametrize("kind", ["stance", "trot", "crawl"])
def test_library_formula_table_reproduces_the_restatement(kind):
    """The descriptor table the device kernel walks (built in C++ from the plan's own row lists, csrc/wbc.cuh), evaluated
    on the host with the kernel's operation order, gives the restatement's value rows bit for bit."""
    from apf_quadruped_b200 import qpswift
    S = wbc.random_states(24, start=5)
    rows = wbc.value_rows(kind, S)
    plan = qpswift.Plan(rows.pattern)
    d = wbc.descriptors(plan, kind)
    got = wbc.evaluate_descriptors(d, S)
    want = np.concatenate([rows.Ppr, rows.Apr, rows.Gpr, rows.c, rows.h, rows.b], axis=1)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    # the wrong kind for a pattern is refused
    other = "trot" if kind != "trot" else "stance"
    with pytest.raises(RuntimeError):
        wbc.descriptors(plan, other)
    plan.close()
