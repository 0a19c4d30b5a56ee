"""Boundary hygiene of the single-problem C API that needs no GPU: QP_SETUP keeps no table of live QPs, and
its per-pattern plan cache is bounded ("no hidden global tables that grow per setup")."""
import numpy as np

from apf_quadruped_b200 import dogbot


def _setup(qps, batch, i, cleanup):
    pat = batch.pattern
    lb = qps.lib()
    keep = [np.ascontiguousarray(a, dtype=np.int64) for a in (pat.Pjc, pat.Pir, pat.Ajc, pat.Air, pat.Gjc, pat.Gir)]
    vals = [np.ascontiguousarray(a, dtype=np.float64).copy() for a in (batch.Ppr[i], batch.Apr[i], batch.Gpr[i], batch.c[i], batch.h[i], batch.b[i])]
    pl, pd = qps._pl, qps._pd
    qp = lb.QP_SETUP(pat.n, pat.m, pat.p, pl(keep[0]), pl(keep[1]), pd(vals[0]), pl(keep[2]), pl(keep[3]), pd(vals[1]),
                     pl(keep[4]), pl(keep[5]), pd(vals[2]), pd(vals[3]), pd(vals[4]), pd(vals[5]), 0.0, None)
    assert qp.contents.stats.contents.AMD_RESULT >= 0
    if cleanup:
        lb.QP_CLEANUP(qp)
    return keep, vals


def test_setups_without_cleanup_grow_nothing_inside_the_library(qps):
    lb = qps.lib()
    batch = dogbot.generate("trot", 4)
    _setup(qps, batch, 0, True)
    before = lb.QPB_PLAN_CACHE_INFO(0)
    held = [_setup(qps, batch, i % 4, False) for i in range(2000)]  # the controller never calls QP_CLEANUP (main.cpp:1649)
    assert lb.QPB_PLAN_CACHE_INFO(0) == before  # one pattern, one plan; and there is no table of QPs to grow
    del held


def test_plan_cache_is_bounded(qps):
    lb = qps.lib()
    cap = lb.QPB_PLAN_CACHE_INFO(1)
    rng = np.random.default_rng(3)
    base = dogbot.generate("trot", 1)
    pat = base.pattern
    n, m, p = pat.n, pat.m, pat.p
    for trial in range(cap + 8):  # distinct patterns: a dense G with one entry knocked out
        G = rng.standard_normal((m, n))
        G[trial % m, (7 * trial) % n] = 0.0
        P = np.eye(n)
        A = rng.standard_normal((p, n))
        c, h, b = np.zeros(n), np.ones(m), np.zeros(p)
        import ctypes as C
        qp = lb.QP_SETUP_dense(n, m, p, qps._pd(np.asfortranarray(P).ravel(order="F").copy()), qps._pd(np.asfortranarray(A).ravel(order="F").copy()),
                               qps._pd(np.asfortranarray(G).ravel(order="F").copy()), qps._pd(c), qps._pd(h), qps._pd(b), None, qps.COLUMN_MAJOR_ORDERING)
        lb.QP_CLEANUP_dense(qp)
    assert lb.QPB_PLAN_CACHE_INFO(0) <= cap
