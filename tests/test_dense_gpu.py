"""GPU parity of the batched dense front-end (include/qpswift_b200.h QP_SOLVE_BATCH_DENSE,
QPB_DENSE_TO_VALUES_DEVICE) against the reference's own dense path: QP_SETUP_dense + QP_SOLVE
(dogbot_controller/src/qpSWIFT/qpSWIFT.c:260-456, densetosparse Auxilary.c:1154-1273), which is what the
controller calls every tick (dogbot_controller/src/client/main.cpp:1649-1667).

Bar: the CCS the reference makes of a dense QP is the plan's pattern with exactly the gathered values, and
x, y, z, s, the exit flag and the iteration count are bit-identical; a QP whose dense zero pattern differs from
the plan's is reported, never silently solved as something else.
"""
import numpy as np
import pytest

from apf_quadruped_b200 import dogbot

pytestmark = pytest.mark.gpu

COL, ROW = 30, 20  # GlobalOptions.h:59-60


@pytest.mark.parametrize("ordering", [COL, ROW])
@pytest.mark.parametrize("kind,count", [("trot", 48), ("stance", 33), ("crawl", 20)])
def test_dense_batch_matches_reference_dense_path(qps, oracle, kind, count, ordering):
    batch = dogbot.generate(kind, count, start=9100)
    pat = batch.pattern
    P, A, G = batch.dense_batch(row_major=ordering == ROW)
    solver = qps.BatchSolver(pat, count)
    out = solver.solve_dense(P, A, G, batch.c, batch.h, batch.b, ordering=ordering)
    assert not out["violation"].any()
    for i in range(count):
        ref = oracle.solve_dense(pat.n, pat.m, pat.p, P[i], A[i], G[i], batch.c[i], batch.h[i], batch.b[i], ordering=ordering)
        for name, jc, ir, vals in (("P", pat.Pjc, pat.Pir, batch.Ppr[i]), ("A", pat.Ajc, pat.Air, batch.Apr[i]), ("G", pat.Gjc, pat.Gir, batch.Gpr[i])):
            rjc, rir, rpr = ref.plan["ccs"][name]
            assert np.array_equal(rjc, jc) and np.array_equal(rir, ir) and np.array_equal(rpr, vals)
        assert ref.flag == out["flag"][i] and ref.iters == out["iters"][i]
        for k in ("x", "y", "z", "s"):
            assert np.array_equal(getattr(ref, k), out[k][i]), (i, k)
    solver.close()


def test_gather_on_the_device_is_exact_and_ragged_counts_work(qps):
    import torch
    for kind, count in (("trot", 1000), ("stance", 257), ("mpc", 65)):
        batch = dogbot.generate(kind, count, start=31)
        pat = batch.pattern
        solver = qps.BatchSolver(pat, count)
        for ordering in (COL, ROW):
            dense = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in batch.dense_batch(row_major=ordering == ROW)]
            rows = [torch.full((count, w), float("nan"), dtype=torch.float64, device="cuda") for w in (pat.nnzP, pat.nnzA, pat.nnzG)]
            viol = torch.full((count,), -1, dtype=torch.int32, device="cuda")
            ptr = lambda t: t.data_ptr() if t.numel() else None
            solver.dense_to_values_device(count, ordering, *[ptr(t) for t in dense], *[ptr(t) for t in rows], viol.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            for got, want in zip(rows, (batch.Ppr, batch.Apr, batch.Gpr)):
                assert np.array_equal(got.cpu().numpy(), want)
            assert not viol.cpu().numpy().any()
        solver.close()


def test_pattern_violations_are_reported(qps, oracle):
    batch = dogbot.generate("trot", 6, start=777)
    pat = batch.pattern
    P, A, G = batch.dense_batch()
    # QP 1: a nonzero where the plan has none (G(0, 29) is outside the trot pattern)
    cell = 29 * pat.m + 0
    assert G[1, cell] == 0.0
    G[1, cell] = 0.25
    # QP 3: an entry of the pattern is exactly zero (the reference would drop it)
    col = int(np.searchsorted(pat.Gjc, 5, side="right") - 1)
    zcell = col * pat.m + int(pat.Gir[5])
    assert G[3, zcell] == batch.Gpr[3, 5] != 0.0
    G[3, zcell] = 0.0
    # QP 4: -0.0 counts as zero, like `!= 0` in densetosparse
    G[4, cell] = -0.0
    solver = qps.BatchSolver(pat, 6)
    out = solver.solve_dense(P, A, G, batch.c, batch.h, batch.b)
    assert list(out["violation"]) == [0, 1, 0, 2, 0, 0]
    # QP 3 is solved as the explicit-pattern QP: identical with the reference's sparse setup on those values
    vals = batch.Gpr[3].copy()
    vals[5] = 0.0
    ref = oracle.solve_sparse(pat, batch.Ppr[3], batch.Apr[3], vals, batch.c[3], batch.h[3], batch.b[3])
    assert ref.flag == out["flag"][3] and ref.iters == out["iters"][3]
    for k in ("x", "y", "z", "s"):
        assert np.array_equal(getattr(ref, k), out[k][3])
    # ... whereas the reference's dense path orders a smaller system for it
    dref = oracle.solve_dense(pat.n, pat.m, pat.p, P[3], A[3], G[3], batch.c[3], batch.h[3], batch.b[3])
    assert dref.plan["ccs"]["G"][2].size == pat.nnzG - 1
    # the untouched QPs are what they were
    plain = solver.solve(batch)
    for i in (0, 2, 4, 5):
        for k in ("x", "y", "z", "s", "flag", "iters"):
            assert np.array_equal(plain[k][i], out[k][i])
    solver.close()
