"""BASELINE.json configs[1] at FULL size (65,536 trot QPs on one GPU), checked through properties that do not
need the oracle on every instance:

* optimality: every QP that exits OPTIMAL satisfies qpSWIFT's own stop test (qpSWIFT.c:520-534) when the
  residuals are recomputed on the host from x, y, z, s with plain numpy -- ||P x + c + A'y + G'z||,
  ||A x - b||, ||G x + s - h|| below reltol/sqrt(3), s'z/m below abstol, s > 0, z > 0;
* composition independence: the two halves solved on their own give the very bits of the whole batch
  (a QP's arithmetic does not depend on its neighbours, the kernel that ran it, or where a shard was cut);
* a sample of 512 instances spread over the batch is bit-identical with the oracle.
"""
import numpy as np
import pytest

from apf_quadruped_b200 import dogbot
from test_parity_gpu import check_against_oracle

pytestmark = pytest.mark.gpu

COUNT = 65536


def _residual_norms(batch, out):
    pat = batch.pattern
    x, y, z, s = out["x"], out["y"], out["z"], out["s"]
    rx = batch.c.copy()
    ry = -batch.b.copy()
    rz = s - batch.h

    def sweep(jc, ir, vals, into_rows, dual):
        cols = np.repeat(np.arange(pat.n), np.diff(np.asarray(jc)))
        for k, (i, j) in enumerate(zip(np.asarray(ir), cols)):
            v = vals[:, k]
            if into_rows is not None:
                into_rows[:, i] += v * x[:, j]   # (M x)_i
            if dual is not None:
                rx[:, j] += v * dual[:, i]        # (M'd)_j

    sweep(pat.Pjc, pat.Pir, batch.Ppr, rx, None)  # P is stored in full (both triangles): (P x)_i only
    sweep(pat.Ajc, pat.Air, batch.Apr, ry, y)
    sweep(pat.Gjc, pat.Gir, batch.Gpr, rz, z)
    norm = lambda a: np.sqrt((a * a).sum(axis=1))
    return norm(rx), norm(ry), norm(rz), (s * z).sum(axis=1) / pat.m


@pytest.mark.parametrize("kernel", ["warp", "stream"])
def test_full_size_batch_properties(qps, oracle, monkeypatch, kernel):
    """configs[1] at full size on the kernel the library picks for it (the specialised warp-per-QP kernel) and on the
    throughput kernel (QPB_STREAM_MIN as for patterns without a specialised warp kernel)."""
    if kernel == "stream":
        monkeypatch.setenv("QPB_STREAM_MIN", "2048")
    batch = dogbot.generate("trot", COUNT)
    solver = qps.BatchSolver(batch.pattern, COUNT)
    out = solver.solve(batch)
    assert solver.last_kernel == kernel and solver.last_kernel_generated
    ok = out["flag"] == qps.QP_OPTIMAL
    assert ok.mean() > 0.99 and set(np.unique(out["flag"])) <= {qps.QP_OPTIMAL, qps.QP_MAXIT}
    assert 8 <= out["iters"][ok].min() and out["iters"][ok].max() <= 40
    # qpSWIFT's stop test, recomputed on the host (reltol = abstol = 1e-6 by default, qpSWIFT.c:178-180)
    n_rx, n_ry, n_rz, mu = _residual_norms(batch, out)
    thr = 1e-6 / np.sqrt(3.0)
    slack = 1.0 + 1e-6  # the host sums in another order than Auxilary.c:745-786
    for name, v in (("rx", n_rx), ("ry", n_ry), ("rz", n_rz)):
        assert (v[ok] < thr * slack).all(), name
    assert (mu[ok] < 1e-6 * slack).all()
    assert (out["s"][ok] > 0).all() and (out["z"][ok] > 0).all()
    # composition independence
    half = COUNT // 2
    for lo, hi in ((0, half), (half, COUNT)):
        part = solver.solve(batch.slice(lo, hi))
        for k in ("x", "y", "z", "s", "flag", "iters"):
            assert np.array_equal(part[k], out[k][lo:hi]), (lo, k)
    # a sample against the oracle
    pick = np.linspace(0, COUNT - 1, 512).astype(int)
    sample = dogbot.Batch(batch.pattern, *[np.ascontiguousarray(getattr(batch, f)[pick]) for f in ("Ppr", "Apr", "Gpr", "c", "h", "b")])
    ref = oracle.solve_batch(sample)
    check_against_oracle({k: out[k][pick] for k in ("x", "y", "z", "s", "flag", "iters")}, ref)
    solver.close()


@pytest.mark.parametrize("kind,count", [("stance", 32768), ("crawl", 32768), ("mpc", 32768)])
def test_other_patterns_at_size(qps, oracle, kind, count):
    """The other BASELINE.json patterns at 32,768 QPs each (configs[2]'s stance and crawl shares on the specialised
    warp-per-QP kernel; configs[4]'s horizon-10 MPC QP on the interpreted one): qpSWIFT's stop test recomputed on the
    host for every QP, composition independence, and a spread sample against the oracle."""
    batch = dogbot.generate_parallel(kind, count)
    solver = qps.BatchSolver(batch.pattern, count)
    out = solver.solve(batch)
    assert solver.last_kernel == "warp" and solver.last_kernel_generated == (kind != "mpc")
    ok = out["flag"] == qps.QP_OPTIMAL
    assert ok.all()
    n_rx, n_ry, n_rz, mu = _residual_norms(batch, out)
    thr, slack = 1e-6 / np.sqrt(3.0), 1.0 + 1e-6
    assert (n_rx < thr * slack).all() and (n_ry < thr * slack).all() and (n_rz < thr * slack).all() and (mu < 1e-6 * slack).all()
    assert (out["s"] > 0).all() and (out["z"] > 0).all()
    lo, hi = count // 3, count // 3 + 5000
    part = solver.solve(batch.slice(lo, hi))
    for k in ("x", "y", "z", "s", "flag", "iters"):
        assert np.array_equal(part[k], out[k][lo:hi]), k
    pick = np.linspace(0, count - 1, 128).astype(int)
    sample = dogbot.Batch(batch.pattern, *[np.ascontiguousarray(getattr(batch, f)[pick]) for f in ("Ppr", "Apr", "Gpr", "c", "h", "b")])
    ref = oracle.solve_batch(sample)
    check_against_oracle({k: out[k][pick] for k in ("x", "y", "z", "s", "flag", "iters")}, ref)
    solver.close()
