"""Boundary behaviour of the C API on the device: verbose output, statistics, argument checks, threads."""
import os
import re
import subprocess
import sys
import threading

import numpy as np
import pytest

from apf_quadruped_b200 import dogbot

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _printed(which, kind, count, maxit):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "helpers", "verbose_solve.py"), which, kind, str(count), str(maxit)],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.splitlines()


@pytest.mark.parametrize("kind,maxit", [("stance", 100), ("trot", 6)])
def test_verbose_prints_the_references_lines(qps, oracle, kind, maxit):
    """verbose = 1: the per-iteration lines (qpSWIFT.c:513-516, 598-601), header and footer, text for text; only
    the three timing lines of the footer carry other numbers."""
    if not oracle.available("reference"):
        pytest.skip("needs the compiled reference (oracle/_ref)")
    ours, ref = _printed("ours", kind, 3, maxit), _printed("ref", kind, 3, maxit)
    timing = re.compile(r"^(Solve Time|KKT_Solve Time|LDL Time)\s*: [-0-9.e+]+ ms$")
    strip = lambda lines: [l for l in lines if not timing.match(l)]
    assert sum(1 for l in ours if l.startswith("It: ")) > 3
    assert strip(ours) == strip(ref)
    assert sum(1 for l in ours if timing.match(l)) == sum(1 for l in ref if timing.match(l))


def test_ldl_numeric_time_is_reported(qps):
    batch = dogbot.generate("stance", 1, start=5)
    got = qps.solve_single(batch.pattern, batch.Ppr[0], batch.Apr[0], batch.Gpr[0], batch.c[0], batch.h[0], batch.b[0])
    st = got["stats"]
    # (iterations + 1) factorisations of a few microseconds each on the device clock (Auxilary.c:476-484)
    assert 1e-7 < st["ldl_numeric"] < st["tsolve"]
    assert st["kkt_time"] > 0


def test_device_entry_checks_the_capacity(qps):
    torch = pytest.importorskip("torch")
    batch = dogbot.generate("trot", 8)
    solver = qps.BatchSolver(batch.pattern, 4)
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(getattr(batch, k)).to(dev) for k in ("Ppr", "Apr", "Gpr", "c", "h", "b")}
    pat = batch.pattern
    out = [torch.empty((8, w), dtype=torch.float64, device=dev) for w in (pat.n, pat.p, pat.m, pat.m)]
    fl, it = torch.empty(8, dtype=torch.int32, device=dev), torch.empty(8, dtype=torch.int32, device=dev)
    with pytest.raises(RuntimeError):
        solver.solve_device(8, qps.default_settings(), *[t[k].data_ptr() for k in ("Ppr", "Apr", "Gpr", "c", "h", "b")],
                            *[o.data_ptr() for o in out], fl.data_ptr(), it.data_ptr())
    assert "capacity" in qps.last_error()
    with pytest.raises(RuntimeError):
        solver.read_stats(8)
    solver.close()


def test_qp_solve_from_several_threads(qps, oracle):
    """Independent QPs of ONE pattern solved concurrently (they share the pattern's single-QP batch): each gets its
    own sigma_d, flag, fval and sigma."""
    batch = dogbot.generate("stance", 8, start=900)
    pat = batch.pattern
    sig = [0.0, 1e-3] * 4
    refs = [oracle.solve_sparse(pat, batch.Ppr[i], batch.Apr[i], batch.Gpr[i], batch.c[i], batch.h[i], batch.b[i], sigma_d=sig[i]) for i in range(8)]
    got = [None] * 8

    def work(i):
        for _ in range(5):
            got[i] = qps.solve_single(pat, batch.Ppr[i], batch.Apr[i], batch.Gpr[i], batch.c[i], batch.h[i], batch.b[i], sigma_d=sig[i])

    threads = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for i in range(8):
        assert got[i]["flag"] == refs[i].flag and got[i]["iters"] == refs[i].iters
        assert got[i]["fval"] == refs[i].fval and got[i]["sigma"] == refs[i].stats["sigma"]
        assert np.array_equal(got[i]["x"], refs[i].x)
