"""SURVEY 8 f2 / f1 on the device: value rows built by wbc_formulate_kernel from packed robot states, the solve on those
rows, and the torques -- each against the numpy restatement (apf_quadruped_b200/wbc.py) and against the compiled
reference's QP_SETUP_dense + QP_SOLVE on the restatement's dense matrices, bit for bit."""
import numpy as np
import pytest

from apf_quadruped_b200 import dogbot, qpswift, wbc

pytestmark = pytest.mark.gpu


def _device_rows(torch, pat, count, dev):
    return {k: torch.zeros((count, w), dtype=torch.float64, device=dev) for k, w in
            (("Ppr", pat.nnzP), ("Apr", pat.nnzA), ("Gpr", pat.nnzG), ("c", pat.n), ("h", pat.m), ("b", pat.p))}


@pytest.mark.parametrize("kind,count", [("stance", 300), ("trot", 1000), ("crawl", 257)])
def test_formulate_solve_torques_on_the_device(kind, count):
    import torch
    from oracle import reference as R
    dev = torch.device("cuda", 0)
    S = wbc.random_states(count, start=11)
    want = wbc.value_rows(kind, S)
    pat = want.pattern
    solver = qpswift.BatchSolver(pat, count, device=0)
    f = wbc.Formulator(solver, kind)
    dS = torch.from_numpy(S).to(dev)
    rows = _device_rows(torch, pat, count, dev)
    st = torch.cuda.current_stream().cuda_stream
    f.formulate(count, dS.data_ptr(), *[rows[k].data_ptr() for k in ("Ppr", "Apr", "Gpr", "c", "h", "b")], stream=st)
    torch.cuda.synchronize()
    for k in ("Ppr", "Apr", "Gpr", "c", "h", "b"):
        assert np.array_equal(rows[k].cpu().numpy(), getattr(want, k)), f"{k} differs from the numpy restatement"
    # solve the rows where they lie, torques from the solution
    out = {"x": torch.empty((count, pat.n), dtype=torch.float64, device=dev), "y": torch.empty((count, pat.p), dtype=torch.float64, device=dev),
           "z": torch.empty((count, pat.m), dtype=torch.float64, device=dev), "s": torch.empty((count, pat.m), dtype=torch.float64, device=dev),
           "flag": torch.empty(count, dtype=torch.int32, device=dev), "iters": torch.empty(count, dtype=torch.int32, device=dev)}
    opts = qpswift.default_settings(reltol=1e-2, abstol=1e-2)  # the controller's tolerances (main.cpp:1651-1652)
    solver.solve_device(count, opts, *[rows[k].data_ptr() for k in ("Ppr", "Apr", "Gpr", "c", "h", "b")],
                        out["x"].data_ptr(), out["y"].data_ptr(), out["z"].data_ptr(), out["s"].data_ptr(), out["flag"].data_ptr(), out["iters"].data_ptr(), None, None, st)
    tau = torch.empty((count, 12), dtype=torch.float64, device=dev)
    f.torques(count, dS.data_ptr(), out["x"].data_ptr(), tau.data_ptr(), stream=st)
    torch.cuda.synchronize()
    x = out["x"].cpu().numpy()
    assert np.array_equal(tau.cpu().numpy(), wbc.torques_reference(kind, S, x)), "torques differ from the numpy restatement"
    # the compiled reference on the restatement's DENSE matrices, as the controller calls it (QP_SETUP_dense + QP_SOLVE)
    Q, A, G, c, h, b = wbc.formulate_reference(kind, S[:24])
    for i in range(24):
        ref = R.solve_dense(pat.n, pat.m, pat.p, Q[i].ravel(order="F"), A[i].ravel(order="F"), G[i].ravel(order="F"), c[i], h[i], b[i], reltol=1e-2, abstol=1e-2)
        assert ref.flag == int(out["flag"][i]) and ref.iters == int(out["iters"][i])
        assert np.array_equal(ref.x, x[i]) and np.array_equal(ref.z, out["z"][i].cpu().numpy()) and np.array_equal(ref.s, out["s"][i].cpu().numpy())
        assert np.array_equal(ref.y, out["y"][i].cpu().numpy())
    solver.close()


def test_listed_robots_only():
    """Index list + count in device memory (what the rollout's per-pattern lists are): listed rows are formulated, solved
    and given torques; the others are not touched."""
    import torch
    dev = torch.device("cuda", 0)
    count, kind = 400, "trot"
    S = wbc.random_states(count, start=3)
    want = wbc.value_rows(kind, S)
    pat = want.pattern
    solver = qpswift.BatchSolver(pat, count, device=0)
    f = wbc.Formulator(solver, kind)
    listed = np.array([7, 399, 0, 128, 129, 33, 250], dtype=np.int32)
    dindex = torch.from_numpy(np.concatenate([listed, np.full(count - len(listed), -1, dtype=np.int32)])).to(dev)
    dcount = torch.tensor([len(listed)], dtype=torch.int32, device=dev)
    dS = torch.from_numpy(S).to(dev)
    rows = _device_rows(torch, pat, count, dev)
    names = ("Ppr", "Apr", "Gpr", "c", "h", "b")
    st = torch.cuda.current_stream().cuda_stream
    f.formulate(count, dS.data_ptr(), *[rows[k].data_ptr() for k in names], dindex=dindex.data_ptr(), dcount=dcount.data_ptr(), stream=st)
    out = {"x": torch.full((count, pat.n), -7.0, dtype=torch.float64, device=dev), "y": torch.empty((count, pat.p), dtype=torch.float64, device=dev),
           "z": torch.empty((count, pat.m), dtype=torch.float64, device=dev), "s": torch.empty((count, pat.m), dtype=torch.float64, device=dev),
           "flag": torch.full((count,), -1, dtype=torch.int32, device=dev), "iters": torch.empty(count, dtype=torch.int32, device=dev)}
    opts = qpswift.default_settings()
    solver.solve_device_indexed(count, opts, dindex.data_ptr(), dcount.data_ptr(), *[rows[k].data_ptr() for k in names],
                                out["x"].data_ptr(), out["y"].data_ptr(), out["z"].data_ptr(), out["s"].data_ptr(), out["flag"].data_ptr(), out["iters"].data_ptr(), st)
    tau = torch.full((count, 12), -7.0, dtype=torch.float64, device=dev)
    f.torques(count, dS.data_ptr(), out["x"].data_ptr(), tau.data_ptr(), dindex=dindex.data_ptr(), dcount=dcount.data_ptr(), stream=st)
    torch.cuda.synchronize()
    mask = np.zeros(count, dtype=bool)
    mask[listed] = True
    for k in names:
        got = rows[k].cpu().numpy()
        assert np.array_equal(got[mask], getattr(want, k)[mask]) and not np.any(got[~mask])
    from oracle import reference as R
    sub = dogbot.Batch(pat, *[np.ascontiguousarray(getattr(want, k)[listed]) for k in names])
    ref, _ = R.run_batch(sub)
    x = out["x"].cpu().numpy()
    assert np.array_equal(x[listed], ref["x"]) and np.all(x[~mask] == -7.0)
    assert np.array_equal(out["flag"].cpu().numpy()[listed], ref["flag"]) and np.all(out["flag"].cpu().numpy()[~mask] == -1)
    t = tau.cpu().numpy()
    assert np.array_equal(t[listed], wbc.torques_reference(kind, S[listed], ref["x"])) and np.all(t[~mask] == -7.0)
    solver.close()
