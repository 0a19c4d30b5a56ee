"""The rollout of apf_quadruped_b200/rollout.py on the CPU.  TEST INFRASTRUCTURE ONLY (nothing under apf_quadruped_b200/ imports it).

The numpy restatement of the tick model (``rollout.tick_reference``) + the CPU oracle solving every QP (``QP_SETUP`` +
``QP_SOLVE`` per QP, cold start, the controller's tolerances) + the numpy restatement of the formulation and of the torques
(``wbc.value_rows``, ``wbc.torques_reference``).  The device loop is compared with this bit for bit, tick by tick
(``tests/test_rollout_gpu.py``, and inside every ``bench.py`` run that includes the rollout)."""
import numpy as np

from apf_quadruped_b200 import wbc
from apf_quadruped_b200.rollout import DEFAULT_SEED, KIND_NAMES, initial, tick_reference


def reference_rollout(robots: int, ticks: int, first: int = 0, seed: int = DEFAULT_SEED, reltol=1e-2, abstol=1e-2, record=None):
    """The rollout on the CPU: the numpy model + the CPU oracle (QP_SETUP + QP_SOLVE per QP, cold start) + the numpy torque
    restatement.  TEST INFRASTRUCTURE (imports oracle/).  Returns the final (S, Rw, x, tau) and per-tick records."""
    from oracle import reference as Rf
    S, Rw = initial(robots, first, seed)
    x = np.zeros((robots, 30))
    tau = np.zeros((robots, 12))
    flags = np.zeros(robots, dtype=np.int64)
    iters = np.zeros(robots, dtype=np.int64)
    log = []
    for t in range(ticks):
        kind = tick_reference(S, Rw, x, t, first, seed)
        for k, name in enumerate(KIND_NAMES):
            idx = np.nonzero(kind == k)[0]
            if idx.size == 0:
                continue
            rows = wbc.value_rows(name, S[idx])
            ref, _ = Rf.run_batch(rows, reltol=reltol, abstol=abstol)
            x[idx] = ref["x"]
            flags[idx], iters[idx] = ref["flag"], ref["iters"]
            tau[idx] = wbc.torques_reference(name, S[idx], ref["x"])
        log.append({"kind": kind.copy(), "x": x.copy(), "tau": tau.copy(), "flag": flags.copy(), "iters": iters.copy()})
        if record is not None:
            record(t, S, Rw, x, tau)
    return S, Rw, x, tau, log


def check_device_rollout(torch, device, robots=64, ticks=50):
    """A small device rollout against this CPU rollout, bit for bit (bench.py's in-run check)."""
    from apf_quadruped_b200.rollout import Rollout
    small = Rollout(torch, robots, device=device, first=0)
    small.step(ticks)
    got = small.snapshot()
    small.close()
    Sr, Rwr, xr, taur, log = reference_rollout(robots, ticks, first=0)
    same = bool(np.array_equal(got["x"], xr) and np.array_equal(got["tau"], taur) and np.array_equal(got["S"], Sr) and np.array_equal(got["Rw"], Rwr)
                and np.array_equal(got["iters"], log[-1]["iters"]))
    if not same:
        raise AssertionError("rollout: the device loop differs from the CPU oracle stepping the same model")
    return {"robots": robots, "ticks": ticks, "bit_identical_with_cpu_oracle": same, "qp_solves_checked": robots * ticks}
