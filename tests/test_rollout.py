"""SURVEY 8 f4 on the host: the rollout model in numpy (apf_quadruped_b200/rollout.py) with the CPU oracle in the loop."""
import numpy as np
import pytest

from apf_quadruped_b200 import rollout, wbc
from oracle import rollout_reference


def test_schedule_keeps_all_three_patterns_busy():
    robots = np.arange(512)
    for t in (0, 13, 99, 100, 777):
        k = rollout.kinds_at(t, robots)
        assert set(np.unique(k)) == {0, 1, 2}
        assert (k == 0).sum() > (k == 1).sum() > 0 and (k == 2).sum() > 0


def test_initial_states_are_reproducible_per_robot():
    S, Rw = rollout.initial(8, first=40)
    S2, Rw2 = rollout.initial(3, first=43)
    assert np.array_equal(S[3:6], S2) and np.array_equal(Rw[3:6], Rw2)
    assert S.shape == (8, wbc.STATE) and Rw.shape == (8, rollout.ROW)


def test_hash_is_the_splitmix_finaliser():
    h = rollout._hash(1, np.array([0, 1]), 0, 0, 0)
    assert h.dtype == np.uint64 and h[0] != h[1]
    assert np.array_equal(h, rollout._hash(1, np.array([0, 1]), 0, 0, 0))


def test_reference_rollout_closes_the_loop():
    """Six robots, 60 ticks (every robot goes through stance, trot and crawl), every QP solved by the CPU oracle: all optimal,
    the state stays bounded, the swing feet approach their footholds, and splitting the robots changes nothing."""
    from oracle import reference as R
    if not (R.available("reference") or R.available("port")):
        pytest.skip("no CPU oracle built")
    seen = set()
    S, Rw, x, tau, log = rollout_reference.reference_rollout(6, 60)
    for rec in log:
        assert np.all(rec["flag"] == 0)
        seen |= set(np.unique(rec["kind"]))
    assert seen == {0, 1, 2}
    assert np.all(np.isfinite(S)) and np.all(np.isfinite(tau)) and np.abs(tau).max() <= 60.0 + 1e-6
    assert np.abs(Rw[:, rollout.RR_DX:rollout.RR_DV]).max() < 0.05
    S2, Rw2, x2, tau2, _ = rollout_reference.reference_rollout(3, 60, first=3)
    assert np.array_equal(S[3:], S2) and np.array_equal(x[3:], x2) and np.array_equal(tau[3:], tau2) and np.array_equal(Rw[3:], Rw2)
