"""SURVEY 8 f4 on the device: the rollout loop (tick kernel -> formulation -> indexed solve -> torques, per contact
pattern, replayed from a CUDA graph) against the CPU oracle stepping the same model in numpy, bit for bit."""
import numpy as np
import pytest

from apf_quadruped_b200 import rollout
from oracle import rollout_reference

pytestmark = pytest.mark.gpu


def _same(snap, S, Rw, x, tau):
    return np.array_equal(snap["S"], S) and np.array_equal(snap["Rw"], Rw) and np.array_equal(snap["x"], x) and np.array_equal(snap["tau"], tau)


def test_device_rollout_matches_the_cpu_oracle_tick_by_tick():
    """64 robots x 50 ticks (cold-start QPs): after every one of the first ticks and at the end, state rows, rollout rows,
    solutions, torques, contact kinds, exit flags and iteration counts are the CPU's."""
    import torch
    ro = rollout.Rollout(torch, 64, device=0, graph=False)
    seen = {}

    def record(t, S, Rw, x, tau):
        if t < 3 or t == 26:
            seen[t] = (S.copy(), Rw.copy(), x.copy(), tau.copy())

    S, Rw, x, tau, log = rollout_reference.reference_rollout(64, 50, record=record)
    for t in range(50):
        ro.step(1)
        if t in seen:
            snap = ro.snapshot()
            assert snap["tick"] == t + 1
            assert np.array_equal(snap["kind"], log[t]["kind"])
            assert np.array_equal(snap["flag"], log[t]["flag"]) and np.array_equal(snap["iters"], log[t]["iters"])
            assert _same(snap, *seen[t]), f"tick {t} differs from the CPU oracle"
    snap = ro.snapshot()
    assert _same(snap, S, Rw, x, tau)
    assert np.all(snap["flag"] == 0)
    ro.close()


def test_graph_replay_equals_eager_and_split_equals_whole():
    import torch
    eager = rollout.Rollout(torch, 96, device=0, graph=False)
    eager.step(60)
    a = eager.snapshot()
    eager.close()
    g = rollout.Rollout(torch, 96, device=0, graph=True)
    g.step(60)
    b = g.snapshot()
    g.close()
    for k in ("S", "Rw", "x", "tau", "flag", "iters", "kind"):
        assert np.array_equal(a[k], b[k]), f"CUDA-graph replay differs from eager launches in {k}"
    # the robots of a second GPU: first = 48 (nothing crosses GPUs during the ticks)
    half = rollout.Rollout(torch, 48, device=0, first=48, graph=True)
    half.step(60)
    c = half.snapshot()
    half.close()
    for k in ("S", "Rw", "x", "tau", "flag", "iters", "kind"):
        assert np.array_equal(a[k][48:], c[k]), f"a split rollout differs from the whole in {k}"
