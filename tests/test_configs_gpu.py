"""GPU parity on the shapes of BASELINE.json's other configs (they are parity cases, not bench lines).

configs[2]  mixed contact-set batch: stance / crawl / trot QPs, one symbolic plan per pattern, each
            pattern's sub-batch split across ranks (shard.mixed_shards) -- here 2 emulated ranks.
configs[3]  Monte-Carlo rollout: robots x sequential control ticks, tick t+1's QP built from tick t's
            solution.  "Warm start" has no reference semantics (kkt_initialize always cold-starts,
            Auxilary.c:992-1089), so every tick is a cold-start solve and must match the oracle bit
            for bit -- which also makes the two trajectories identical tick after tick.
"""
import numpy as np
import pytest

from apf_quadruped_b200 import dogbot, shard
from test_parity_gpu import check_against_oracle

pytestmark = pytest.mark.gpu


def test_mixed_contact_set_batch_sharded(qps, oracle, monkeypatch):
    monkeypatch.setenv("QPB_STREAM_MIN", "1")  # the throughput kernel, as at full size
    counts = {"stance": 150, "crawl": 130, "trot": 170}
    world = 2
    solvers = {k: qps.BatchSolver(dogbot.generate(k, 1).pattern, counts[k]) for k in counts}
    got = {k: [] for k in counts}
    for rank in range(world):
        for kind, (lo, hi) in shard.mixed_shards(counts, world, rank).items():
            batch = dogbot.generate(kind, hi - lo, start=lo)
            got[kind].append(solvers[kind].solve(batch))
    for kind, parts in got.items():
        whole = {f: np.concatenate([p[f] for p in parts]) for f in ("x", "y", "z", "s", "flag", "iters")}
        ref, _ = oracle.run_batch(dogbot.generate(kind, counts[kind]))
        check_against_oracle(whole, ref)
        solvers[kind].close()


def _next_tick(batch, x):
    """A fixed 'integrator': the next tick's linear cost and bounds react to this tick's solution."""
    c = batch.c + 0.05 * np.tanh(x)
    h = batch.h + 0.01 * np.abs(np.sin(x[:, :1]))
    return dogbot.Batch(batch.pattern, batch.Ppr, batch.Apr, batch.Gpr, np.ascontiguousarray(c), np.ascontiguousarray(h), batch.b)


@pytest.mark.parametrize("kernel_min", ["1", "1000000"])  # throughput kernel / latency kernel
def test_rollout_of_sequential_ticks(qps, oracle, monkeypatch, kernel_min):
    monkeypatch.setenv("QPB_STREAM_MIN", kernel_min)
    robots, ticks = 96, 4
    dev = dogbot.generate("stance", robots, start=8000)
    cpu = dogbot.generate("stance", robots, start=8000)
    solver = qps.BatchSolver(dev.pattern, robots)
    for _ in range(ticks):
        out = solver.solve(dev)
        ref, _ = oracle.run_batch(cpu)
        check_against_oracle(out, ref)
        dev, cpu = _next_tick(dev, out["x"]), _next_tick(cpu, ref["x"])
        assert np.array_equal(dev.c, cpu.c) and np.array_equal(dev.h, cpu.h)
    solver.close()


def test_edge_counts_and_errors(qps, oracle, monkeypatch):
    monkeypatch.setenv("QPB_STREAM_MIN", "1")
    batch = dogbot.generate("trot", 33)
    solver = qps.BatchSolver(batch.pattern, 33)
    one = batch.slice(0, 1)
    out = solver.solve(one)  # a single QP through the throughput kernel: 31 idle lanes
    check_against_oracle(out, oracle.solve_batch(one))
    out = solver.solve(batch)  # one full warp and one lane
    check_against_oracle(out, oracle.solve_batch(batch))
    big = dogbot.generate("trot", 34)
    with pytest.raises(RuntimeError):  # count > capacity: QPB_E_ARG, nothing solved
        solver.solve(big)
    solver.close()
