"""GPU parity of the THROUGHPUT kernel (one QP per thread, csrc/stream.cuh) versus the CPU oracle.

Same bar as tests/test_parity_gpu.py (which exercises the latency kernel with its small batches):
identical exit flags and iteration counts, bit-identical x, y, z, s (the 1e-8 relative tolerance of
BASELINE.json's north_star is the documented floor).  QPB_STREAM_MIN=1 sends every batch, however
small, to the throughput kernel; the default threshold is covered by the last test.
"""
import os

import numpy as np
import pytest

from apf_quadruped_b200 import dogbot
from test_parity_gpu import check_against_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture()
def stream_only(monkeypatch):
    monkeypatch.setenv("QPB_STREAM_MIN", "1")
    monkeypatch.setenv("QPB_STREAM_MIN_WARPS", "1")  # also patterns that would default to the latency kernel (mpc)


GENERATED_KINDS = ("trot", "stance", "crawl")  # Makefile: GEN_KINDS


@pytest.mark.parametrize("program", ["generated", "interpreted"])
@pytest.mark.parametrize("kind,count", [("stance", 300), ("crawl", 97), ("trot", 300), ("mpc", 70)])
def test_stream_batch_matches_oracle(qps, oracle, stream_only, monkeypatch, kind, count, program):
    """ragged counts: partial warps, lanes that run out of work at different times"""
    if program == "interpreted":
        monkeypatch.setenv("QPB_STREAM_INTERPRET", "1")
    batch = dogbot.generate(kind, count, start=40000)
    solver = qps.BatchSolver(batch.pattern, count)
    out = solver.solve(batch)
    assert solver.last_kernel == "stream"
    assert solver.last_kernel_generated == (program == "generated" and kind in GENERATED_KINDS)
    ref = oracle.solve_batch(batch)
    check_against_oracle(out, ref)
    fval, stat = solver.read_stats(count)
    assert np.array_equal(fval, ref["fval"])
    solver.close()


def test_stream_controller_tolerances_and_maxit(qps, oracle, stream_only):
    batch = dogbot.generate("trot", 130, start=5000)
    solver = qps.BatchSolver(batch.pattern, 130)
    out = solver.solve(batch, qps.default_settings(reltol=1e-2, abstol=1e-2))
    check_against_oracle(out, oracle.solve_batch(batch, reltol=1e-2, abstol=1e-2))
    out = solver.solve(batch, qps.default_settings(maxit=0))
    assert np.all(out["flag"] == qps.QP_MAXIT) and np.all(out["iters"] == 0)
    check_against_oracle(out, oracle.solve_batch(batch, maxit=0))
    out = solver.solve(batch, qps.default_settings(maxit=3))
    check_against_oracle(out, oracle.solve_batch(batch, maxit=3))
    assert solver.last_kernel == "stream"
    solver.close()


def test_stream_user_permutation_and_sigma_d(qps, oracle, stream_only):
    """sigma_d > 0 reaches the centering-only branch (qpSWIFT.c:572-579)"""
    batch = dogbot.generate("crawl", 64, start=900)
    pat = batch.pattern
    perm = np.random.default_rng(5).permutation(pat.N)
    solver = qps.BatchSolver(pat, 64, sigma_d=0.05, permut=perm)
    out = solver.solve(batch)
    assert solver.last_kernel == "stream"
    check_against_oracle(out, oracle.solve_batch(batch, sigma_d=0.05, permut=perm))
    solver.close()


def test_generated_and_interpreted_programs_agree_on_a_full_device(qps, stream_only, monkeypatch):
    """40,000 trot QPs: every warp of every SM busy, lanes refilled from the queue; the specialised
    kernel and the interpreter must agree bit for bit (the oracle checks smaller batches above)."""
    batch = dogbot.generate("trot", 40000, start=123456)
    solver = qps.BatchSolver(batch.pattern, batch.count)
    out = solver.solve(batch)
    assert solver.last_kernel_generated
    solver.close()
    monkeypatch.setenv("QPB_STREAM_INTERPRET", "1")
    solver = qps.BatchSolver(batch.pattern, batch.count)
    out2 = solver.solve(batch)
    assert solver.last_kernel == "stream" and not solver.last_kernel_generated
    for k in ("x", "y", "z", "s", "flag", "iters"):
        assert np.array_equal(out2[k], out[k]), k
    solver.close()


def test_stream_and_latency_kernels_agree_at_default_threshold(qps, oracle, monkeypatch):
    """4,096 trot QPs on the throughput kernel (its threshold for patterns without a specialised warp-per-QP
    kernel); 512 of them on the warp-per-QP kernel; both are bit-identical with the oracle, so results do not
    depend on how a batch is cut."""
    monkeypatch.setenv("QPB_STREAM_MIN", "2048")
    batch = dogbot.generate("trot", 4096, start=70000)
    solver = qps.BatchSolver(batch.pattern, 4096)
    out = solver.solve(batch)
    assert solver.last_kernel == "stream" and solver.last_kernel_generated
    ref, _ = oracle.run_batch(batch)
    check_against_oracle(out, ref)
    small = dogbot.generate("trot", 512, start=70000)
    out2 = solver.solve(small)
    assert solver.last_kernel == "warp"
    for k in ("x", "y", "z", "s", "flag", "iters"):
        assert np.array_equal(out2[k], out[k][:512])
    solver.close()


@pytest.mark.parametrize("spill", ["32", "0", "5"])
def test_stragglers_are_handed_to_the_latency_kernel(qps, oracle, stream_only, monkeypatch, spill):
    """trot instance 264420 runs into maxit = 100; the throughput kernel hands QPs that exceed
    QPB_STREAM_SPILL iterations to the latency kernel (second launch), which must not change a bit.
    spill = 5 sends most of the batch through the hand-over, 0 disables it."""
    monkeypatch.setenv("QPB_STREAM_SPILL", spill)
    batch = dogbot.generate("trot", 256, start=264320)
    solver = qps.BatchSolver(batch.pattern, 256)
    out = solver.solve(batch)
    assert solver.last_kernel == "stream" and solver.last_launches == (1 if spill == "0" else 2)
    ref = oracle.solve_batch(batch)
    assert ref["flag"][100] == qps.QP_MAXIT and ref["iters"][100] == 100
    check_against_oracle(out, ref)
    fval, stat = solver.read_stats(256)
    assert np.array_equal(fval, ref["fval"])
    solver.close()


def test_results_written_straight_into_pinned_buffers(qps, oracle, stream_only):
    """QP_SOLVE_BATCH lets the throughput kernel write finished rows directly into page-locked caller buffers
    (no staging copy); pageable buffers take the staged path.  Same bits either way, stragglers included."""
    import torch
    batch = dogbot.generate("trot", 700, start=264000)  # contains the maxit instance 264420
    solver = qps.BatchSolver(batch.pattern, batch.count)
    plain = solver.solve(batch)
    pat = batch.pattern
    pinned = {"x": torch.empty((batch.count, pat.n), dtype=torch.float64).pin_memory(),
              "y": torch.empty((batch.count, pat.p), dtype=torch.float64).pin_memory(),
              "z": torch.empty((batch.count, pat.m), dtype=torch.float64).pin_memory(),
              "s": torch.empty((batch.count, pat.m), dtype=torch.float64).pin_memory()}
    out = {k: v.numpy() for k, v in pinned.items()}
    out["flag"] = np.empty(batch.count, dtype=np.int64)
    out["iters"] = np.empty(batch.count, dtype=np.int64)
    for v in pinned.values():
        v.fill_(float("nan"))
    solver.solve(batch, out=out)
    for k in ("x", "y", "z", "s", "flag", "iters"):
        assert np.array_equal(out[k], plain[k]), k
    assert out["iters"].max() == 100
    check_against_oracle(plain, oracle.solve_batch(batch))
    solver.close()
