"""Host check of the compiled factor / solve schedules (csrc/program.cpp) -- no GPU needed.

A small interpreter executes the instruction words exactly as the kernel's warp does (32 lanes, all
loads of a wave before its stores, accumulators private to a lane) and the results are compared bit
for bit with the reference's LDL_numeric / LDL_lsolve / dsolve / ltsolve (ldl.c:253-326, 495-597),
obtained from the compiled reference through oracle.reference.  This pins the scheduler: operation
order per value, operand readiness, scratch allocation and the in-place solution.
"""
import numpy as np
import pytest

from apf_quadruped_b200 import dogbot

FIRST, LAST, NEGZ, ADD, ACTIVE = 1 << 14, 1 << 15, 1 << 30, 1 << 31, 1 << 47
SUMWAVE, DIVWAVE = 1 << 62, 1 << 63
MASK = (1 << 14) - 1


def regularise(d):
    """ldl.c:319-320"""
    sign = -1.0 if d <= 0.0 else 1.0
    return sign * 0.0000001 if sign * d <= 0.00000000000001 else d


class LaneLoads:
    def __init__(self):
        self.by_addr, self.lane = {}, 0

    def add(self, addr):
        self.by_addr.setdefault(addr, set()).add(self.lane)

    def update(self, addrs):
        for a in addrs:
            self.add(a)


def run_segment(words, S, check_hazards=True):
    """Execute super-waves `words` (shape (w, 32, width)) on the slice S (float64 array) in place.
    A super-wave is `width` instruction words per lane, executed between two __syncwarp()s."""
    words = words.reshape(words.shape[0], -1)
    nl = words.shape[1]
    acc = [0.0] * nl
    for wave in words:
        w0 = int(wave[0])
        stores = []
        loads = LaneLoads()
        if w0 & DIVWAVE:
            for lane in range(nl):
                w = int(wave[lane])
                assert w & DIVWAVE
                if w & ACTIVE:
                    a, b, cs = w & MASK, (w >> 32) & MASK, (w >> 48) & MASK
                    loads.lane = lane
                    loads.update((a, b))
                    stores.append((cs, float(np.float64(S[a]) / np.float64(regularise(float(S[b])))), lane))
        elif w0 & SUMWAVE:
            for lane in range(nl):
                w = int(wave[lane])
                assert w & SUMWAVE
                if w & ACTIVE:
                    a, ci, cnt, cs = w & MASK, (w >> 16) & MASK, (w >> 32) & MASK, (w >> 48) & MASK
                    t = float(S[ci])
                    loads.lane = lane
                    loads.add(ci)
                    for i in range(cnt):
                        t = t - float(S[a + i])
                        loads.add(a + i)
                    stores.append((cs, t, lane))
        else:
            for lane in range(nl):
                w = int(wave[lane])
                assert not (w & (DIVWAVE | SUMWAVE))
                a, ci, b, cs = w & MASK, (w >> 16) & MASK, (w >> 32) & MASK, (w >> 48) & MASK
                t = acc[lane]
                loads.lane = lane
                if w & FIRST:
                    t = float(np.float64(S[ci]) + np.float64(-0.0 if w & NEGZ else 0.0))
                    loads.add(ci)
                prod = float(np.float64(S[a]) * np.float64(S[b]))
                loads.update((a, b))
                t = t + prod if w & ADD else t - prod
                acc[lane] = t
                if w & LAST:
                    stores.append((cs, t, lane))
        if check_hazards:
            targets = [cs for cs, _, _ in stores]
            assert len(set(targets)) == len(targets), "two lanes store to the same double in one wave"
            for cs, _, lane in stores:  # nothing may read, in the same wave, what another lane stores
                assert loads.by_addr.get(cs, {lane}) <= {lane}, "load and store of one double in the same wave"
        for cs, v, _ in stores:
            S[cs] = v
    return S


def make_slice(plan, batch, i):
    lay = plan.layout()
    pat = batch.pattern
    S = np.full(lay["total"], np.nan)
    S[lay["zero"]] = 0.0
    S[lay["zero"] + 1] = -0.0
    vals = np.concatenate([batch.Ppr[i], batch.Apr[i], batch.Gpr[i]])
    S[lay["V"]:lay["V"] + vals.size] = vals
    S[lay["c"]:lay["c"] + pat.n] = batch.c[i]
    S[lay["h"]:lay["h"] + pat.m] = batch.h[i]
    S[lay["b"]:lay["b"] + pat.p] = batch.b[i]
    S[lay["zd"]:lay["zd"] + pat.m] = -1.0
    return S, lay


@pytest.mark.parametrize("kind", ["trot", "stance", "crawl", "mpc"])
def test_schedules_reproduce_reference_factor_and_solve(qps, oracle, kind):
    batch = dogbot.generate(kind, 2, start=123)
    pat = batch.pattern
    plan = qps.Plan(pat)
    assert plan.info("program_ok") == 1
    fs, ss = plan.segment("factor"), plan.segment("solve")
    N, lnz = plan.info("N"), plan.info("lnz")
    for i in range(batch.count):
        S, lay = make_slice(plan, batch, i)
        ref = oracle.solve_sparse(pat, batch.Ppr[i], batch.Apr[i], batch.Gpr[i], batch.c[i], batch.h[i], batch.b[i],
                                  maxit=0, want_factor=True)
        run_segment(plan.program("factor"), S)
        assert np.array_equal(S[lay["Lx"]:lay["Lx"] + lnz], ref.init["Lx"])
        # the chains store raw pivots; ldl.c:319-320 is applied wherever a pivot is used as a divisor
        assert np.array_equal(np.array([regularise(float(d)) for d in S[lay["D"]:lay["D"] + N]]), ref.init["D"])
        rhs = np.concatenate([-batch.c[i], batch.b[i], batch.h[i]])
        S[lay["rhs"]:lay["rhs"] + N] = rhs
        run_segment(plan.program("solve"), S)
        sol = S[lay["rhs"]:lay["rhs"] + N]
        assert np.array_equal(sol[:pat.n], ref.init["x"])
        assert np.array_equal(sol[pat.n:pat.n + pat.p], ref.init["y"])
        assert not np.isnan(sol).any()


def test_schedule_statistics_are_consistent(qps):
    """Work is conserved: every multiply-subtract and quotient of the reference appears exactly once."""
    for kind in ("trot", "stance", "mpc"):
        pat = dogbot.generate(kind, 1).pattern
        plan = qps.Plan(pat)
        fs, ss = plan.segment("factor"), plan.segment("solve")
        lnz, N = plan.info("lnz"), plan.info("N")
        lp = plan.array("Lp")
        li = plan.array("Li")
        # LDL_numeric: sum over rows k, over pattern columns i, of the entries stored in column i before row k
        cnt = np.diff(lp)
        updates = sum(int(np.sum(li[lp[j]:lp[j + 1]] > r)) for j in range(N) for r in li[lp[j]:lp[j + 1]])
        assert fs["divs"] == lnz
        assert fs["terms"] == updates + lnz  # triangular-solve updates + one l*y term per entry for the pivots
        assert ss["divs"] == N
        assert ss["terms"] >= 2 * lnz
        assert fs["waves"] % 3 == 0 and ss["waves"] % 3 == 0 and fs["width"] in (1, 2, 4) and ss["width"] in (1, 2, 4)
        lay = plan.layout()
        assert lay["total"] == plan.info("slice_doubles") <= 16384
        assert all(v % 2 == 0 for k, v in lay.items() if k not in ("total", "tsize")) and cnt.sum() == lnz
