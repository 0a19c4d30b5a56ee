"""Host check of the throughput kernel's micro-programs (csrc/stream_program.cpp) -- no GPU needed.

A small interpreter executes the micro-ops exactly as one thread of the kernel does (accumulators t
and d in registers, prefetched HBM operands read at the START of their group, late and shared
operands at execution time) and the results are compared bit for bit with the reference's
LDL_numeric / LDL_lsolve / dsolve / ltsolve (ldl.c:253-326, 495-597) and computeresiduals
(Auxilary.c:745-786), obtained from the compiled reference through oracle.reference.
"""
import numpy as np
import pytest

from apf_quadruped_b200 import dogbot

TERM, DIVACC, DIV, INITD, ENDD, LOADS, NOP = 0, 2, 3, 4, 5, 6, 8
FIRSTZ, INIT, ST_SHARED, ST_HBM, LATE1, ONE, SLOW = (1 << (52 + i) for i in range(7))


def regularise(d):
    """ldl.c:319-320"""
    sign = -1.0 if d <= 0.0 else 1.0
    return sign * 0.0000001 if sign * d <= 0.00000000000001 else d


def decode(w):
    w = int(w)
    return dict(f1=w & 0xffff, f2=(w >> 16) & 0xffff, c=(w >> 32) & 0xffff, op=(w >> 48) & 15, w=w)


def run_stream(words, group, G, S, gs):
    """Execute micro-ops on the HBM slice G and the shared window S (float64 arrays), in place, with
    the kernel's software pipeline: the HBM operands of group i+depth-1 are read BEFORE group i executes.
    gs: residual program (the window holds x | y | z, results go to HBM) instead of factor / solve."""
    assert len(words) % group == 0
    f = np.float64
    t = d = f(np.nan)
    groups = [[decode(w) for w in words[g0:g0 + group]] for g0 in range(0, len(words), group)]

    def request(ins):
        return [f(G[i["f1"]]) for i in ins]

    depth = 32 // group  # STREAM_FS_DEPTH
    assert len(groups) % depth == 0, "the kernel's pipeline turns `depth` groups at a time"
    with np.errstate(all="ignore"):
        inflight = [request(g) for g in groups[:depth - 1]]
        for gi, ins in enumerate(groups):
            if gi + depth - 1 < len(groups):
                inflight.append(request(groups[gi + depth - 1]))
            p1 = inflight.pop(0)
            slow = bool(ins[0]["w"] & SLOW)
            if not slow:  # fast group: every second operand is read first, every result stored last
                early = [f(S[i["f2"]]) for i in ins]
            pending = []
            for u, i in enumerate(ins):
                w, op = i["w"], i["op"]
                assert bool(w & SLOW) == slow, "a group is fast or slow as a whole"
                if not slow:  # the branch-free path knows four flags and nothing else
                    assert op == TERM and not w & LATE1 and not w & (ST_SHARED if gs else ST_HBM)
                P = f(G[i["f1"]]) if w & LATE1 else p1[u]
                if gs:
                    assert op in (TERM, NOP) and not w & ST_SHARED
                if w & ONE:
                    Sv = f(1.0)
                elif not slow:
                    Sv = early[u]
                else:
                    Sv = f(S[i["f2"]]) if op in (TERM, DIVACC, DIV, LOADS) else None
                if op == TERM:
                    r = (f(0.0) if w & FIRSTZ else t) - P * Sv
                    t = P if w & INIT else r
                elif op == DIV:
                    t = Sv / P
                elif op == INITD:
                    d = P
                elif op == DIVACC:
                    t = Sv / P
                    d = d - t * Sv
                elif op == ENDD:
                    t = f(regularise(float(d)))
                elif op == LOADS:
                    t = Sv
                else:
                    assert op == NOP
                if slow:
                    if w & ST_HBM:
                        G[i["c"]] = t
                    if w & ST_SHARED:
                        S[i["c"]] = t
                else:
                    pending.append((w, i["c"], t))
            for w, c, v in pending:
                if w & ST_HBM:
                    G[c] = v
                if w & ST_SHARED:
                    S[c] = v
    return G


def make_slice(info, batch, i):
    lay = info["layout"]
    pat = batch.pattern
    G = np.full(lay["total"], np.nan)
    G[lay["zero"]] = 0.0
    G[lay["one"]] = 1.0
    G[lay["mone"]] = -1.0
    vals = np.concatenate([batch.Ppr[i], batch.Apr[i], batch.Gpr[i]]) + 0.0
    G[lay["V"]:lay["V"] + vals.size] = vals
    G[lay["c"]:lay["c"] + pat.n] = batch.c[i]
    G[lay["h"]:lay["h"] + pat.m] = batch.h[i]
    G[lay["b"]:lay["b"] + pat.p] = batch.b[i]
    G[lay["zd"]:lay["zd"] + pat.m] = -1.0
    return G, lay, np.full(lay["scratch"], np.nan)


@pytest.mark.parametrize("kind", ["trot", "stance", "crawl", "mpc"])
def test_stream_programs_reproduce_reference_factor_and_solve(qps, oracle, kind):
    batch = dogbot.generate(kind, 2, start=321)
    pat = batch.pattern
    plan = qps.Plan(pat)
    info = plan.stream_info()
    assert info["ok"] == 1
    N, lnz = plan.info("N"), plan.info("lnz")
    for i in range(batch.count):
        G, lay, S = make_slice(info, batch, i)
        ref = oracle.solve_sparse(pat, batch.Ppr[i], batch.Apr[i], batch.Gpr[i], batch.c[i], batch.h[i], batch.b[i],
                                  maxit=0, want_factor=True)
        run_stream(plan.stream_words("factor"), info["group"], G, S, False)
        assert np.array_equal(G[lay["Lx"]:lay["Lx"] + lnz], ref.init["Lx"])
        assert np.array_equal(G[lay["D"]:lay["D"] + N], ref.init["D"])
        G[lay["rhs"]:lay["rhs"] + N] = np.concatenate([-batch.c[i], batch.b[i], batch.h[i]])
        S[:] = np.nan
        run_stream(plan.stream_words("solve"), info["group"], G, S, False)
        sol = G[lay["rhs"]:lay["rhs"] + N]
        assert np.array_equal(sol[:pat.n], ref.init["x"])
        assert np.array_equal(sol[pat.n:pat.n + pat.p], ref.init["y"])
        assert not np.isnan(sol).any()


def _csc_dense(jc, ir, pr, nrows, ncols):
    A = np.zeros((nrows, ncols))
    for c in range(ncols):
        for q in range(jc[c], jc[c + 1]):
            A[ir[q], c] = pr[q]
    return A


@pytest.mark.parametrize("kind", ["trot", "stance", "mpc"])
def test_stream_residual_programs(qps, kind):
    """rx, ry, rz and -(P x) in the reference's accumulation order (Auxilary.c:745-786, 802-860)."""
    batch = dogbot.generate(kind, 1, start=77)
    pat = batch.pattern
    plan = qps.Plan(pat)
    info = plan.stream_info()
    G, lay, S = make_slice(info, batch, 0)
    rng = np.random.default_rng(5)
    n, m, p = pat.n, pat.m, pat.p
    x, y, z, s = rng.normal(size=n), rng.normal(size=p), rng.uniform(0.1, 2, m), rng.uniform(0.1, 2, m)
    for name, v in (("x", x), ("y", y), ("z", z), ("s", s)):
        G[lay[name]:lay[name] + v.size] = v
    c, h, b = batch.c[0], batch.h[0], batch.b[0]
    G[lay["hs"]:lay["hs"] + m] = s - h
    G[lay["nb"]:lay["nb"] + p] = -b
    S[:n + p + m] = np.concatenate([x, y, z])  # the kernel copies the iterate into the window first
    run_stream(plan.stream_words("resid"), info["group"], G, S, True)
    # the reference's loops, literally
    Ppr, Apr, Gpr = batch.Ppr[0] + 0.0, batch.Apr[0] + 0.0, batch.Gpr[0] + 0.0
    rx = np.zeros(n)
    for col in range(n):  # y -= P x, column scatter
        for q in range(pat.Pjc[col], pat.Pjc[col + 1]):
            rx[pat.Pir[q]] -= x[col] * Ppr[q]
    px = rx.copy()
    for col in range(n):  # y -= G'z, column gather
        for q in range(pat.Gjc[col], pat.Gjc[col + 1]):
            rx[col] -= Gpr[q] * z[pat.Gir[q]]
    if p:
        for col in range(n):
            for q in range(pat.Ajc[col], pat.Ajc[col + 1]):
                rx[col] -= Apr[q] * y[pat.Air[q]]
    rx += c * -1.0
    ry = np.zeros(p)
    if p:
        for col in range(n):
            for q in range(pat.Ajc[col], pat.Ajc[col + 1]):
                ry[pat.Air[q]] -= x[col] * Apr[q]
        ry += b * 1.0
    rz = np.zeros(m)
    for col in range(n):
        for q in range(pat.Gjc[col], pat.Gjc[col + 1]):
            rz[pat.Gir[q]] -= x[col] * Gpr[q]
    rz0 = rz.copy()
    rz += h - s
    r = G[lay["r"]:lay["r"] + n + p + m]
    assert np.array_equal(r[:n], rx) and np.array_equal(r[n:n + p], ry) and np.array_equal(r[n + p:], rz)
    assert np.array_equal(G[lay["px"]:lay["px"] + n], px)
    # RESZ: hs <- -(G x) + hs in place (the initial point's r = h - G x with hs = h)
    G[lay["hs"]:lay["hs"] + m] = -h
    run_stream(plan.stream_words("resz"), info["group"], G, S, True)
    assert np.array_equal(G[lay["hs"]:lay["hs"] + m], rz0 + h)


def test_stream_program_statistics(qps):
    for kind in ("trot", "stance", "crawl", "mpc"):
        pat = dogbot.generate(kind, 1).pattern
        plan = qps.Plan(pat)
        info = plan.stream_info()
        lnz, N = plan.info("lnz"), plan.info("N")
        f, s = info["segments"]["factor"], info["segments"]["solve"]
        assert f["divs"] == lnz and s["divs"] == N and s["terms"] == 2 * lnz
        # one multiply-subtract per update of the reference's row solves; the pivots' l*y terms are in DIVACC
        li, lp = plan.array("Li"), plan.array("Lp")
        updates = sum(int(np.sum(li[lp[j]:lp[j + 1]] > r)) for j in range(N) for r in li[lp[j]:lp[j + 1]])
        assert f["terms"] == updates
        assert info["layout"]["scratch"] >= N and info["layout"]["total"] <= 65535
        for seg in info["segments"].values():
            assert seg["ops"] == seg["groups"] * info["group"]


def test_library_carries_specialised_kernels_for_the_wbc_patterns(qps):
    """`make` generates straight-line kernels from the micro-programs of the three WBC patterns
    (apf_quadruped_b200/codegen.py) and compiles them in; each registers under its program's key."""
    import ctypes
    keys = (ctypes.c_uint64 * 16)()
    cnt = int(qps.lib().QPB_GENERATED_KERNELS(keys, 16))
    have = {int(keys[i]) for i in range(min(cnt, 16))}
    for kind in ("trot", "stance", "crawl"):
        plan = qps.Plan(dogbot.generate(kind, 1).pattern)
        assert plan.stream_key() in have, f"no specialised kernel for {kind}: run `make`"
        plan.close()
    mpc = qps.Plan(dogbot.generate("mpc", 1).pattern)
    assert mpc.stream_key() not in have  # runs on the interpreter / the latency kernel
    mpc.close()

