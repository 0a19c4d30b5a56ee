"""Host check of the kernel generator (apf_quadruped_b200/codegen.py) -- no GPU needed.

The generator rewrites the throughput kernel's micro-programs (row blocks of LDL_numeric with Y in
registers, operands reused from registers, loads hoisted); its IR is executed here on the host and must
leave every element of the HBM slice exactly as the untransformed program does -- which
tests/test_stream_program.py in turn pins to the reference (ldl.c:253-326, 495-597; Auxilary.c:745-786).
"""
import numpy as np
import pytest

from apf_quadruped_b200 import codegen, dogbot
from test_stream_program import make_slice, run_stream


def same(a, b):
    return np.array_equal(a, b, equal_nan=True) and np.array_equal(np.signbit(a), np.signbit(b))


@pytest.mark.parametrize("options", [{}, {"share": 0}, {"ring": 8, "group": 3, "ahead": 2, "dist": 40}, {"rows": 2, "reuse": 16, "dist": 8}, {"rows": 0}, {"ybudget": 12, "fn_ops": 50, "chunk": 37, "rows": 4}])
@pytest.mark.parametrize("kind", ["trot", "stance", "crawl", "mpc"])
def test_ir_leaves_the_slice_as_the_program_does(qps, kind, options):
    batch = dogbot.generate(kind, 2, start=4242)
    pat = batch.pattern
    plan = qps.Plan(pat)
    info = plan.stream_info()
    segs, opt = codegen.build_all(plan, **options)
    rng = np.random.default_rng(9)
    n, m, p, N = pat.n, pat.m, pat.p, pat.N
    for i in range(batch.count):
        G, lay, S = make_slice(info, batch, i)
        G[lay["zd"]:lay["zd"] + m] = -rng.uniform(0.05, 20.0, m)  # -s/z of some iterate
        G2, S2 = G.copy(), S.copy()
        run_stream(plan.stream_words("factor"), info["group"], G, S, False)
        codegen.evaluate(segs["factor"], G2, S2)
        assert same(G, G2)
        rhs = rng.normal(size=N)
        for g, s in ((G, S), (G2, S2)):
            g[lay["rhs"]:lay["rhs"] + N] = rhs
            s[:] = np.nan
        run_stream(plan.stream_words("solve"), info["group"], G, S, False)
        codegen.evaluate(segs["solve"], G2, S2)
        assert same(G, G2) and same(S, S2)
        x, y, z, s_ = rng.normal(size=n), rng.normal(size=p), rng.uniform(0.1, 2, m), rng.uniform(0.1, 2, m)
        for g, s in ((G, S), (G2, S2)):
            for name, v in (("x", x), ("y", y), ("z", z), ("s", s_)):
                g[lay[name]:lay[name] + v.size] = v
            g[lay["hs"]:lay["hs"] + m] = s_ - batch.h[i]
            g[lay["nb"]:lay["nb"] + p] = -batch.b[i]
            s[:] = np.nan
            s[:n + p + m] = np.concatenate([x, y, z])
        run_stream(plan.stream_words("resid"), info["group"], G, S, True)
        codegen.evaluate(segs["resid"], G2, S2)
        assert same(G, G2)
        for g in (G, G2):
            g[lay["hs"]:lay["hs"] + m] = -batch.h[i]
        run_stream(plan.stream_words("resz"), info["group"], G, S, True)
        codegen.evaluate(segs["resz"], G2, S2)
        assert same(G, G2)
    plan.close()


def test_row_blocks_share_operands_and_keep_y_in_registers(qps):
    """what the transformation is for: fewer HBM loads, no shared-memory traffic in the factorisation"""
    for kind, least in (("trot", 0.15), ("stance", 0.25), ("crawl", 0.25)):
        plan = qps.Plan(dogbot.generate(kind, 1).pattern)
        plain, _ = codegen.build_all(plan, rows=0, reuse=0)
        blocked, _ = codegen.build_all(plan)
        a, b = plain["factor"].stats, blocked["factor"].stats
        assert a["ops"] == b["ops"] and a["stg"] == b["stg"]
        assert b["hbm_loads"] <= (1.0 - least) * a["hbm_loads"], (kind, a, b)
        assert b["lds"] + b["sts"] < 0.6 * (a["lds"] + a["sts"])
        for seg in ("solve", "resid", "resz"):
            assert plain[seg].stats["ops"] == blocked[seg].stats["ops"]
        plan.close()


def test_loads_never_pass_a_store_to_their_element(qps):
    plan = qps.Plan(dogbot.generate("trot", 1).pattern)
    segs, _ = codegen.build_all(plan)
    for B in segs.values():
        for fn in B.functions:
            stored = {}
            for idx, st in enumerate(codegen.statements(fn)):
                if st[0] == "stg":
                    stored[st[1]] = idx
            for idx, st in enumerate(codegen.statements(fn)):
                if st[0] == "ldg" and st[2] in stored:
                    # a load of an element this function stores comes after the store, or before it in the
                    # program (write after read) -- the evaluation test above decides which was meant
                    assert idx != stored[st[2]]
    plan.close()
