"""Host check of the pattern-specialised warp-per-QP programs (apf_quadruped_b200/codegen_warp.py) -- no GPU.

The generator's IR is executed by its emulator (32 lanes in lock step, with the cross-lane hazard check: a
double written by one lane and read by another needs a __syncwarp in between) and compared bit for bit
with the reference's LDL_numeric, LDL solves (ldl.c:253-326, 495-597, through oracle.reference) and with a
plain restatement of computeresiduals' products (Auxilary.c:753-785)."""
import numpy as np
import pytest

from apf_quadruped_b200 import codegen_warp as W
from apf_quadruped_b200 import dogbot


def make_slice(lay, batch, i, st=None):
    pat = batch.pattern
    S = np.full(lay["total"], np.nan)
    S[lay["zero"]] = 0.0
    S[lay["zero"] + 1] = -0.0
    S[lay["vzero"]] = 0.0  # (the zero behind the value row, in the value row's address space)
    vals = np.concatenate([batch.Ppr[i], batch.Apr[i], batch.Gpr[i]])
    if st is not None:
        st.place_values(S, vals)  # (storage order, or residual order for a value row in the scratch row)
    else:
        S[lay["V"]:lay["V"] + vals.size] = vals + 0.0  # (the kernel normalises -0.0 to +0.0 when it loads the values)
    S[lay["c"]:lay["c"] + pat.n] = batch.c[i]
    S[lay["h"]:lay["h"] + pat.m] = batch.h[i]
    S[lay["b"]:lay["b"] + pat.p] = batch.b[i]
    S[lay["zd"]:lay["zd"] + pat.m] = -1.0
    return S


def residual_products(pat, Ppr, Apr, Gpr, c, h, b, x, y, z, s):
    """Auxilary.c:753-785: column scatter for P x, A x, G x; column gather for G'z, A'y."""
    def scatter(jc, ir, pr, xin, out):
        for col in range(len(jc) - 1):
            for q in range(jc[col], jc[col + 1]):
                out[ir[q]] -= xin[col] * pr[q]

    def gather(jc, ir, pr, xin, out):
        for col in range(len(jc) - 1):
            for q in range(jc[col], jc[col + 1]):
                out[col] -= pr[q] * xin[ir[q]]

    rx = np.zeros(pat.n)
    scatter(pat.Pjc, pat.Pir, Ppr, x, rx)
    px = rx.copy()
    gather(pat.Gjc, pat.Gir, Gpr, z, rx)
    if pat.p:
        gather(pat.Ajc, pat.Air, Apr, y, rx)
    rx += c * (-1.0)
    ry = np.zeros(pat.p)
    if pat.p:
        scatter(pat.Ajc, pat.Air, Apr, x, ry)
        ry += b * 1.0
    rz = np.zeros(pat.m)
    scatter(pat.Gjc, pat.Gir, Gpr, x, rz)
    rz += h - s
    return np.concatenate([rx, ry, rz]), px


@pytest.mark.parametrize("kind,vglobal,frows,pbsize", [("trot", True, 0, 64), ("stance", True, 40, 64), ("crawl", True, 200, 64), ("crawl", True, 0, 42), ("trot", False, 0, 64)])
def test_generated_warp_programs_reproduce_the_reference(qps, oracle, kind, vglobal, frows, pbsize):
    """vglobal: the value row P | A | G lies outside the shared-memory part of the slice (the warp's scratch row in global
    memory; loads hoisted far ahead of their use) -- the default; False: the whole working set in shared memory."""
    batch = dogbot.generate(kind, 2, start=321)
    pat = batch.pattern
    plan = qps.Plan(pat)
    # frows: the factorisation's value gathers with rows of their own behind the value row; pbsize: the backward sweep's
    # product buffer (42 doubles is what crawl needs -- and what lets an 11th slice fit, codegen_warp.generate)
    st = W.Structure(plan, vglobal, frows, pbsize)
    assert (st.lay["V"] >= st.lay["smem"]) == vglobal
    fac, sol, res = W.gen_factor(st), W.gen_solve(st), W.gen_resid(st)
    lay = st.lay
    assert len(st.frows) == (min(frows, len(st.frows)) if vglobal else 0) and (len(st.frows) > 0) == (vglobal and frows > 0)
    if vglobal:  # no statement mixes the two address spaces, and every value-row load is marked for the printer
        assert fac.vloads and res.vloads and not sol.vloads
    assert all(len(rows) <= 32 for rows in fac.passes) and sorted(k for rows in fac.passes for k in rows) == st.R
    N, lnz, n, m, p = plan.info("N"), plan.info("lnz"), pat.n, pat.m, pat.p
    for i in range(batch.count):
        S = make_slice(lay, batch, i, st)
        ref = oracle.solve_sparse(pat, batch.Ppr[i], batch.Apr[i], batch.Gpr[i], batch.c[i], batch.h[i], batch.b[i],
                                  maxit=0, want_factor=True)
        W.Emulator(fac, S).run()
        assert np.array_equal(S[lay["Lx"]:lay["Lx"] + lnz], ref.init["Lx"])
        assert np.array_equal(S[lay["D"]:lay["D"] + N], ref.init["D"])  # regularised pivots (ldl.c:319-320)
        S[lay["rhs"]:lay["rhs"] + N] = np.concatenate([-batch.c[i], batch.b[i], batch.h[i]])
        W.Emulator(sol, S).run()
        x = S[lay["rhs"]:lay["rhs"] + N]
        assert np.array_equal(x[:n], ref.init["x"]) and np.array_equal(x[n:n + p], ref.init["y"])
        assert not np.isnan(x).any()
        rng = np.random.default_rng(i)
        xv, yv, zv, sv = rng.standard_normal(n), rng.standard_normal(p), rng.random(m) + 0.1, rng.random(m) + 0.1
        for name, v in (("x", xv), ("y", yv), ("z", zv), ("s", sv)):
            S[lay[name]:lay[name] + v.size] = v
        W.Emulator(res, S).run()
        r, px = residual_products(pat, batch.Ppr[i], batch.Apr[i], batch.Gpr[i], batch.c[i], batch.h[i], batch.b[i], xv, yv, zv, sv)
        assert np.array_equal(S[lay["r"]:lay["r"] + N], r)
        assert np.array_equal(S[lay["T"]:lay["T"] + n], px)


def test_zero_right_hand_side_takes_the_exact_quotient(qps, oracle):
    """LDL_dsolve on exact zeros (the trot QP's b is all zero): x / D must stay +-0 with the reference's sign."""
    batch = dogbot.generate("trot", 1, start=5)
    plan = qps.Plan(batch.pattern)
    st = W.Structure(plan)
    fac, sol = W.gen_factor(st), W.gen_solve(st)
    lay, N = st.lay, plan.info("N")
    S = make_slice(lay, batch, 0, st)
    W.Emulator(fac, S).run()
    rhs = np.zeros(N)
    rhs[::3] = -0.0
    S[lay["rhs"]:lay["rhs"] + N] = rhs
    W.Emulator(sol, S).run()
    out = S[lay["rhs"]:lay["rhs"] + N]
    assert np.all(out == 0.0)


def test_emulator_flags_missing_syncs():
    ir = W.IR("t")
    v = ir.lds(None, 0, 1)
    ir.emit("sync")
    ir.emit("sts", None, 8, 1, v, None)   # lane l writes double l + 1 ...
    ir.lds(None, 0, 1)                    # ... which lane l + 1 reads without a sync
    with pytest.raises(AssertionError):
        W.Emulator(ir, np.zeros(64)).run()
    ir2 = W.IR("t")
    v = ir2.lds(None, 0, 1)
    ir2.emit("sync")                      # (the store below overwrites what the neighbour has just read)
    ir2.emit("sts", None, 8, 1, v, None)
    ir2.emit("sync")
    ir2.lds(None, 0, 1)
    W.Emulator(ir2, np.zeros(64)).run()


def test_layout_fits_more_warps_than_the_interpreted_one(qps):
    for kind in ("trot", "stance", "crawl"):
        plan = qps.Plan(dogbot.generate(kind, 1).pattern)
        st = W.Structure(plan)
        full = plan.layout()
        assert st.lay["smem"] < st.lay["total"] and st.lay["smem"] < full["total"] - plan.pattern.nnzG and st.lay["tsize"] >= 2 * plan.pattern.m
        assert st.lay["V"] >= st.lay["smem"], "the value row lies outside the shared-memory part of the slice"
        assert all(st.lay[k] % 2 == 0 for k in W.LAYOUT_FIELDS if k not in ("total", "tsize"))


def test_table_rows_are_read_two_per_load(qps, monkeypatch):
    """pack_tables: two `ldt` within the window become one `ldt2` over a 32-bit table whose halves are the two rows'
    16-bit offsets; the rows left single move to a compacted 16-bit table; the programs see the same offsets."""
    plan = qps.Plan(dogbot.generate("trot", 1).pattern)
    st = W.Structure(plan)
    monkeypatch.setenv("QPB_GEN_PACK", "0")
    plain = W.gen_solve(st)
    monkeypatch.setenv("QPB_GEN_PACK", "24")
    packed = W.gen_solve(st)
    assert not plain.tab32 and not any(s[0] == "ldt2" for s in plain.st)
    singles = [s for s in packed.st if s[0] == "ldt"]
    pairs = [s for s in packed.st if s[0] == "ldt2"]
    loads_before = sum(1 for s in plain.st if s[0] == "ldt")
    assert pairs and len(singles) + 2 * len(pairs) == loads_before and len(singles) + len(pairs) < 0.6 * loads_before
    assert len(packed.tab) == 32 * len(singles) and len(packed.tab32) == 32 * len(pairs)
    # every variable keeps its row of offsets
    rows_plain = {s[1]: tuple(plain.tab[s[2]:s[2] + 32]) for s in plain.st if s[0] == "ldt"}
    rows_packed = {s[1]: tuple(packed.tab[s[2]:s[2] + 32]) for s in singles}
    for s in pairs:
        words = packed.tab32[s[3]:s[3] + 32]
        rows_packed[s[1]] = tuple(w & 0xffff for w in words)
        rows_packed[s[2]] = tuple(w >> 16 for w in words)
    assert rows_packed == rows_plain
    # a variable is still defined before its first use
    seen = set()
    for s in packed.st:
        if s[0] == "ldt":
            seen.add(s[1])
        elif s[0] == "ldt2":
            seen.update(s[1:3])
        elif s[0] == "lds" and s[2] is not None:
            assert s[2] in seen


@pytest.mark.parametrize("kind", ["trot", "stance", "crawl"])
def test_value_row_in_residual_order(qps, kind):
    """Structure.vslots: every entry of P | A | G has one slot per product of computeresiduals that reads it (P: 1, A and
    G: 2), no slot holds two entries, a slot's lane and step are those of the chain that reads it, and the slot tables
    the kernel places a QP's values by (vdst, gslot in the generated source) say the same."""
    pat = dogbot.generate(kind, 1).pattern
    plan = qps.Plan(pat)
    st = W.Structure(plan, True, 24)  # (with a few factor-order rows behind the value row)
    W.gen_factor(st)
    nP, nA, nG = len(pat.Pir), len(pat.Air), len(pat.Gir)
    assert [len(v) for v in st.vslots] == [1] * nP + [2] * (nA + nG)
    flat = [s for v in st.value_slots() for s in v]
    assert len(flat) == len(set(flat)) and max(flat) < st.lay["total"] - st.lay["V"]
    # the chain that reads slot 32 T + lane is lane's term at step T
    T = 0
    for grp in st.resid_groups:
        for t in range(max(len(tl) for tl in grp)):
            for lane, tl in enumerate(grp):
                if t < len(tl):
                    assert 32 * T + lane in st.vslots[tl[t][0]]
            T += 1
    assert 32 * T == st.lay["vzero"] - st.lay["V"]
    # factor-order rows: 24 rows of 32 behind the zero that follows the value row
    assert len(st.frows) == 24 and st.F0 == st.lay["vzero"] + 2 and st.lay["total"] == st.F0 + 32 * 24
    # the generated source carries the same tables
    import re
    src, stats = W.generate(kind)
    st2 = W.Structure(plan, True, 0, st.lay["pbsize"] if stats["layout"]["tsize"] == st.lay["tsize"] else 42)
    vdst = [int(v) for v in re.search(r"unsigned short vdst\[\] = \{([^}]*)\}", src).group(1).split(",")]
    slots = st2.value_slots()
    k = max(len(v) for v in slots)
    assert vdst == [v[c] if c < len(v) else 0xffff for c in range(k) for v in slots] and f"kVDst = {k};" in src
    gslot = [int(v) for v in re.search(r"unsigned short gslot\[\] = \{([^}]*)\}", src).group(1).split(",")]
    rows = [[] for _ in range(pat.m)]
    for col in range(pat.n):
        for q in range(int(pat.Gjc[col]), int(pat.Gjc[col + 1])):
            rows[int(pat.Gir[q])].append(slots[nP + nA + q][0])
    assert gslot == [v for r in rows for v in r] and stats["layout"]["value_row_doubles"] == st2.lay["total"] - st2.lay["smem"]
