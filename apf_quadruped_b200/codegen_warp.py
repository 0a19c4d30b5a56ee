"""Pattern-specialised WARP-PER-QP kernels: LDL_numeric and the triangular solves of one sparsity pattern
as straight-line warp code over a factor that never leaves shared memory.

One warp owns one QP; the QP's whole working set (value row, factor, iterate: SliceLayout, csrc/program.h)
sits in the warp's slice of shared memory.  The generator maps the reference's loops onto the 32 lanes so
that every number is still produced by the reference's own sequence of IEEE-754 operations:

* LDL_numeric (ldl.c:253-326).  LANE = ROW of L.  The rows with a non-empty pattern are cut into passes
  of at most 32; in a pass every lane runs the up-looking triangular solve of ITS row, all lanes walking
  one common sequence of columns i (a linear extension of every row's Pattern order and of the
  elimination tree -- it exists for the DogBot patterns; the generator refuses a pattern without one).
  At step i a lane holds y = Y[i] of its row in a register, forms l = y / D[i], d -= l*y, stores L(k,i),
  and then applies column i to its work vector: Y[j] -= L(j,i)*y for the rows j of column i -- L(j,i) is
  ONE broadcast read of shared memory for all rows, Y[j] is a register with a static name.  A lane whose
  row does not hold column i carries y = +0.0 through the same instructions, which changes nothing
  (x - (+-0) = x for every x that is not -0.0, and no accumulator of the factorisation can be -0.0: each
  starts as 0 + v).  Columns that are leaves of the elimination tree (the inequality rows the ordering
  eliminates first) receive no updates: their pivots and their entries L(k,i) = K(i,k)/D[i] are computed
  beforehand, 32 at a time.
  "Y[i] += value" from Y[i] == 0 (ldl.c:284) is v + 0.0, i.e. v except that -0.0 becomes +0.0: the kernel
  normalises the matrix values that way once, when it loads them (the residual products cannot tell the
  difference: their accumulators are never -0.0 either).
* LDL_lsolve / dsolve / ltsolve (ldl.c:495-557).  Forward: lane = row r, w_r -= L(r,j)*w_j in ascending j
  (leaf columns: every lane walks its own row through per-lane tables; other columns: w_j comes by shuffle).
  Backward: the sum of a column is sequential by definition; its products are formed 32 at a time, the
  sum runs over broadcast reads.  Leaf columns are independent chains, one per lane.

The output is an IR (statements over 32-lane values) that is printed as CUDA and that `Emulator` executes
on the host with a cross-lane hazard check (a value written by one lane and read by another needs a
__syncwarp in between), so the generated code is checked against the reference bit for bit without a GPU
(tests/test_codegen_warp.py).

    python -m apf_quadruped_b200.codegen_warp --out build/gen trot stance crawl
"""
from __future__ import annotations

import argparse
import heapq
import os

import numpy as np

from . import codegen as CG

NLANE = 32


class IR:
    """Statement list + the table of 16-bit byte offsets the code indexes by lane."""

    def __init__(self, name):
        self.name = name
        self.st = []
        self.tab = []
        self.nd = self.ni = self.np_ = 0

    def d(self):
        self.nd += 1
        return f"d{self.nd}"

    def i(self):
        self.ni += 1
        return f"i{self.ni}"

    def p(self):
        self.np_ += 1
        return f"p{self.np_}"

    def table(self, entries):
        assert len(entries) <= NLANE
        off = len(self.tab)
        row = list(entries) + [0] * (NLANE - len(entries))
        assert all(0 <= e <= 0xffff for e in row), "byte offset does not fit 16 bits"
        self.tab.extend(row)
        return off

    def emit(self, *s):
        self.st.append(tuple(s))

    # ---- helpers; every address is a BYTE offset into the warp's slice -----------------------------
    def ldt(self, entries):
        v = self.i()
        self.emit("ldt", v, self.table(entries))
        return v

    def lds(self, ivar, imm, lc=0, care=NLANE):
        """care: lanes [care, 32) never use what they load (the emulator poisons it and skips their hazard check)."""
        v = self.d()
        self.emit("lds", v, ivar, imm, lc, care)
        return v

    def gather(self, offsets, fill, tail_any=False):
        """One double per lane from the given byte offsets (None -> `fill`).  tail_any: the lanes beyond the
        list may read any double of the slice (their value is never used)."""
        offs = [fill if o is None else o for o in offsets]
        care = NLANE
        if tail_any and len(offs) < NLANE and all(offs[l] == offs[0] + 8 * l for l in range(len(offs))):
            care = len(offs)
            offs += [offs[0] + 8 * l for l in range(len(offs), NLANE)]
        offs += [fill] * (NLANE - len(offs))
        if all(o == offs[0] for o in offs):
            return self.lds(None, offs[0], 0)
        if all(offs[l] == offs[0] + 8 * l for l in range(NLANE)):
            return self.lds(None, offs[0], 1, care)
        return self.lds(self.ldt(offs), 0, 0)

    def scatter_plan(self, offsets):
        """Where the lanes whose offset is not None will store: (ivar, imm, lc, predicate or None)."""
        live = [l for l, o in enumerate(offsets) if o is not None]
        if not live:
            return None
        full = list(offsets) + [None] * (NLANE - len(offsets))
        if len(live) == NLANE and all(full[l] == full[0] + 8 * l for l in range(NLANE)):
            return (None, full[0], 1, None)
        lo, hi = live[0], live[-1]
        if live == list(range(lo, hi + 1)) and all(full[l] == full[lo] + 8 * (l - lo) for l in live):
            p = self.p()  # a contiguous run of lanes: lo <= lane <= hi
            self.emit("prange", p, lo, hi)
            return (None, full[lo] - 8 * lo, 1, p)
        iv = self.ldt([0xffff if o is None else o for o in full])
        p = self.p()
        self.emit("pne", p, iv, 0xffff)
        return (iv, 0, 0, p)

    def scatter(self, offsets, v, plan=None):
        """Store v of the lanes whose offset is not None."""
        plan = plan or self.scatter_plan(offsets)
        if plan is not None:
            self.emit("sts", plan[0], plan[1], plan[2], v, plan[3])

    def op(self, kind, *a):
        v = self.d()
        self.emit(kind, v, *a)
        return v

    def gather2(self, st, offsets):
        """gather() of KKT entries that may lie in the value row (global scratch) AND in shared memory (the (z,z)
        diagonal): one load per address space, each padded with its own zero, joined by a lane mask."""
        inv = [st.in_v(o) for o in offsets]
        has_v = any(inv)
        has_s = any(o is not None and not v for o, v in zip(offsets, inv))
        if not has_v:
            return self.gather(offsets, st.ZERO)
        row = st.frow([o if v else None for o, v in zip(offsets, inv)])
        if not has_s:
            return self.gather(offsets, st.VZERO) if row is None else self.lds(None, row, 1)
        gv = self.gather([o if v else None for o, v in zip(offsets, inv)], st.VZERO) if row is None else self.lds(None, row, 1)
        gs = self.gather([None if v else o for o, v in zip(offsets, inv)], st.ZERO)
        p = self.p()
        self.emit("pmask", p, sum(1 << l for l, v in enumerate(inv) if v))
        return self.op("sel", p, gv, gs)

    def chain(self, acc, terms_per_lane, fill):
        """acc -= S[a_t] * S[b_t] for t = 0, 1, ... per lane: terms_per_lane[l] is lane l's list of (a, b) byte
        offsets; shorter lists are padded with the constant +0.0 (acc - (+0)(+0) leaves every acc unchanged).
        One table-driven loop in the generated code."""
        steps = max((len(t) for t in terms_per_lane), default=0)
        if steps == 0:
            return acc
        ta = tb = None
        for t in range(steps):
            ra = [tl[t][0] if t < len(tl) else fill for tl in terms_per_lane]
            rb = [tl[t][1] if t < len(tl) else fill for tl in terms_per_lane]
            ra += [fill] * (NLANE - len(ra))
            rb += [fill] * (NLANE - len(rb))
            oa, ob = self.table(ra), self.table(rb)
            if t == 0:
                ta, tb = oa, ob
            assert oa == ta + 2 * NLANE * t and ob == tb + 2 * NLANE * t
        v = self.d()
        self.emit("chain", v, acc, ta, tb, steps)
        return v


def regularise(d):
    sign = np.where(d <= 0.0, -1.0, 1.0)
    return np.where(sign * d <= 0.00000000000001, sign * 0.0000001, d)


class Emulator:
    """Runs an IR on one slice S (float64, indexed by double) with 32 lanes in lock step, and checks that
    every cross-lane communication through shared memory is separated by a ("sync",)."""

    def __init__(self, ir: IR, S):
        self.ir, self.S = ir, S
        self.tab = np.asarray(ir.tab, dtype=np.int64)
        self.lanes = np.arange(NLANE)
        self.written = {}  # double index -> set of lanes that stored it since the last sync
        self.read = {}

    def _addr(self, env, ivar, imm, lc, live=None):
        a = np.full(NLANE, imm, dtype=np.int64) + lc * 8 * self.lanes
        if ivar is not None:
            a = a + env[ivar]
        if live is None:
            live = np.ones(NLANE, bool)
        assert np.all(a[live] % 8 == 0) and np.all(a[live] >= 0) and np.all(a[live] < 8 * len(self.S))
        return np.where(live, a // 8, 0)

    def run(self):
        S, env, f64 = self.S, {}, np.float64
        L = self.lanes
        with np.errstate(all="ignore"):
            for st in self.ir.st:
                k = st[0]
                if k == "ldt":
                    env[st[1]] = self.tab[st[2] + L]
                elif k == "ldt2":
                    w = np.asarray(self.ir.tab32, dtype=np.int64)[st[3] + L]
                    env[st[1]] = w & 0xffff
                    env[st[2]] = w >> 16
                elif k == "lds":
                    a = self._addr(env, st[2], st[3], st[4])
                    for l in range(st[5]):
                        w = self.written.get(int(a[l]))
                        assert not w or w <= {l}, f"lane {l} reads double {a[l]} stored by {w} without a sync"
                        self.read.setdefault(int(a[l]), set()).add(l)
                    v = S[a].astype(f64)
                    v[st[5]:] = np.nan
                    env[st[1]] = v
                elif k == "sts":
                    pred = env[st[5]] if st[5] is not None else np.ones(NLANE, bool)
                    a = self._addr(env, st[1], st[2], st[3], pred)
                    val = env[st[4]]
                    targets = {}
                    for l in range(NLANE):
                        if not pred[l]:
                            continue
                        ai = int(a[l])
                        r = self.read.get(ai)
                        assert not r or r <= {l}, f"lane {l} overwrites double {ai} read by {r} without a sync"
                        w = self.written.get(ai)
                        same = all(np.array_equal(val[l], val[o], equal_nan=True) for o in (w or ()))
                        assert not w or w <= {l} or (ai in targets and same), f"two lanes store double {ai} without a sync"
                        if ai in targets:
                            assert np.array_equal(targets[ai], val[l], equal_nan=True), "lanes store different values to one double"
                        targets[ai] = val[l]
                        self.written.setdefault(ai, set()).add(l)
                        S[ai] = val[l]
                elif k == "sync":
                    self.written.clear()
                    self.read.clear()
                elif k == "const":
                    env[st[1]] = np.full(NLANE, st[2], dtype=f64)
                elif k == "add0":
                    env[st[1]] = env[st[2]] + f64(0.0)
                elif k == "msub":
                    env[st[1]] = env[st[2]] - env[st[3]] * env[st[4]]
                elif k == "msubp":
                    env[st[1]] = np.where(env[st[2]], env[st[3]] - env[st[4]] * env[st[5]], env[st[3]])
                elif k == "mul":
                    env[st[1]] = env[st[2]] * env[st[3]]
                elif k == "sub":
                    env[st[1]] = env[st[2]] - env[st[3]]
                elif k == "div":
                    env[st[1]] = env[st[2]] / env[st[3]]
                elif k == "regu":
                    env[st[1]] = regularise(env[st[2]])
                elif k == "selc":
                    env[st[1]] = np.where(env[st[2]], env[st[3]], f64(st[4]))
                elif k == "sel":
                    env[st[1]] = np.where(env[st[2]], env[st[3]], env[st[4]])
                elif k == "pnez":
                    env[st[1]] = env[st[2]] != 0.0
                elif k == "add":
                    env[st[1]] = env[st[2]] + env[st[3]]
                elif k == "mulc":
                    env[st[1]] = env[st[2]] * f64(st[3])
                elif k == "chain":
                    acc = env[st[2]].copy()
                    for t in range(st[5]):
                        ia = self.tab[st[3] + 2 * NLANE * t + L] // 8
                        ib = self.tab[st[4] + 2 * NLANE * t + L] // 8
                        for arr in (ia, ib):
                            for l in range(NLANE):
                                w = self.written.get(int(arr[l]))
                                assert not w or w <= {l}, f"chain: lane {l} reads double {arr[l]} stored by {w} without a sync"
                                self.read.setdefault(int(arr[l]), set()).add(l)
                        acc = acc - S[ia].astype(f64) * S[ib].astype(f64)
                    env[st[1]] = acc
                elif k == "shfl":
                    env[st[1]] = np.full(NLANE, env[st[2]][st[3]], dtype=f64)
                elif k == "pne":
                    env[st[1]] = env[st[2]] != st[3]
                elif k == "prange":
                    env[st[1]] = (L >= st[2]) & (L <= st[3])
                elif k == "pmask":
                    env[st[1]] = ((st[2] >> L) & 1).astype(bool)
                else:
                    raise ValueError(k)
        return S


# ---------------------------------------------------------------------------------------------------
# Structure of one pattern: everything the generators need, taken from the plan and from the VERIFIED
# micro-programs of the throughput kernel (tests/test_stream_program.py pins those to the reference).
# ---------------------------------------------------------------------------------------------------
LAYOUT_FIELDS = ["zero", "V", "c", "h", "b", "zd", "Lx", "D", "T", "rhs", "x", "y", "z", "s", "r", "lam", "ds", "dz", "dsl", "total", "tsize"]


def compact_layout(n, m, p, N, nV, lnz, vglobal=True, vsize=None, pbsize=64):
    """SliceLayout (csrc/program.h) with the scratch area T cut down to what the generated programs use: the
    permuted work vector of the solves (N) and the products of one column of the backward sweep (64); the
    kernel's own scratch (-(P x): n, rho's products: 2m) fits in the same doubles.

    vglobal: the value row V = P | A | G (read-only during a solve, a third of the slice) is NOT kept in shared
    memory: it lies in the warp's scratch row in global memory (L2-resident; written once per QP with the v + 0.0
    normalisation), followed by a zero.  The programs keep ONE address space -- a double of the slice is in shared
    memory below `smem`, in the scratch row from `V` on -- so the host emulator is unchanged; only the printed CUDA
    distinguishes the two (print_function).  `total` covers both, `smem` is what a warp occupies on chip."""
    even = lambda v: (v + 1) & ~1
    lay, at = {}, 0

    def take(name, count):
        nonlocal at
        lay[name] = at
        at += even(count)

    take("zero", 2)
    head = (("V", nV),) if not vglobal else ()
    for name, cnt in head + (("c", n), ("h", m), ("b", p), ("zd", m), ("Lx", lnz), ("D", N)):
        take(name, cnt)
    lay["pbsize"] = even(pbsize)  # (the products of one column of the backward sweep: gen_solve)
    lay["tsize"] = max(even(N) + lay["pbsize"], even(2 * m), even(n))
    take("T", lay["tsize"])
    for name, cnt in (("rhs", N), ("x", n), ("y", p), ("z", m), ("s", m), ("r", N), ("lam", max(m, n)), ("ds", m), ("dz", m), ("dsl", m)):
        take(name, cnt)
    lay["smem"] = at
    if vglobal:
        take("V", nV if vsize is None else vsize)  # (vsize: the row in residual order, Structure.vslots)
        take("vzero", 2)
    else:
        lay["vzero"] = lay["zero"]
    lay["total"] = at
    lay["vglobal"] = bool(vglobal)
    return lay


def resid_chains(pat):
    """computeresiduals' sparse products (Auxilary.c:753-785) as the chains gen_resid runs, one output row per lane:
    a list of groups (the rows of one chain() call), each a list of per-lane term lists [(value index into P | A | G,
    vector name, vector index), ...] in the reference's order -- column scatter delivers a row's terms in ascending
    column order, column gather a column's terms in storage order.  Groups: per 32 rows of rx [P x] then [G'z + A'y];
    then per 32 rows of ry | rz [A x or G x]."""
    n, m, p = pat.n, pat.m, pat.p
    nnzP, nnzA = len(pat.Pir), len(pat.Air)

    def rows_of(jc, ir_, nrows, voff, vec):
        out = [[] for _ in range(nrows)]
        for col in range(len(jc) - 1):
            for q in range(int(jc[col]), int(jc[col + 1])):
                out[int(ir_[q])].append((voff + q, vec, col))
        return out

    def cols_of(jc, ir_, voff, vec):
        return [[(voff + q, vec, int(ir_[q])) for q in range(int(jc[col]), int(jc[col + 1]))] for col in range(len(jc) - 1)]

    Px = rows_of(pat.Pjc, pat.Pir, n, 0, "x")
    Gtz = cols_of(pat.Gjc, pat.Gir, nnzP + nnzA, "z")
    Aty = cols_of(pat.Ajc, pat.Air, nnzP, "y") if p else [[] for _ in range(n)]
    Ax = rows_of(pat.Ajc, pat.Air, p, nnzP, "x") if p else []
    Gx = rows_of(pat.Gjc, pat.Gir, m, nnzP + nnzA, "x")
    groups = []
    for w0 in range(0, n, NLANE):
        rows = list(range(w0, min(n, w0 + NLANE)))
        groups.append([Px[i] for i in rows])
        groups.append([Gtz[i] + Aty[i] for i in rows])
    outs = Ax + Gx
    for w0 in range(0, len(outs), NLANE):
        groups.append(outs[w0:w0 + NLANE])
    return groups


FROW_BUDGET = {}  # factor-order rows per pattern (see Structure): none by default


class Structure:
    def __init__(self, plan, vglobal=True, frow_budget=0, pbsize=64):
        info = plan.stream_info()
        assert info["ok"] == 1
        sl = info["layout"]
        self.N = N = plan.info("N")
        pat0 = plan.pattern
        nV0 = len(pat0.Pir) + len(pat0.Air) + len(pat0.Gir)
        # Value row in the global scratch row: laid out in RESIDUAL ORDER -- slot 32 T + lane holds the value that lane
        # multiplies at step T of the residual products (resid_chains), +0.0 where the lane has no term -- so that the
        # residuals read it with coalesced loads at constant addresses instead of table-driven gathers.  A value that two
        # products use (A and G: A x and A'y, G x and G'z) has two slots; the factorisation reads the first.
        self.resid_groups = resid_chains(pat0)
        self.vslots = None
        vsize = None
        if vglobal and int(os.environ.get("QPB_GEN_RESID_ORDER", "1")):
            self.vslots = [[] for _ in range(nV0)]
            T = 0
            for grp in self.resid_groups:
                for t in range(max((len(tl) for tl in grp), default=0)):
                    for lane, tl in enumerate(grp):
                        if t < len(tl):
                            self.vslots[tl[t][0]].append(NLANE * T + lane)
                    T += 1
            assert all(1 <= len(v) <= 2 for v in self.vslots), "a value the residual products never read (or read more than twice)"
            vsize = NLANE * T
        self.lay = compact_layout(pat0.n, pat0.m, pat0.p, N, nV0, plan.info("lnz"), vglobal, vsize, pbsize)
        lay = self.lay
        # ... and the gathers of the factorisation (a pivot row's K entries, a leaf column's entries by row) get rows of their
        # own BEHIND the value row, 32 slots each in lane order: a third and fourth copy of an entry, coalesced as well
        # -- as long as the scratch rows of all warps in flight stay well inside L2.  Measured on B200 (65,536 QPs, M solves/s;
        # rows: 0 / 20 / 40 / 55 / 70 / 90 / all): trot (131 gathers) 1.840 / 1.858 / 1.872 / 1.882 / 1.887 / 1.855 / 1.595,
        # stance 1.601 / 1.488 / 1.398 / - / 1.404 / - / 0.950, crawl 1.661 / 1.657 / 1.611 / - / 1.503 / - / 0.991: the budget is
        # a property of the pattern AND of the warps per SM -- with the 12 warps per SM that trot runs since its loads became
        # cheap (below), rows 0 / 30 / 55: 2.047 / 2.004 / 2.032 -- so there are none by default (frow_budget, FROW_BUDGET;
        # QPB_GEN_FROWS overrides); gathers beyond the budget stay table-driven gathers into the residual-order slots.
        self.frows, self.frow_of, self.F0 = [], {}, lay["total"]
        self.frow_budget = int(os.environ.get("QPB_GEN_FROWS", str(frow_budget))) if self.vslots is not None else 0
        self.Lp = [int(v) for v in plan.array("Lp")]
        self.Li = [int(v) for v in plan.array("Li")]
        self.P = [int(v) for v in plan.array("P")]
        self.Parent = [int(v) for v in plan.array("Parent")]
        self.lnz = self.Lp[N]
        pat = plan.pattern
        self.pattern = pat
        self.n, self.m, self.p = pat.n, pat.m, pat.p
        nV = len(pat.Pir) + len(pat.Air) + len(pat.Gir)
        self.ZERO = lay["zero"] * 8
        self.VZERO = lay["vzero"] * 8          # the zero of the value row's address space
        self.VBYTE = lay["V"] * 8 if lay["vglobal"] else None

        def slice_offset(e):
            """stream-layout element -> byte offset in the warp kernel's slice (None: the constant +0.0)"""
            if e == sl["zero"]:
                return None
            if sl["V"] <= e < sl["V"] + nV:
                return (lay["V"] + (self.vslots[e - sl["V"]][0] if self.vslots else e - sl["V"])) * 8
            assert sl["zd"] <= e < sl["zd"] + self.m, "KKT entry that is neither a matrix value nor the (z,z) diagonal"
            return (lay["zd"] + e - sl["zd"]) * 8

        rows = CG.parse_rows(CG.live_ops(plan.stream_words("factor"), sl["zero"]))
        assert rows is not None and len(rows) == N
        self.pat, self.slot, self.kinit, self.kdiag = [], {}, {}, []
        colpos = {}
        for j in range(N):
            for q in range(self.Lp[j], self.Lp[j + 1]):
                colpos[(self.Li[q], j)] = q
        self.colpos = colpos
        for k, r in enumerate(rows):
            cols = [o.f1 - sl["D"] for o in r["divacc"]]
            self.pat.append(cols)
            self.kdiag.append(slice_offset(r["initd"].f1))
            for t, ch in enumerate(r["chains"]):
                i = cols[t]
                q = r["divacc"][t].c - sl["Lx"]
                assert colpos[(k, i)] == q
                first = ch[0]
                self.kinit[(k, i)] = slice_offset(first.f1) if (first.w & CG.INIT) else None
                # the chain's terms are exactly the earlier pattern columns whose column holds row i, in pattern order
                terms = [o for o in ch if not (o.w & CG.INIT)]
                want = [(colpos[(i, cols[b])], b) for b in range(t) if (i, cols[b]) in colpos]
                got = [(o.f1 - sl["Lx"], o.f2) for o in terms]
                assert got == want, "micro-program and structure disagree"
        self.R = [k for k in range(N) if self.pat[k]]
        self.leaf = [k for k in range(N) if not self.pat[k]]
        self.is_leaf = [not self.pat[k] for k in range(N)]

    def col(self, j):
        return [(self.Li[q], q) for q in range(self.Lp[j], self.Lp[j + 1])]

    def frow(self, offsets):
        """Byte offset of the factor-order row that holds, in lane order, the value-row entries at `offsets` (None: +0.0);
        None when there are no such rows (value row in shared memory or in storage order, budget used up)."""
        if not self.frow_budget:
            return None
        first = {self.lay["V"] * 8 + 8 * sl_[0]: e for e, sl_ in enumerate(self.vslots)}
        key = tuple(None if o is None else first[o] for o in offsets) + (None,) * (NLANE - len(offsets))
        if key not in self.frow_of:
            if len(self.frows) >= self.frow_budget:
                return None
            self.frow_of[key] = len(self.frows)
            self.frows.append(key)
            self.lay["total"] += NLANE
        return (self.F0 + NLANE * self.frow_of[key]) * 8

    def value_slots(self):
        """Per entry of P | A | G: its slots in the scratch row (doubles from the row's start): residual order first."""
        out = [list(v) for v in self.vslots]
        for r, key in enumerate(self.frows):
            for lane, e in enumerate(key):
                if e is not None:
                    out[e].append(self.F0 - self.lay["V"] + NLANE * r + lane)
        return out

    def place_values(self, S, vals):
        """The value row P | A | G (as v + 0.0) where the programs expect it in the slice S: in storage order, or -- value
        row in the scratch row -- in residual order with +0.0 in the slots without a term (what the kernel does on load)."""
        lay = self.lay
        if self.vslots is None:
            S[lay["V"]:lay["V"] + len(vals)] = np.asarray(vals) + 0.0
            return
        S[lay["V"]:lay["total"]] = 0.0
        for v, slots in zip(vals, self.value_slots()):
            for sl_ in slots:
                S[lay["V"] + sl_] = v + 0.0

    def in_v(self, off):
        """Is the byte offset in the value row's address space (global scratch row) rather than in shared memory?"""
        return self.VBYTE is not None and off is not None and off >= self.VBYTE

    def Lx(self, q):
        return (self.lay["Lx"] + q) * 8

    def D(self, i):
        return (self.lay["D"] + i) * 8


def common_order(st: Structure, rows):
    """A sequence of the columns in the patterns of `rows` that contains every row's Pattern as a subsequence
    and puts every column after its descendants in the elimination tree (smallest index first among the
    free ones).  None if the rows' orders contradict each other."""
    cols = sorted({i for k in rows for i in st.pat[k]})
    inset = set(cols)
    succ = {i: set() for i in cols}
    indeg = {i: 0 for i in cols}

    def edge(a, b):
        if b not in succ[a]:
            succ[a].add(b)
            indeg[b] += 1

    for k in rows:
        for a, b in zip(st.pat[k], st.pat[k][1:]):
            edge(a, b)
    for i in cols:
        a = st.Parent[i]
        while a >= 0 and a not in inset:
            a = st.Parent[a]
        if a >= 0:
            edge(i, a)
    heap = [i for i in cols if indeg[i] == 0]
    heapq.heapify(heap)
    out = []
    while heap:
        i = heapq.heappop(heap)
        out.append(i)
        for b in succ[i]:
            indeg[b] -= 1
            if indeg[b] == 0:
                heapq.heappush(heap, b)
    return out if len(out) == len(cols) else None


def split_passes(st: Structure):
    """Consecutive groups of at most 32 rows of R; the cut that makes the passes cheapest (a pass pays for every
    entry (j, i) of L with j below its last row)."""
    R = st.R
    if len(R) <= NLANE:
        return [R]
    assert len(R) <= 2 * NLANE, "more than 64 rows with a pattern: not handled"

    def cost(rows):
        hi = rows[-1]
        cols = {i for k in rows for i in st.pat[k]}
        return sum(3 * sum(1 for j, _ in st.col(i) if j < hi) + (8 if st.is_leaf[i] else 30) for i in cols)

    best = None
    for cut in range(len(R) - NLANE, NLANE + 1):
        c = cost(R[:cut]) + cost(R[cut:])
        if best is None or c < best[0]:
            best = (c, cut)
    return [R[:best[1]], R[best[1]:]]


def gen_factor(st: Structure, passes=None) -> IR:
    ir = IR("factor")
    lay, ZERO = st.lay, st.ZERO
    # ---- leaf pivots: D[i] = regularise(0 + K(i,i))  (ldl.c:284, 299, 319-320) -----------------------
    for w in range(0, len(st.leaf), NLANE):
        chunk = st.leaf[w:w + NLANE]
        v = ir.gather2(st, [st.kdiag[i] for i in chunk])
        v = ir.op("regu", v)
        ir.scatter([st.D(i) for i in chunk], v)
    ir.emit("sync")
    # ---- leaf columns of L: L(k,i) = (0 + K(i,k)) / D[i] ----------------------------------------------
    ent = [(q, k, i) for i in st.leaf for k, q in st.col(i)]
    for w in range(0, len(ent), NLANE):
        chunk = ent[w:w + NLANE]
        for q, k, i in chunk:
            assert st.kinit[(k, i)] is not None, "entry of a leaf column without a KKT entry"
        y = ir.gather2(st, [st.kinit[(k, i)] for q, k, i in chunk])
        dv = ir.gather([st.D(i) for q, k, i in chunk], st.D(chunk[0][2]))
        l = ir.op("div", y, dv)
        ir.scatter([st.Lx(q) for q, k, i in chunk], l)
    ir.emit("sync")
    # ---- the rows with a pattern, a pass of up to 32 at a time -------------------------------------------
    passes = passes or split_passes(st)
    ir.passes = passes
    for rows in passes:
        lane_of = {k: l for l, k in enumerate(rows)}
        hi = rows[-1]
        order = common_order(st, rows)
        if order is None:
            raise ValueError("the rows' Pattern orders have no common extension")
        inpat = {k: set(st.pat[k]) for k in rows}
        d = ir.gather2(st, [st.kdiag[k] for k in rows])
        Y, published = {}, set()

        def work(j):
            if j not in Y:  # Y[j] = 0 + K(j,k) where row k has that entry, +0.0 elsewhere (ldl.c:284)
                offs = [st.kinit.get((k, j)) if j in inpat[k] else None for k in rows]
                if all(o is None for o in offs):
                    v = ir.d()
                    ir.emit("const", v, 0.0)
                else:
                    v = ir.gather2(st, offs)
                Y[j] = v
            return Y[j]

        for i in order:
            if st.is_leaf[i]:
                y = ir.gather2(st, [st.kinit[(k, i)] if i in inpat[k] else None for k in rows])
                l = ir.gather([st.Lx(st.colpos[(k, i)]) if i in inpat[k] else None for k in rows], ZERO)
                d = ir.op("msub", d, l, y)  # D[k] -= l*y  (ldl.c:306)
            else:
                y = work(i)
                if i in lane_of:  # the pivot of a row of this pass: final now, every later column is an ancestor
                    dr = ir.op("regu", d)
                    di = ir.op("shfl", dr, lane_of[i])
                    ir.emit("sts", None, st.D(i), 0, di, None)
                    published.add(i)
                else:
                    di = ir.lds(None, st.D(i), 0)
                # l = y / D[i]  (ldl.c:305).  A lane whose row does not hold column i carries y = +0.0; it divides 1.0
                # instead (its l*y is still +-0, its l is never stored): a zero numerator would send the whole warp
                # through the slow path of the FP64 division.
                where = ir.scatter_plan([st.Lx(st.colpos[(k, i)]) if i in inpat[k] else None for k in rows])
                num = y if where[3] is None else ir.op("selc", where[3], y, 1.0)
                l = ir.op("div", num, di)
                d = ir.op("msub", d, l, y)
                ir.scatter(None, l, where)
                ir.emit("sync")
                del Y[i]
            for j, q in st.col(i):  # Y[j] -= L(j,i) * y  (ldl.c:302-304), for the rows below j that hold column i
                if j >= hi or not any(k > j and i in inpat[k] for k in rows):
                    continue
                lji = ir.lds(None, st.Lx(q), 0)
                Y[j] = ir.op("msub", work(j), lji, y)
        left = [k for k in rows if k not in published]
        if left:
            dr = ir.op("regu", d)
            ir.scatter([st.D(k) if k in left else None for k in rows], dr)
        ir.emit("sync")
    return pack_tables(hoist_v_loads(ir, st))


def gen_solve(st: Structure, passes=None) -> IR:
    ir = IR("solve")
    lay, ZERO, N = st.lay, st.ZERO, st.N
    Xs = lay["T"] * 8                 # the permuted work vector w[N]
    PB = (lay["T"] + N + (N & 1)) * 8  # products of one column of the backward sweep
    assert N + (N & 1) + lay["pbsize"] <= lay["tsize"], "scratch too small"
    X = lambda j: Xs + 8 * j
    passes = passes or split_passes(st)
    # ---- LDL_perm: w[j] = b[P[j]] -------------------------------------------------------------------
    for w0 in range(0, N, NLANE):
        js = list(range(w0, min(N, w0 + NLANE)))
        v = ir.gather([(lay["rhs"] + st.P[j]) * 8 for j in js], ZERO)
        ir.scatter([X(j) for j in js], v)
    ir.emit("sync")
    # ---- LDL_lsolve: w[r] -= L(r,j) w[j], j ascending (ldl.c:495-513) ---------------------------------
    rowent = {r: [] for r in st.R}
    for j in range(N):
        for r, q in st.col(j):
            rowent[r].append((j, q))
    w = [ir.gather([X(r) for r in rows], ZERO) for rows in passes]
    lane_of = {}
    for pi, rows in enumerate(passes):
        for l, r in enumerate(rows):
            lane_of[r] = (pi, l)
    nonleaf = [j for j in st.R if st.Lp[j] < st.Lp[j + 1]]
    done = {r: 0 for r in st.R}  # entries of row r consumed so far
    for stop in nonleaf + [N]:
        # every lane walks its own row through the leaf columns below `stop`
        for pi, rows in enumerate(passes):
            todo = {}
            for r in rows:
                e = rowent[r]
                t = done[r]
                while t < len(e) and e[t][0] < stop:
                    assert st.is_leaf[e[t][0]], "a non-leaf column must be consumed at its own step"
                    t += 1
                todo[r] = e[done[r]:t]
                done[r] = t
            for t in range(max((len(v) for v in todo.values()), default=0)):
                l = ir.gather([st.Lx(todo[r][t][1]) if t < len(todo[r]) else None for r in rows], ZERO)
                x = ir.gather([X(todo[r][t][0]) if t < len(todo[r]) else None for r in rows], ZERO)
                w[pi] = ir.op("msub", w[pi], l, x)
        if stop == N:
            break
        j = stop  # a column whose w[j] is a lane's accumulator: broadcast it
        pj, lj = lane_of[j]
        xj = ir.op("shfl", w[pj], lj)
        for pi, rows in enumerate(passes):
            offs = []
            for r in rows:
                e = rowent[r]
                if done[r] < len(e) and e[done[r]][0] == j:
                    offs.append(st.Lx(e[done[r]][1]))
                    done[r] += 1
                else:
                    offs.append(None)
            if all(o is None for o in offs):
                continue
            full = offs + [None] * (NLANE - len(offs))
            iv = ir.ldt([ZERO if o is None else o for o in full])
            p = ir.p()
            ir.emit("pne", p, iv, ZERO)
            l = ir.lds(iv, 0, 0)
            nw = ir.d()
            ir.emit("msubp", nw, p, w[pi], l, xj)
            w[pi] = nw
    assert all(done[r] == len(rowent[r]) for r in st.R)
    # the leaf columns before the first non-leaf one were consumed with stale w[j]?  No: a leaf ROW has no
    # entries, so w[j] of a leaf column is the right-hand side itself and is final from the start.
    for pi, rows in enumerate(passes):
        ir.scatter([X(r) for r in rows], w[pi])
    ir.emit("sync")
    # ---- LDL_dsolve: w[j] /= D[j] (ldl.c:520-532) -------------------------------------------------------
    quot = []
    for w0 in range(0, N, NLANE):  # (all loads before the first store: a lane beyond N reads its neighbour's doubles)
        js = list(range(w0, min(N, w0 + NLANE)))
        x = ir.gather([X(j) for j in js], X(js[0]), tail_any=True)
        dv = ir.gather([st.D(j) for j in js], st.D(js[0]), tail_any=True)
        # x / D with x == +-0 (right-hand sides are full of zeros) would take the slow path of the FP64 division:
        # divide 1.0 there instead and return x * D, which is the same +-0 (D is finite and not zero)
        nz = ir.p()
        ir.emit("pnez", nz, x)
        q = ir.op("div", ir.op("selc", nz, x, 1.0), dv)
        quot.append((js, ir.op("sel", nz, q, ir.op("mul", x, dv))))
    for js, v in quot:
        ir.scatter([X(j) for j in js], v)
    ir.emit("sync")
    # ---- LDL_ltsolve: w[j] -= Lx[q]*w[Li[q]], q ascending, j descending (ldl.c:539-557) ------------------
    # The columns of the last pass (the dense tail of L): lane = column.  When w[r] becomes final it is broadcast by
    # shuffle and every lane forms ITS product L(r,j) w[r] into a register named after r; the sum of column j then
    # runs over registers, in ascending r as the reference does (it is computed by all lanes, kept by lane(j)).
    last = passes[-1]
    lane_last = {j: l for l, j in enumerate(last)}
    acc = ir.gather([X(j) for j in last], ZERO)
    ir.emit("sync")  # (every lane stores the final w[r] back over what its owner has just read)
    prod = {}
    for r in reversed(last):
        ent = st.col(r)
        assert all(rr in lane_last for rr, _ in ent)
        if ent:
            t = acc
            for rr, _ in ent:
                t = ir.op("sub", t, prod[rr])
            mine = ir.p()
            ir.emit("prange", mine, lane_last[r], lane_last[r])
            acc = ir.op("sel", mine, t, acc)
        xr = ir.op("shfl", acc, lane_last[r])
        ir.emit("sts", None, X(r), 0, xr, None)
        offs = [st.Lx(st.colpos[(r, j)]) if (r, j) in st.colpos else None for j in last]
        if any(o is not None for o in offs):
            prod[r] = ir.op("mul", ir.gather(offs, ZERO), xr)
    ir.emit("sync")
    for j in reversed([c for c in nonleaf if c not in lane_last]):  # earlier passes: products through shared memory
        ent = st.col(j)
        assert len(ent) <= lay["pbsize"], "product buffer too small (Structure pbsize)"
        acc = ir.lds(None, X(j), 0)  # (read before the sync below: every lane stores the sum back over it)
        for r0 in range(0, len(ent), NLANE):
            chunk = ent[r0:r0 + NLANE]
            l = ir.gather([st.Lx(q) for r, q in chunk], st.Lx(chunk[0][1]), tail_any=True)
            x = ir.gather([X(r) for r, q in chunk], X(chunk[0][0]), tail_any=True)
            ir.scatter([PB + 8 * (r0 + t) for t in range(len(chunk))], ir.op("mul", l, x))
        ir.emit("sync")
        for t in range(len(ent)):
            acc = ir.op("sub", acc, ir.lds(None, PB + 8 * t, 0))
        ir.emit("sts", None, X(j), 0, acc, None)
        ir.emit("sync")
    leafcols = sorted((j for j in st.leaf if st.Lp[j] < st.Lp[j + 1]), key=lambda j: -(st.Lp[j + 1] - st.Lp[j]))
    for c0 in range(0, len(leafcols), NLANE):  # independent chains: one leaf column per lane
        cols = leafcols[c0:c0 + NLANE]
        acc = ir.gather([X(j) for j in cols], ZERO)
        ents = [st.col(j) for j in cols]
        for t in range(len(ents[0])):
            l = ir.gather([st.Lx(e[t][1]) if t < len(e) else None for e in ents], ZERO)
            x = ir.gather([X(e[t][0]) if t < len(e) else None for e in ents], ZERO)
            acc = ir.op("msub", acc, l, x)
        ir.scatter([X(j) for j in cols], acc)
    ir.emit("sync")
    # ---- LDL_permt: x[P[j]] = w[j] ------------------------------------------------------------------------
    for w0 in range(0, N, NLANE):
        js = list(range(w0, min(N, w0 + NLANE)))
        v = ir.gather([X(j) for j in js], X(js[0]), tail_any=True)
        ir.scatter([(lay["rhs"] + st.P[j]) * 8 for j in js], v)
    ir.emit("sync")
    return pack_tables(hoist_v_loads(ir, st))  # (no value-row loads here: the tables of per-lane offsets are requested early)


def gen_resid(st: Structure) -> IR:
    """computeresiduals' sparse products (Auxilary.c:753-785): rx = -(P x) - G'z - A'y - c, ry = -(A x) + b,
    rz = -(G x) + (h - s), and tmp = -(P x) for obj_value -- one output row per lane, its terms in the
    reference's order (column scatter delivers a row's terms in ascending column order, column gather a
    column's terms in storage order), as table-driven chains."""
    ir = IR("resid")
    lay, ZERO = st.lay, st.ZERO
    pat = st.pattern
    n, m, p = st.n, st.m, st.p
    B = lambda name, i: (lay[name] + i) * 8
    groups = iter(st.resid_groups)
    tmp = lay["T"]  # scratch: -(P x), read by the kernel's sums
    zero = ir.d()
    ir.emit("const", zero, 0.0)
    step = [0]  # steps of the residual products so far: row 32 T of the value row in residual order

    def chain(acc, grp):
        """acc -= V[a_t] * S[b_t], t ascending, per lane.  Value row in shared memory: one table-driven loop.  Value row
        in global memory: unrolled, so that the loads of V can be issued far ahead of their use (hoist_v_loads); in
        residual order they are coalesced loads at constant addresses."""
        terms_per_lane = [[(B("V", v), B(vec, i)) for v, vec, i in tl] for tl in grp]
        if not st.lay["vglobal"]:
            return ir.chain(acc, terms_per_lane, ZERO)
        for t in range(max((len(tl) for tl in terms_per_lane), default=0)):
            if st.vslots is not None:
                a = ir.lds(None, B("V", NLANE * step[0]), 1)
                step[0] += 1
            else:
                a = ir.gather([tl[t][0] if t < len(tl) else None for tl in terms_per_lane], st.VZERO)
            b = ir.gather([tl[t][1] if t < len(tl) else None for tl in terms_per_lane], ZERO)
            acc = ir.op("msub", acc, a, b)  # (a padded term is acc - (+0)(+0): no accumulator here is ever -0.0)
        return acc

    for w0 in range(0, n, NLANE):
        rows = list(range(w0, min(n, w0 + NLANE)))
        acc = chain(zero, next(groups))
        ir.scatter([(tmp + i) * 8 for i in rows], acc)
        acc = chain(acc, next(groups))
        cv = ir.gather([B("c", i) for i in rows], ZERO)
        acc = ir.op("add", acc, ir.op("mulc", cv, -1.0))  # rx[i] += c[i] * (-1)
        ir.scatter([B("r", i) for i in rows], acc)
    # ry and rz rows share waves: r[n + i] = acc + (u - v) with (u, v) = (b, +0.0) for ry (b - 0 = b*1 exactly), (h, s) for rz
    outs = [(None, B("b", i), None, B("r", n + i)) for i in range(p)] + [(None, B("h", i), B("s", i), B("r", n + p + i)) for i in range(m)]
    for w0 in range(0, len(outs), NLANE):
        chunk = outs[w0:w0 + NLANE]
        acc = chain(zero, next(groups))
        u = ir.gather([o[1] for o in chunk], ZERO)
        v = ir.gather([o[2] for o in chunk], ZERO)
        acc = ir.op("add", acc, ir.op("sub", u, v))
        ir.scatter([o[3] for o in chunk], acc)
    ir.emit("sync")
    return pack_tables(hoist_v_loads(ir, st))


def hoist_v_loads(ir: IR, st: Structure, dist=None, tdist=None):
    """Value-row loads (global memory, an L2 round trip each) are moved `dist` statements ahead of where the program
    needs them -- together with the table load that feeds them -- so that they are in flight while the warp works on
    shared memory.  Legal anywhere: the value row is read-only while the programs run.  Marks them in ir.vloads."""
    ir.vloads = set()
    if not st.lay["vglobal"]:
        return ir
    dist = int(os.environ.get("QPB_GEN_VDIST", "100")) if dist is None else dist
    # the tables of per-lane offsets are global loads too, and with the slices filling shared memory the L1 that would
    # hold them is a few KB: they come from L2, so they are requested early as well
    # (60 statements ahead with one table row per load; 30 since two rows come with one load -- pack_tables: 10 / 20 / 30 / 40
    # measured 1.730 / 1.727 / 1.741 / 1.727 M solves/s on trot)
    tdist = int(os.environ.get("QPB_GEN_TDIST", "30")) if tdist is None else tdist
    tab_of = {}
    keyed = []
    for idx, s in enumerate(ir.st):
        key = float(idx)
        if s[0] == "ldt":
            tab_of[s[1]] = s[2]
            key = idx - tdist
        elif s[0] == "lds":
            _, v, ivar, imm, lc, care = s
            lanes = range(care)
            addrs = [imm + lc * 8 * l + (ir.tab[tab_of[ivar] + l] if ivar is not None else 0) for l in lanes]
            inv = [st.in_v(a) for a in addrs]
            assert all(inv) or not any(inv), "a load that mixes the value row's address space with shared memory"
            if inv and inv[0]:
                ir.vloads.add(v)
                key = idx - dist
                if ivar is not None and keyed and keyed[-1][2][0] == "ldt" and keyed[-1][2][1] == ivar:
                    keyed[-1] = (key - 0.5, keyed[-1][1], keyed[-1][2])
        keyed.append((key, idx, s))
    keyed.sort(key=lambda t: (t[0], t[1]))
    ir.st = [s for _, _, s in keyed]
    return ir


def pack_tables(ir: IR, window=None):
    """Two table rows per load: an `ldt` and the next one within `window` statements become one `ldt2` that reads a
    32-bit word per lane (low half = the first row's offset, high half = the second's) from a second, 32-bit table;
    the rows left single move to a compacted 16-bit table.  Halves the table loads (LSU instructions and wavefronts).
    The second row is read a few statements earlier than before: the tables are read-only, nothing else changes."""
    window = int(os.environ.get("QPB_GEN_PACK", "24")) if window is None else window
    ir.tab32 = []
    if window <= 0 or any(s[0] == "chain" for s in ir.st):
        return ir
    st, out, tab16, second = ir.st, [], [], set()
    for idx, s in enumerate(st):
        if idx in second:
            continue
        if s[0] != "ldt":
            out.append(s)
            continue
        partner = next((j for j in range(idx + 1, min(len(st), idx + 1 + window)) if st[j][0] == "ldt" and j not in second), None)
        if partner is None:
            out.append(("ldt", s[1], len(tab16)))
            tab16.extend(ir.tab[s[2]:s[2] + NLANE])
        else:
            a, b = s[2], st[partner][2]
            out.append(("ldt2", s[1], st[partner][1], len(ir.tab32)))
            ir.tab32.extend(ir.tab[a + l] | (ir.tab[b + l] << 16) for l in range(NLANE))
            second.add(partner)
    ir.st, ir.tab = out, tab16
    return ir


# ---------------------------------------------------------------------------------------------------
# CUDA
# ---------------------------------------------------------------------------------------------------
def print_function(ir: IR, out, fname, tabname, chunk=1500, bar=0):
    """The statements as CUDA, in functions of about `chunk` statements cut at __syncwarp()s; values alive
    across a cut travel through the arguments (they are few: the pivots and the work registers of a pass)."""
    sts = ir.st
    # cut points: indices of sync statements
    cuts, last = [0], 0
    for idx, s in enumerate(sts):
        if s[0] == "sync" and idx + 1 - last >= chunk:
            cuts.append(idx + 1)
            last = idx + 1
    if cuts[-1] != len(sts):
        cuts.append(len(sts))
    defined_in = {}
    for idx, s in enumerate(sts):
        if s[0] == "ldt2":
            defined_in[s[1]] = defined_in[s[2]] = idx
        if s[0] in ("ldt", "lds", "const", "add0", "msub", "msubp", "mul", "sub", "div", "regu", "shfl", "pne", "prange", "pmask", "selc", "sel", "pnez", "add", "mulc", "chain"):
            defined_in[s[1]] = idx

    def uses(s):
        k = s[0]
        if k == "lds":
            return [s[2]] if s[2] else []
        if k == "sts":
            return [v for v in (s[1], s[4], s[5]) if v]
        if k in ("add0", "regu"):
            return [s[2]]
        if k in ("msub",):
            return list(s[2:5])
        if k == "msubp":
            return list(s[2:6])
        if k in ("mul", "sub", "div"):
            return list(s[2:4])
        if k == "shfl":
            return [s[2]]
        if k == "selc":
            return [s[2], s[3]]
        if k == "sel":
            return list(s[2:5])
        if k in ("pnez", "mulc", "chain"):
            return [s[2]]
        if k == "add":
            return list(s[2:4])
        if k == "pne":
            return [s[2]]
        return []

    # peephole: two loads of adjacent doubles (same addressing, 16-byte aligned, no store or sync between them)
    # become one ld.shared.v2.f64
    vloads = getattr(ir, "vloads", set())
    paired, skip = {}, set()
    for idx, s in enumerate(sts):
        if s[0] != "lds" or idx in skip or s[2] is not None or s[4] != 0 or s[3] % 16 != 0 or s[1] in vloads:
            continue
        for j in range(idx + 1, min(idx + 12, len(sts))):
            t = sts[j]
            if t[0] in ("sts", "sync", "chain"):
                break
            if t[0] == "lds" and j not in skip and t[2] is None and t[4] == 0 and t[3] == s[3] + 8 and t[1] not in vloads:
                paired[idx] = t[1]
                skip.add(j)
                break
    names = []
    for c in range(len(cuts) - 1):
        lo, hi = cuts[c], cuts[c + 1]
        live_in = []
        seen = set()
        for s in sts[lo:hi]:
            for u in uses(s):
                if defined_in[u] < lo and u not in seen:
                    seen.add(u)
                    live_in.append(u)
        live_out = []
        seen = set()
        for s in sts[hi:]:
            for u in uses(s):
                if lo <= defined_in[u] < hi and u not in seen:
                    seen.add(u)
                    live_out.append(u)
        names.append((f"{fname}_{c}", live_in, live_out, lo, hi))

    def ctype(v):
        return {"d": "double", "i": "unsigned", "p": "bool"}[v[0]]

    for name, live_in, live_out, lo, hi in names:
        args = ["const unsigned sb", "const unsigned sbl", "const unsigned lane", "const unsigned short *__restrict__ tl", "const unsigned *__restrict__ tl2", "const unsigned nbar", "const char *vb", "const char *vbl"]
        args += [f"const {ctype(v)} {v}" for v in live_in]
        args += [f"{ctype(v)} &o_{v}" for v in live_out]
        out.append(f"static __device__ __noinline__ void {name}({', '.join(args)})\n{{")
        since = 0
        for idx in range(lo, hi):
            s = sts[idx]
            k = s[0]
            since += 1
            if idx in skip:
                continue
            if idx in paired:
                out.append(f"\tdouble {s[1]}, {paired[idx]}; asm volatile(\"ld.shared.v2.f64 {{%0, %1}}, [%2+{s[3]}];\" : \"=d\"({s[1]}), \"=d\"({paired[idx]}) : \"r\"(sb));")
                continue
            if bar and k == "sync" and since >= bar:
                # the warps of the CTA that run this program meet here: they then fetch the same code at the same time
                out.append('\tasm volatile("bar.sync 1, %0;" :: "r"(nbar) : "memory");')
                since = 0
                continue
            if k == "ldt":
                out.append(f"\tconst unsigned {s[1]} = __ldg(tl + {s[2]});")
            elif k == "ldt2":
                out.append(f"\tconst unsigned w{s[1]} = __ldg(tl2 + {s[3]}); const unsigned {s[1]} = w{s[1]} & 0xffffu, {s[2]} = w{s[1]} >> 16;")
            elif k == "lds" and s[1] in vloads:
                # the value row: global memory (the warp's scratch row; vb is pre-offset so that slice offsets apply), L2 only
                base = "vbl" if s[4] else "vb"
                addr = f"{base} + {s[2]}" if s[2] else base
                out.append(f"\tdouble {s[1]}; asm volatile(\"ld.global.cg.f64 %0, [%1+{s[3]}];\" : \"=d\"({s[1]}) : \"l\"({addr}));")
            elif k == "lds":
                base = "sbl" if s[4] else "sb"
                if s[2]:
                    out.append(f"\tdouble {s[1]}; asm volatile(\"ld.shared.f64 %0, [%1+{s[3]}];\" : \"=d\"({s[1]}) : \"r\"({base} + {s[2]}));")
                else:
                    out.append(f"\tdouble {s[1]}; asm volatile(\"ld.shared.f64 %0, [%1+{s[3]}];\" : \"=d\"({s[1]}) : \"r\"({base}));")
            elif k == "sts":
                base = "sbl" if s[3] else "sb"
                addr = f"{base} + {s[1]}" if s[1] else base
                st_ = f"asm volatile(\"st.shared.f64 [%0+{s[2]}], %1;\" :: \"r\"({addr}), \"d\"({s[4]}) : \"memory\");"
                out.append(f"\tif ({s[5]}) {{ {st_} }}" if s[5] else f"\t{st_}")
            elif k == "sync":
                out.append("\t__syncwarp();")
            elif k == "const":
                out.append(f"\tconst double {s[1]} = {s[2]!r};")
            elif k == "add0":
                out.append(f"\tconst double {s[1]} = {s[2]} + 0.0;")
            elif k == "msub":
                out.append(f"\tconst double {s[1]} = {s[2]} - {s[3]} * {s[4]};")
            elif k == "msubp":
                out.append(f"\tconst double {s[1]} = {s[2]} ? {s[3]} - {s[4]} * {s[5]} : {s[3]};")
            elif k == "mul":
                out.append(f"\tconst double {s[1]} = {s[2]} * {s[3]};")
            elif k == "sub":
                out.append(f"\tconst double {s[1]} = {s[2]} - {s[3]};")
            elif k == "div":
                out.append(f"\tconst double {s[1]} = {s[2]} / {s[3]};")
            elif k == "regu":
                out.append(f"\tconst double {s[1]} = regularise({s[2]});")
            elif k == "selc":
                out.append(f"\tconst double {s[1]} = {s[2]} ? {s[3]} : {s[4]!r};")
            elif k == "sel":
                out.append(f"\tconst double {s[1]} = {s[2]} ? {s[3]} : {s[4]};")
            elif k == "pnez":
                out.append(f"\tconst bool {s[1]} = {s[2]} != 0.0;")
            elif k == "add":
                out.append(f"\tconst double {s[1]} = {s[2]} + {s[3]};")
            elif k == "mulc":
                out.append(f"\tconst double {s[1]} = {s[2]} * {s[3]!r};")
            elif k == "chain":
                out.append(f"\tconst double {s[1]} = lane_chain({s[2]}, sb, tl + {s[3]}, tl + {s[4]}, {s[5]});")
            elif k == "shfl":
                out.append(f"\tconst double {s[1]} = __shfl_sync(0xffffffffu, {s[2]}, {s[3]});")
            elif k == "pne":
                out.append(f"\tconst bool {s[1]} = {s[2]} != {s[3]}u;")
            elif k == "prange":
                out.append(f"\tconst bool {s[1]} = lane - {s[2]}u <= {s[3] - s[2]}u;")
            elif k == "pmask":
                out.append(f"\tconst bool {s[1]} = ({s[2]:#x}u >> lane) & 1u;")
            else:
                raise ValueError(k)
        for v in live_out:
            out.append(f"\to_{v} = {v};")
        out.append("}\n")
    out.append(f"static __device__ __forceinline__ void {fname}(const unsigned sb, const unsigned nbar, const char *vb)\n{{")
    out.append(f"\tconst unsigned lane = threadIdx.x & 31u, sbl = sb + 8u * lane;\n\tconst unsigned short *tl = {tabname} + lane;\n\tconst unsigned *tl2 = {tabname}32 + lane;\n\tconst char *vbl = vb + 8u * lane;")
    declared = set()
    for name, live_in, live_out, lo, hi in names:
        for v in live_out:
            if v not in declared:
                out.append(f"\t{ctype(v)} {v};")
                declared.add(v)
        out.append(f"\t{name}({', '.join(['sb', 'sbl', 'lane', 'tl', 'tl2', 'nbar', 'vb', 'vbl'] + live_in + live_out)});")
    out.append("}\n")


def generate(kind: str, chunk=1500, bar=0, vglobal=True):
    from . import dogbot, qpswift as Q
    plan = Q.Plan(dogbot.generate(kind, 1).pattern)
    key = plan.stream_key()
    st = Structure(plan, vglobal, FROW_BUDGET.get(kind, 0))
    if vglobal:
        # One more slice per CTA where a product buffer of the size the pattern needs (instead of 64 doubles) makes room for
        # it: crawl 10 -> 11 warps (2,660 -> 2,638 doubles per slice; 11 slices of 2,641 fit).
        fits = lambda doubles: min(int(os.environ.get("QPB_GEN_MAX_WARPS", "12")), (227 * 1024) // (doubles * 8))
        last = set(split_passes(st)[-1])
        need = max([len(st.col(j)) for j in st.R if st.Lp[j] < st.Lp[j + 1] and j not in last], default=0)
        slim = compact_layout(st.n, st.m, st.p, st.N, 0, st.lnz, True, 0, need)
        if fits(slim["smem"]) > fits(st.lay["smem"]) and int(os.environ.get("QPB_GEN_SLIM", "1")):
            st = Structure(plan, vglobal, FROW_BUDGET.get(kind, 0), need)
    progs = (gen_factor(st), gen_solve(st), gen_resid(st))
    ns = f"genw_{kind}" if vglobal else f"genws_{kind}"  # (throughput variant / latency variant: value row in shared memory)
    out = [f"// GENERATED by apf_quadruped_b200/codegen_warp.py for the '{kind}' pattern (pattern key {key:#018x}); do not edit.",
           f"// passes: {[(r[0], r[-1]) for r in progs[0].passes]}; value row {'in the global scratch row' if vglobal else 'in shared memory'}",
           '#include "warp_kernel.cuh"', "", "namespace qpb {", f"namespace {ns} {{", ""]
    if st.vslots is not None:  # where the kernel puts a QP's values: first and second slot of every entry of P | A | G
        slots = st.value_slots()
        kdst = max(len(v) for v in slots)
        dst = [v[k] if k < len(v) else 0xffff for k in range(kdst) for v in slots]
        assert max(v for v in dst if v != 0xffff) < 0xffff
        out.append(f"__device__ const unsigned short vdst[] = {{{', '.join(str(v) for v in dst)}}};\n")
        # the entries of G by rows (ascending column inside a row: the plan's grow lists, symbolic.cpp) -> a slot each: the
        # initial point's h - G x (kernels.cuh ipm_init_point) reads the value row through it
        pat = st.pattern
        nnzPA = len(pat.Pir) + len(pat.Air)
        grow = [[] for _ in range(st.m)]
        for col in range(len(pat.Gjc) - 1):
            for q in range(int(pat.Gjc[col]), int(pat.Gjc[col + 1])):
                grow[int(pat.Gir[q])].append(slots[nnzPA + q][0])
        out.append(f"__device__ const unsigned short gslot[] = {{{', '.join(str(v) for r in grow for v in r)}}};\n")
    for ir in progs:
        tabname = f"tab_{ir.name}"
        vals = ", ".join(str(v) for v in ir.tab) or "0"
        out.append(f"__device__ const unsigned short {tabname}[] = {{{vals}}};\n")
        vals32 = ", ".join(f"{v}u" for v in getattr(ir, "tab32", [])) or "0u"
        out.append(f"__device__ const unsigned {tabname}32[] = {{{vals32}}};\n")
        print_function(ir, out, ir.name, tabname, chunk, bar if ir.name != "resid" else 0)
    lay = st.lay
    # what the kernel gets: the shared-memory part of the slice (V outside it: in the warp's global scratch row)
    clay = dict(lay)
    if lay["vglobal"]:
        clay["V"] = lay["smem"]
    clay["total"] = lay["smem"]
    nv = len(st.pattern.Pir) + len(st.pattern.Air) + len(st.pattern.Gir)
    # warps per CTA: as many slices as fit, at most 12 (168 registers each).  Until the table loads were halved and the value
    # row went into residual order, the LSU / MIO queues saturated at 10 warps (11 / 12: no gain); since then trot
    # 10 / 11 / 12 warps: 1.904 / 2.021 / 2.047 M solves/s (stance: 9 and crawl: 10 slices fit anyway)
    max_warps = max(1, min(int(os.environ.get("QPB_GEN_MAX_WARPS", "12")) if vglobal else 8, (227 * 1024) // (lay["smem"] * 8)))
    out += ["struct Prog", "{",
            "\tstatic constexpr bool own_residuals = true;",
            "\tstatic constexpr bool realign = %s; // barriers inside factor / solve need every warp's count" % ("true" if bar else "false"),
            "\tstatic constexpr bool value_row_global = %s; // V = P | A | G in the warp's scratch row in global memory, not in the slice" % ("true" if lay["vglobal"] else "false"),
            "\tstatic constexpr int value_row_doubles = %d; // the row + the zero behind it, padded" % (lay["total"] - lay["smem"] if lay["vglobal"] else 0),
            "\tstatic constexpr bool value_row_resid_order = %s; // the row holds the values in the order of the residual products (slots: vdst)" % ("true" if st.vslots is not None else "false"),
            "\tstatic constexpr int kVDst = %d; // slots per entry of P | A | G (vdst[k * kNV + e], 0xffff: none)" % (max(len(v) for v in st.value_slots()) if st.vslots is not None else 0),
            "\tstatic __device__ __forceinline__ unsigned vdst(const int i) { return %s; }" % (f"__ldg({ns}::vdst + i)" if st.vslots is not None else "0u"),
            "\tstatic __device__ __forceinline__ unsigned gslot(const int t) { return %s; }" % (f"__ldg({ns}::gslot + t)" if st.vslots is not None else "0u"),
            "\tstatic constexpr bool static_shape = true;",
            "\tstatic constexpr int kN = %d, kM = %d, kP = %d, kOffA = %d, kOffG = %d, kNV = %d;" % (st.n, st.m, st.p, len(st.pattern.Pir), len(st.pattern.Pir) + len(st.pattern.Air), nv),
            "\tstatic __device__ __forceinline__ SliceLayout layout() { SliceLayout L; %s return L; }" % " ".join(f"L.{f} = {clay[f]};" for f in LAYOUT_FIELDS),
            "\tstatic __device__ __forceinline__ void factor(const ProgramView &, const unsigned sb, const unsigned nbar, const char *vb) { %s::factor(sb, nbar, vb); }" % ns,
            "\tstatic __device__ __forceinline__ void solve(const ProgramView &, const unsigned sb, const unsigned nbar, const char *vb) { %s::solve(sb, nbar, vb); }" % ns,
            "\tstatic __device__ __forceinline__ void resid(const unsigned sb, const char *vb) { %s::resid(sb, 32u, vb); }" % ns,
            "};", "", f"}} // namespace {ns}", "",
            "static const SliceLayout layout_%s = {%s};" % (ns, ", ".join(str(clay[f]) for f in LAYOUT_FIELDS)),
            f"QPB_REGISTER_GENERATED_WARP({ns}, {key:#018x}ull, layout_{ns}, {ns}::Prog, {max_warps})", "", "} // namespace qpb", ""]
    stats = {ir.name: dict(statements=len(ir.st), table=len(ir.tab), table32=len(getattr(ir, "tab32", [])),
                           **{k: sum(1 for s in ir.st if s[0] == k) for k in ("lds", "sts", "ldt", "ldt2", "msub", "div", "sync", "shfl", "chain")})
             for ir in progs}
    stats["layout"] = dict(smem=st.lay["smem"], bytes=st.lay["smem"] * 8, tsize=st.lay["tsize"], value_row_global=lay["vglobal"], nV=nv, max_warps=max_warps,
                           value_row_doubles=lay["total"] - lay["smem"] if lay["vglobal"] else 0, resid_order=st.vslots is not None, factor_rows=len(st.frows))
    plan.close()
    return "\n".join(out), stats


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("kinds", nargs="*", default=["trot", "stance", "crawl"])
    ap.add_argument("--out", default="build/gen")
    ap.add_argument("--chunk", type=int, default=1500)
    ap.add_argument("--bar", type=int, default=0, help="a CTA barrier about every so many statements of factor / solve (0: none)")
    ap.add_argument("--vsmem", action="store_true", help="ONLY the variant that keeps the value row P | A | G in shared memory (default: both variants -- "
                    "the throughput one with the value row in the warp's global scratch row, and this one for batches that are in flight all at once)")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for kind in a.kinds:
        for vglobal in ((False,) if a.vsmem else (True, False)):
            src, stats = generate(kind, a.chunk, a.bar, vglobal)
            path = os.path.join(a.out, f"warp_gen_{kind}.cu" if vglobal else f"warp_gens_{kind}.cu")
            old = open(path).read() if os.path.exists(path) else None
            if old != src:
                with open(path, "w") as f:
                    f.write(src)
            print(f"{path}:")
            for seg, st_ in stats.items():
                print(f"  {seg}: {st_}")


if __name__ == "__main__":
    main()
