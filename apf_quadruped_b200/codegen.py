"""Pattern-specialised throughput kernels: the micro-programs of csrc/stream_program.cpp turned into
straight-line CUDA.

The interpreter of csrc/stream.cuh spends ~20 instructions per micro-op on instruction supply (shuffles,
flag tests, address arithmetic); for a KNOWN sparsity pattern every index and every flag is a constant,
so the same micro-op sequence can be emitted as source: one `ld.global` with an immediate offset, one
multiply and one subtract per term.  The arithmetic of every number (which operands, in which order) is
the program's, word for word -- the programs themselves are checked against the reference on the host
(tests/test_stream_program.py) -- so the generated kernels stay bit-identical with reference qpSWIFT
(ldl.c:253-326 LDL_numeric, 495-597 the solves, Auxilary.c:745-786 computeresiduals).  What the generator
is free to choose, and chooses for the memory system, is WHEN independent numbers are computed and
WHERE a value waits for its next use:

* HBM operands are requested `dist` micro-ops before they are used (never before the last store to
  the same element);
* an operand that was loaded or stored at most `reuse` micro-ops ago is taken from its register instead
  of HBM; constants (+0.0, 1.0, -1.0) are literals;
* residual products: an entry of G or A feeds two chains ((G x)_k and (G'z)_i); the chains are interleaved so
  that both take it from one load (`share`);
* LDL_numeric: up to `rows` consecutive rows of L are computed TOGETHER (their chains interleaved), so
  that the entries of L they both read are fetched once; the row-local work vectors Y of these rows
  live in registers (every index is known) as far as `ybudget` doubles go, and in disjoint parts of
  the shared-memory window beyond; the quotient L(k,i) = y_i / D_i and the pivot update follow their
  chain immediately (same operations, same order per number as ldl.c:299-320).

The result is an IR (a list of statements over single-assignment variables) that is printed as CUDA and
that `evaluate` executes on the host, so the transformation itself is checked against the untransformed
program bit for bit without a GPU (tests/test_codegen.py).

    python -m apf_quadruped_b200.codegen --out build/gen trot stance crawl

writes build/gen/stream_gen_<kind>.cu; the Makefile compiles whatever it finds there into the library,
where each file registers its kernel under the key of its micro-program (QPB_PLAN_STREAM_KEY).  A
pattern without a generated kernel runs on the interpreter.
"""
from __future__ import annotations

import argparse
import os

import numpy as np

TERM, DIVACC, DIV, INITD, ENDD, LOADS, NOP = 0, 2, 3, 4, 5, 6, 8
FIRSTZ, INIT, ST_SHARED, ST_HBM, LATE1, ONE, SLOW = (1 << (52 + i) for i in range(7))

DEFAULTS = dict(dist=24, reuse=16, chunk=256, rows=3, ybudget=60, fn_ops=400, block_ops=4000, pdist=0, ring=0, group=4, ahead=6, share=1)


class Op:
    __slots__ = ("f1", "f2", "c", "op", "w", "uses_p", "uses_s")

    def __init__(self, w: int):
        self.w = w
        self.f1 = w & 0xffff
        self.f2 = (w >> 16) & 0xffff
        self.c = (w >> 32) & 0xffff
        self.op = (w >> 48) & 15
        self.uses_p = self.op in (TERM, DIVACC, DIV, INITD)
        self.uses_s = (self.op == TERM and not (w & ONE) and not (w & INIT)) or self.op in (DIVACC, DIV, LOADS)


def live_ops(words, zero):
    """The micro-ops that do something: padding (a TERM on the constant +0.0 in a fast group, a NOP
    without a store in a slow one) is dropped -- the accumulator is dead wherever the compiler pads."""
    out = []
    for w in words:
        w = int(w)
        if (w & ~SLOW) == zero:
            continue
        o = Op(w)
        if o.op == NOP and not (w & (ST_SHARED | ST_HBM)):
            continue
        out.append(o)
    return out


# ---------------------------------------------------------------------------------------------------
# IR.  Statements (tuples) over single-assignment double variables (strings); operands may also be the
# literals "0.0", "1.0", "-1.0".
#   ("ldg", v, e)            v = HBM slice[e]                 ("lds", v, e)   v = window[e]
#   ("stg", e, a)            HBM slice[e] = a                 ("sts", e, a)   window[e] = a
#   ("msub", v, base, a, b)  v = base - a*b   (b None: v = base - a)
#   ("div", v, a, b)  v = a / b               ("regu", v, a)  v = regularise(a)  (ldl.c:319-320)
# A Function is a list of statements; a "carried" one takes the accumulators (t, d) of the program in
# and hands them on (chains of the program may cross its boundaries).
# ---------------------------------------------------------------------------------------------------
class Function:
    def __init__(self, name, carried, start):
        self.name = name
        self.carried = carried  # True: double2 f(g, sb, t, d) returning the accumulators; False: void f(g, sb)
        self.items = []         # (position in micro-ops, statement)
        self.loads = []         # load records: dict(v, e, at = wanted position of the request, use = position of the first use)
        self.prefetch = []      # (position, element): prefetch.global.L2
        self.t = "t_in"
        self.d = "d_in"
        self.start = start
        self.nops = 0


class Builder:
    """Emits the statements of one segment, function by function, with the operand cache."""

    def __init__(self, seg, consts, opt, scratch=0):
        self.seg = seg
        self.scratch = scratch       # doubles of shared-memory window per lane
        self.consts = consts  # element -> literal
        self.opt = opt
        self.functions = []
        self.fn = None
        self.nvar = 0
        self.pos = 0                 # micro-ops emitted so far in this segment
        self.cache = {}              # element -> [variable, position of the last use]
        self.last_store = {}         # element -> position of the last store
        self.stats = dict(ops=0, hbm_loads=0, reused=0, lds=0, sts=0, stg=0, functions=0)

    def var(self, prefix="v"):
        self.nvar += 1
        return f"{prefix}{self.nvar}"

    def begin(self, carried):
        self.fn = Function(f"{self.seg}_{len(self.functions)}", carried, self.pos)
        self.fn.opt = self.opt
        self.functions.append(self.fn)
        self.cache = {}  # registers do not survive a call
        self.stats["functions"] += 1

    def put(self, stmt):
        self.fn.items.append((self.pos, stmt))

    def operand(self, e):
        """The value of HBM element e as a variable: literal, cached register, or a (scheduled) load."""
        if e in self.consts:
            return self.consts[e]
        hit = self.cache.get(e)
        if hit is not None and self.pos - hit[1] <= self.opt["reuse"]:
            hit[1] = self.pos
            self.stats["reused"] += 1
            return hit[0]
        v = self.var("p")
        at = max(self.pos - self.opt["dist"], self.last_store.get(e, -1) + 1, self.fn.start)
        rec = dict(v=v, e=e, at=min(at, self.pos), use=self.pos)
        self.fn.loads.append(rec)
        self.put(("use", rec))
        self.cache[e] = [v, self.pos]
        self.stats["hbm_loads"] += 1
        return v

    def stg(self, e, a):
        self.put(("stg", e, a))
        self.last_store[e] = self.pos
        self.cache[e] = [a, self.pos]
        self.stats["stg"] += 1

    def lds(self, e):
        v = self.var("s")
        self.put(("lds", v, e))
        self.stats["lds"] += 1
        return v

    def sts(self, e, a):
        self.put(("sts", e, a))
        self.stats["sts"] += 1

    def step(self):
        self.pos += 1
        self.fn.nops += 1
        self.stats["ops"] += 1

    # -- one micro-op with the window in shared memory (the program's own semantics, stream.cuh slow_op) --
    def generic(self, o):
        f, w = self.fn, o.w
        P = self.operand(o.f1) if o.uses_p else None
        S = self.lds(o.f2) if o.uses_s else None
        if o.op == TERM:
            if w & INIT:
                f.t = P
            else:
                v = self.var("t")
                self.put(("msub", v, "0.0" if w & FIRSTZ else f.t, P, None if w & ONE else S))
                f.t = v
        elif o.op == DIVACC:
            v = self.var("t")
            self.put(("div", v, S, P))
            d = self.var("d")
            self.put(("msub", d, f.d, v, S))
            f.t, f.d = v, d
        elif o.op == DIV:
            v = self.var("t")
            self.put(("div", v, S, P))
            f.t = v
        elif o.op == INITD:
            f.d = P
        elif o.op == ENDD:
            v = self.var("t")
            self.put(("regu", v, f.d))
            f.t = v
        elif o.op == LOADS:
            f.t = S
        if w & ST_HBM:
            self.stg(o.c, f.t)
        if w & ST_SHARED:
            self.sts(o.c, f.t)
        self.step()

    def generic_run(self, ops):
        """ops in program order, a new function every `chunk` micro-ops, (t, d) carried across."""
        for o in ops:
            if self.fn is None or not self.fn.carried or self.fn.nops >= self.opt["chunk"]:
                self.begin(True)
            self.generic(o)


def parse_rows(ops):
    """LDL_numeric as stream_program.cpp emits it: per row of L the chains of its pattern (each ends with
    the store of Y[t] into window slot t), INITD, one DIVACC per pattern entry, ENDD.  None: not that shape."""
    rows, cur, chain, tail = [], None, [], False
    for o in ops:
        if cur is None:
            cur = dict(chains=[], initd=None, divacc=[], endd=None, ops=0)
        cur["ops"] += 1
        if o.op == TERM and not tail:
            chain.append(o)
            if o.w & ST_HBM:
                return None
            if o.w & ST_SHARED:
                if o.c != len(cur["chains"]):
                    return None
                cur["chains"].append(chain)
                chain = []
        elif o.op == INITD and not chain and not tail:
            cur["initd"], tail = o, True
        elif o.op == DIVACC and tail and (o.w & ST_HBM) and not (o.w & ST_SHARED) and o.f2 == len(cur["divacc"]):
            cur["divacc"].append(o)
        elif o.op == ENDD and tail and (o.w & ST_HBM) and not (o.w & ST_SHARED) and len(cur["divacc"]) == len(cur["chains"]):
            cur["endd"] = o
            rows.append(cur)
            cur, tail = None, False
        else:
            return None
    if cur is not None or chain:
        return None
    for r in rows:
        for t, ch in enumerate(r["chains"]):
            if any(o.uses_s and not o.f2 < t for o in ch):
                return None
    return rows


def build_factor(B: Builder, ops):
    """Row blocks with Y in registers (module docstring); rows that do not fit keep the program's form."""
    opt = B.opt
    rows = parse_rows(ops)
    if rows is None:
        B.generic_run(ops)
        return
    # blocks of consecutive rows; the first rows of a block keep Y in registers while the budget lasts,
    # the others get their own part of the window
    k, blocks = 0, []
    while k < len(rows):
        blk, regs, smem, nops = [], 0, 0, 0
        while len(blk) < max(1, opt["rows"]) and k + len(blk) < len(rows):
            nx = rows[k + len(blk)]
            n = len(nx["chains"])
            if blk and nops + nx["ops"] > opt["block_ops"]:
                break
            if regs + n <= opt["ybudget"] and not any(c == "smem" for c, _ in blk):
                blk.append(("regs", 0))
                regs += n
            elif smem + n <= B.scratch:
                blk.append(("smem", smem))
                smem += n
            else:
                break
            nops += nx["ops"]
        assert blk, "a row's pattern exceeds the shared-memory window"
        blocks.append((k, blk))
        k += len(blk)

    for k0, classes in blocks:
        blk = list(range(k0, k0 + len(classes)))
        where = {r: classes[r - k0] for r in blk}
        if B.fn is None or B.fn.nops >= opt["fn_ops"]:
            B.begin(False)
        pending = set()  # elements this block has yet to store: nobody may read them before
        for r in blk:
            pending.update(o.c for o in rows[r]["divacc"])
            pending.add(rows[r]["endd"].c)
        nxt = {r: 0 for r in blk}
        dacc, Y, left = {}, {}, list(blk)
        while left:
            ready, finished = [], False
            for r in list(left):
                R, t = rows[r], nxt[r]
                if r not in dacc:  # D[k] = Y[k] (ldl.c:299): the diagonal entry of K, or +0.0
                    dacc[r] = B.operand(R["initd"].f1)
                    B.step()
                if t == len(R["chains"]):
                    v = B.var("D")  # ldl.c:319-320
                    B.put(("regu", v, dacc[r]))
                    B.stg(R["endd"].c, v)
                    pending.discard(R["endd"].c)
                    B.step()
                    left.remove(r)
                    finished = True
                elif not (any(o.f1 in pending for o in R["chains"][t]) or R["divacc"][t].f1 in pending):
                    ready.append(r)
            if not ready:
                assert finished, "row block cannot make progress"
                continue
            # the ready chains (one per row, mostly the same column of different rows), term by term: a chain
            # keeps its own order, and chains whose next operand is the same element of L take it together
            todo = {r: list(rows[r]["chains"][nxt[r]]) for r in ready}
            acc = {r: None for r in ready}
            turn = 0
            while any(todo.values()):
                order = [r for r in ready if todo[r]]
                lead = order[turn % len(order)]
                turn += 1
                x = todo[lead][0].f1
                for r in order:
                    if todo[r][0].f1 != x:
                        continue
                    o = todo[r].pop(0)
                    in_regs, base = where[r][0] == "regs", where[r][1]
                    P = B.operand(o.f1)  # y_t = K(t, k) - sum L(t, i) y_i in the row's own order (ldl.c:300-304)
                    if o.w & INIT:
                        acc[r] = P
                    else:
                        yi = None if o.w & ONE else (Y[(r, o.f2)] if in_regs else B.lds(base + o.f2))
                        v = B.var("y")
                        B.put(("msub", v, "0.0" if o.w & FIRSTZ else acc[r], P, yi))
                        acc[r] = v
                    B.step()
            for r in ready:
                t, dv = nxt[r], rows[r]["divacc"][nxt[r]]
                if where[r][0] == "regs":
                    Y[(r, t)] = acc[r]
                else:
                    B.sts(where[r][1] + t, acc[r])
                Dv = B.operand(dv.f1)  # l = y / D[i]; D[k] -= l*y; L(k, i) = l  (ldl.c:305-309)
                l = B.var("l")
                B.put(("div", l, acc[r], Dv))
                d2 = B.var("d")
                B.put(("msub", d2, dacc[r], l, acc[r]))
                dacc[r] = d2
                B.stg(dv.c, l)
                pending.discard(dv.c)
                B.step()
                nxt[r] += 1
    B.fn = None


def build_shared_chains(B: Builder, ops):
    """Residual products: a matrix entry is used by two chains (G(k,i) in (G x)_k and in (G'z)_i, A likewise).
    Every chain keeps its own order; a chain whose next term wants the element the leading chain has just
    loaded takes its term right behind it (one load for both), and a chain that shares nothing any more is
    finished on the spot, so that its accumulator does not stay around.  Returns False (nothing emitted) when
    the program is not a set of independent multiply-subtract chains."""
    chains = []
    for o in ops:
        if o.op not in (TERM, NOP) or (o.w & ST_SHARED):
            return False
        if not chains or (o.op == TERM and (o.w & (FIRSTZ | INIT))):
            chains.append([])
        chains[-1].append(o)
    where = {}
    for ci, ch in enumerate(chains):
        for o in ch:
            if o.w & ST_HBM:
                where.setdefault(o.c, set()).add(ci)
    for ci, ch in enumerate(chains):  # a chain may only load what it stores itself
        if any(o.uses_p and where.get(o.f1, {ci}) != {ci} for o in ch):
            return False
    uses = {}
    for ch in chains:
        for o in ch:
            if o.uses_p:
                uses[o.f1] = uses.get(o.f1, 0) + 1
    if not any(v > 1 for v in uses.values()):
        return False
    B.begin(False)
    nxt = [0] * len(chains)
    acc = [None] * len(chains)

    def emit(ci):
        o = chains[ci][nxt[ci]]
        nxt[ci] += 1
        P = B.operand(o.f1) if o.uses_p else None
        if o.uses_p:
            uses[o.f1] -= 1
        S = B.lds(o.f2) if o.uses_s else None
        if o.op == TERM:
            if o.w & INIT:
                acc[ci] = P
            else:
                v = B.var("t")
                B.put(("msub", v, "0.0" if o.w & FIRSTZ else acc[ci], P, None if o.w & ONE else S))
                acc[ci] = v
        if o.w & ST_HBM:
            B.stg(o.c, acc[ci])
        B.step()

    def private(ci):
        return all(not o.uses_p or uses[o.f1] == 1 for o in chains[ci][nxt[ci]:])

    for lead in range(len(chains)):
        while nxt[lead] < len(chains[lead]):
            o = chains[lead][nxt[lead]]
            emit(lead)
            if not o.uses_p or uses[o.f1] == 0:
                continue
            for ci in range(len(chains)):
                if ci != lead and nxt[ci] < len(chains[ci]):
                    n = chains[ci][nxt[ci]]
                    if n.uses_p and n.f1 == o.f1:
                        emit(ci)
                        if nxt[ci] > 0 and private(ci):
                            while nxt[ci] < len(chains[ci]):
                                emit(ci)
    B.fn = None
    return True


def add_l2_prefetch(B: Builder):
    """Every HBM load is announced `pdist` micro-ops ahead by a prefetch into L2 (no register, no
    scoreboard, may sit in an earlier function): the load itself then finds its line in L2."""
    pdist = B.opt["pdist"]
    if pdist <= 0:
        return
    starts = [fn.start for fn in B.functions]
    for fn in B.functions:
        for rec in fn.loads:
            at, e = rec["at"], rec["e"]
            want = max(0, at - pdist)
            k = max(i for i, s0 in enumerate(starts) if s0 <= want)
            B.functions[k].prefetch.append((want, e))
            B.stats["prefetch"] = B.stats.get("prefetch", 0) + 1


def build_segment(seg, ops, consts, opt, scratch):
    B = Builder(seg, consts, opt, scratch)
    if seg == "factor" and opt["rows"] > 0:
        build_factor(B, ops)
    elif not (seg in ("resid", "resz") and opt["share"] and build_shared_chains(B, ops)):
        B.generic_run(ops)
    add_l2_prefetch(B)
    return B


def statements(fn: Function, opt=None):
    """The function's final statement list: every HBM operand requested where it was scheduled.

    With a ring (opt["ring"] slots of one double per lane in shared memory) an operand is requested by
    cp.async -- ("cpa", slot, e) -- and taken out of its slot at its first use -- ("ldr", v, slot) -- after a
    cp.async.wait_group that covers its commit group; a warp's ld.global throughput is capped (about one
    256-byte request per 160 cycles, scripts/micro/randline3.cu), cp.async's is not.  A slot is reused
    once its previous occupant has been taken out; an operand that cannot be requested at least `ahead`
    micro-ops before its use (stored just before, or no free slot) is an ordinary ld.global."""
    if getattr(fn, "final", None) is not None:
        return fn.final
    opt = opt or getattr(fn, "opt", None) or DEFAULTS
    ring, per_group, ahead = opt["ring"], max(1, opt["group"]), opt["ahead"]
    sched = [(p[0], 0, ("pfl2", p[1])) for p in fn.prefetch]  # (position, order, statement)
    recs = sorted(fn.loads, key=lambda r: (r["at"], r["use"]))
    import heapq
    free = [(-1, k) for k in range(ring)]  # (position of the last use of the slot's occupant, slot)
    heapq.heapify(free)
    copies, last_at = [], -1
    for r in recs:
        r["slot"] = None
        if ring > 0:
            at = max(r["at"], last_at)
            freed, slot = free[0]
            at = max(at, freed + 1)
            if at <= r["use"] - ahead:
                heapq.heapreplace(free, (r["use"], slot))
                r["slot"], r["issue"], last_at = slot, at, at
                copies.append(r)
                continue
        sched.append((r["at"], 2, ("ldg", r["v"], r["e"])))
    # commit groups: closed after `group` copies, and before a copy that would be issued after a member's use
    groups, cur = [], []
    for r in copies:
        if cur and (len(cur) >= per_group or r["issue"] >= min(c["use"] for c in cur)):
            groups.append(cur)
            cur = []
        cur.append(r)
    if cur:
        groups.append(cur)
    commit_at = []
    for j, grp in enumerate(groups):
        for r in grp:
            r["group"] = j
            sched.append((r["issue"], 1, ("cpa", r["slot"], r["e"])))
        commit_at.append(grp[-1]["issue"])
        sched.append((grp[-1]["issue"], 1.5, ("commit",)))
        assert all(c["use"] > grp[-1]["issue"] for c in grp)
    sched.sort(key=lambda x: (x[0], x[1]))
    out, si, waited = [], 0, -1
    import bisect
    for pos, st in fn.items:
        while si < len(sched) and sched[si][0] <= pos:
            out.append(sched[si][2])
            si += 1
        if st[0] == "use":
            r = st[1]
            if r["slot"] is None:
                continue
            j = r["group"]
            if j > waited:
                committed = bisect.bisect_right(commit_at, pos)  # groups committed before the statements of `pos`
                assert committed > j
                out.append(("waitg", committed - 1 - j))
                waited = j
            out.append(("ldr", r["v"], r["slot"]))
        else:
            out.append(st)
    for x in sched[si:]:  # requests of a trailing micro-op without statements (d = P)
        assert x[2][0] in ("ldg", "pfl2")
        out.append(x[2])
    fn.final = out
    return out


def pressure(fn: Function):
    """Most doubles alive at once in the function's statement order (what the register file must hold)."""
    sts = statements(fn)
    last, born = {}, {}
    for i, st in enumerate(sts):
        k = st[0]
        if k in ("pfl2", "cpa", "commit", "waitg"):
            continue
        uses = {"ldg": (), "lds": (), "ldr": (), "stg": (st[2],), "sts": (st[2],), "msub": st[2:5], "div": st[2:4], "regu": (st[2],)}[k]
        for u in uses:
            if u is not None and u in born:
                last[u] = i
        if k in ("ldg", "lds", "ldr", "msub", "div", "regu"):
            born[st[1]] = i
    for v in (fn.t, fn.d):
        if v in born:
            last[v] = len(sts)
    events = []
    for v, b in born.items():
        events.append((b, 1))
        events.append((last.get(v, b) + 1, -1))
    live = most = 0
    for _, d in sorted(events):
        live += d
        most = max(most, live)
    return most


def regularise(d):
    """ldl.c:319-320"""
    sign = -1.0 if d <= 0.0 else 1.0
    return np.float64(sign * 0.0000001) if sign * d <= 0.00000000000001 else d


def evaluate(B: Builder, G, S):
    """Run a segment's IR on one HBM slice G and one window S (float64 arrays), in place."""
    f64 = np.float64
    t = d = f64(np.nan)
    with np.errstate(all="ignore"):
        for fn in B.functions:
            env = {"0.0": f64(0.0), "1.0": f64(1.0), "-1.0": f64(-1.0), "t_in": t, "d_in": d}
            ringv, inflight, committed = {}, [], 0
            for st in statements(fn):
                k = st[0]
                if k == "pfl2":
                    pass
                elif k == "cpa":  # lands when its group is waited for; the element must not change until then
                    inflight.append([st[1], st[2], f64(G[st[2]]), None])
                elif k == "commit":
                    for c in inflight:
                        if c[3] is None:
                            c[3] = committed
                    committed += 1
                elif k == "waitg":
                    done = committed - st[1]
                    for c in [c for c in inflight if c[3] is not None and c[3] < done]:
                        assert f64(G[c[1]]) == c[2] or (np.isnan(c[2]) and np.isnan(G[c[1]])), "element stored while its copy was in flight"
                        ringv[c[0]] = c[2]
                        inflight.remove(c)
                elif k == "ldr":
                    assert not any(c[0] == st[2] for c in inflight), "slot read before its copy was waited for"
                    env[st[1]] = ringv[st[2]]
                elif k == "ldg":
                    env[st[1]] = f64(G[st[2]])
                elif k == "lds":
                    env[st[1]] = f64(S[st[2]])
                elif k == "stg":
                    G[st[1]] = env[st[2]]
                elif k == "sts":
                    S[st[1]] = env[st[2]]
                elif k == "msub":
                    env[st[1]] = env[st[2]] - (env[st[3]] if st[4] is None else env[st[3]] * env[st[4]])
                elif k == "div":
                    env[st[1]] = env[st[2]] / env[st[3]]
                elif k == "regu":
                    env[st[1]] = regularise(env[st[2]])
                else:
                    raise ValueError(k)
            if fn.carried:
                t, d = env[fn.t], env[fn.d]


def print_cuda(B: Builder, out, div):
    """Every function takes `mine` (does this warp have work here?) and is entered by ALL warps of the
    CTA: they meet at a barrier at its top, so that they fetch the same instructions at about the same
    time (the code is far larger than the instruction caches), and a warp with nothing to do skips the
    work.  (Barriers INSIDE a function -- `if (mine) {...}` blocks -- were measured and lose 30 %: values
    that live across the blocks get spilled.)"""
    for fn in B.functions:
        if fn.carried:
            out.append(f"static __device__ __noinline__ double2 {fn.name}(char *g, const unsigned sb, const unsigned rb, const bool mine, const double t_in, const double d_in)\n{{")
        else:
            out.append(f"static __device__ __noinline__ void {fn.name}(char *g, const unsigned sb, const unsigned rb, const bool mine)\n{{")
        names = [l["v"] for l in fn.loads] + [st[1] for _, st in fn.items if st[0] in ("lds", "msub", "div", "regu")]
        for i in range(0, len(names), 16):
            out.append("\tdouble " + ", ".join(names[i:i + 16]) + ";")
        body = []

        def flush():
            if body:
                out.append("\t__syncthreads();\n\tif (mine) {")
                out.extend(body)
                out.append("\t}")
                body.clear()

        def emit(st):
            k = st[0]
            if k == "pfl2":
                body.append(f'\t\tasm volatile("prefetch.global.L2 [%0+{st[1] << 8}];" :: "l"(g));')
            elif k == "cpa":
                body.append(f'\t\tasm volatile("cp.async.ca.shared.global [%0+{st[1] << 8}], [%1+{st[2] << 8}], 8;" :: "r"(rb), "l"(g) : "memory");')
            elif k == "commit":
                body.append('\t\tasm volatile("cp.async.commit_group;" ::: "memory");')
            elif k == "waitg":
                body.append(f'\t\tasm volatile("cp.async.wait_group {st[1]};" ::: "memory");')
            elif k == "ldr":
                body.append(f'\t\tasm volatile("ld.shared.f64 %0, [%1+{st[2] << 8}];" : "=d"({st[1]}) : "r"(rb) : "memory");')
            elif k == "ldg":
                body.append(f'\t\tasm volatile("ld.global.f64 %0, [%1+{st[2] << 8}];" : "=d"({st[1]}) : "l"(g) : "memory");')
            elif k == "lds":
                body.append(f'\t\tasm volatile("ld.shared.f64 %0, [%1+{st[2] << 8}];" : "=d"({st[1]}) : "r"(sb));')
            elif k == "stg":
                body.append(f'\t\tasm volatile("st.global.f64 [%0+{st[1] << 8}], %1;" :: "l"(g), "d"({st[2]}) : "memory");')
            elif k == "sts":
                body.append(f'\t\tasm volatile("st.shared.f64 [%0+{st[1] << 8}], %1;" :: "r"(sb), "d"({st[2]}) : "memory");')
            elif k == "msub":
                body.append(f"\t\t{st[1]} = {st[2]} - {st[3]}{'' if st[4] is None else ' * ' + st[4]};")
            elif k == "div":
                body.append(f"\t\t{st[1]} = {div}({st[2]}, {st[3]});")
            elif k == "regu":
                body.append(f"\t\t{st[1]} = regularise({st[2]});")

        for st in statements(fn):
            emit(st)
        flush()
        if fn.carried:
            t = fn.t if fn.t[0].isdigit() or fn.t[0] == "-" or fn.t == "t_in" else f"mine ? {fn.t} : t_in"
            d = fn.d if fn.d[0].isdigit() or fn.d[0] == "-" or fn.d == "d_in" else f"mine ? {fn.d} : d_in"
            out.append(f"\treturn make_double2({t}, {d});")
        out.append("}\n")
    out.append(f"static __device__ __forceinline__ void {B.seg}(double *slice, const unsigned sb, const unsigned rb, const bool mine)\n{{")
    out.append("\tchar *g = reinterpret_cast<char *>(slice);\n\tdouble2 td = make_double2(0.0, 0.0);")
    for fn in B.functions:
        out.append(f"\ttd = {fn.name}(g, sb, rb, mine, td.x, td.y);" if fn.carried else f"\t{fn.name}(g, sb, rb, mine);")
    out.append("\t(void)td;\n}\n")


def build_all(plan, **options):
    """The four segments of a plan's micro-programs as IR: ({segment: Builder}, options)."""
    opt = dict(DEFAULTS)
    opt.update({k: v for k, v in options.items() if v is not None})
    info = plan.stream_info()
    assert info["ok"] == 1
    lay = info["layout"]
    consts = {lay["zero"]: "0.0", lay["one"]: "1.0", lay["mone"]: "-1.0"}
    return {seg: build_segment(seg, live_ops(plan.stream_words(seg), lay["zero"]), consts, opt, lay["scratch"])
            for seg in ("factor", "solve", "resid", "resz")}, opt


def generate(kind: str, div: str = "div_inline", **options):
    from . import dogbot, qpswift as Q
    plan = Q.Plan(dogbot.generate(kind, 1).pattern)
    key = plan.stream_key()
    segs, opt = build_all(plan, **options)
    out = [f"// GENERATED by apf_quadruped_b200/codegen.py for the '{kind}' pattern (micro-program key {key:#018x}); do not edit.",
           f"// options: {opt}",
           '#include "stream.cuh"', '#include "stream_gen.cuh"', "", "namespace qpb {", f"namespace gen_{kind} {{", ""]
    for seg in ("factor", "solve", "resid", "resz"):
        print_cuda(segs[seg], out, div)
    out += ["struct Prog", "{", "\tstatic constexpr bool cta_sync = true;", f"\tstatic constexpr int ring_slots = {opt['ring']};"]
    for seg in ("factor", "solve", "resid", "resz"):
        out.append(f"\tstatic __device__ __forceinline__ void {seg}(const StreamView &, double *slice, const unsigned sb, const unsigned rb, const bool mine) {{ gen_{kind}::{seg}(slice, sb, rb, mine); }}")
    out += ["};", "", f"}} // namespace gen_{kind}", "",
            f"QPB_REGISTER_GENERATED({kind}, {key:#018x}ull, gen_{kind}::Prog)", "", "} // namespace qpb", ""]
    plan.close()
    for b in segs.values():
        b.stats["max_live"] = max(pressure(fn) for fn in b.functions)
        kinds = [st[0] for fn in b.functions for st in statements(fn)]
        b.stats.update(ring_loads=kinds.count("cpa"), direct_loads=kinds.count("ldg"), waits=kinds.count("waitg"))
    return "\n".join(out), {seg: b.stats for seg, b in segs.items()}


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("kinds", nargs="*", default=["trot", "stance", "crawl"])
    ap.add_argument("--out", default="build/gen")
    ap.add_argument("--div", default="div_inline", help="div_rn: out-of-line division; div_inline: inlined")
    for k, v in DEFAULTS.items():
        ap.add_argument(f"--{k}", type=int, default=None, help=f"default {v}")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for kind in a.kinds:
        src, stats = generate(kind, a.div, **{k: getattr(a, k) for k in DEFAULTS})
        path = os.path.join(a.out, f"stream_gen_{kind}.cu")
        old = open(path).read() if os.path.exists(path) else None
        if old != src:  # keep the time stamp when nothing changed (make)
            with open(path, "w") as f:
                f.write(src)
        print(f"{path}:")
        for seg, st in stats.items():
            print(f"  {seg}: {st}")


if __name__ == "__main__":
    main()
