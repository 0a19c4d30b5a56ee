"""The DogBot whole-body-control QP from the robot's dynamic quantities and the joint torques from its
solution (8 f1): host-side mirror of ``csrc/wbc.cuh``.

``dogbot_controller/src/client/main.cpp`` formulates one dense QP per control tick from iDynTree's mass matrix, Jacobians
and bias forces (stance 1471-1647, trot 1734-1992, crawl 2943-3230), solves it and turns the solution into joint torques
(1667 / 2023 / 3246).  Here a robot is ONE packed state row of ``STATE`` = 468 doubles

    Jc[12][6] | Jj[12][12] | Mcom[6][6] | Mjj[12][12] | bias[18] | Jdqd[12] | W[6] | q[12] | dq[12] | vdot_sw[12]

(feet br, bl, fl, fr; row-major blocks).  Two things live in this module:

* ``formulate_reference`` / ``torques_reference`` -- the numpy RESTATEMENT of the controller's formulation with every
  operation order written out (sums from 0.0 in ascending index order, no fused multiply-add): dense Q, A, G, c, h, b
  exactly as ``QP_SETUP_dense`` would receive them, and the plan-ordered value rows.  It is what the device kernels are
  checked against, bit for bit (``tests/test_wbc.py``, ``tests/test_wbc_gpu.py``), and it runs without a GPU.
* ``Formulator`` -- ctypes access to ``QPB_WBC_FORMULATE_DEVICE`` / ``QPB_WBC_TORQUES_DEVICE`` (device pointers in, device
  pointers out; there is no CPU fallback).

``random_states`` draws synthetic states with the recipe of ``dogbot._Physics``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import dogbot

KINDS = {"stance": 0, "trot": 1, "crawl": 2}
STANCE_FEET = {"stance": [0, 1, 2, 3], "trot": [1, 3], "crawl": [1, 2, 3]}  # of br, bl, fl, fr (main.cpp:1734-1741, 2943-2950)
JC, JJ, MCOM, MJJ, BIAS, JDQD, W, QJ, DQJ, VSW, STATE = 0, 72, 216, 252, 396, 414, 426, 432, 444, 456, 468
SLACK_WEIGHT = {"stance": 0.0, "trot": 1e8, "crawl": 1e4}  # main.cpp:1751 / 2976


def pack_states(Jc, Jj, M, bias, Jdqd, Wd, q, dq, vdot_sw) -> np.ndarray:
    """State rows [B, 468] from the controller's quantities (M is the full 18 x 18 mass matrix)."""
    B = Jc.shape[0]
    S = np.empty((B, STATE))
    S[:, JC:JJ] = Jc.reshape(B, 72)
    S[:, JJ:MCOM] = Jj.reshape(B, 144)
    S[:, MCOM:MJJ] = M[:, 0:6, 0:6].reshape(B, 36)
    S[:, MJJ:BIAS] = M[:, 6:18, 6:18].reshape(B, 144)
    S[:, BIAS:JDQD] = bias
    S[:, JDQD:W] = Jdqd
    S[:, W:QJ] = Wd
    S[:, QJ:DQJ] = q
    S[:, DQJ:VSW] = dq
    S[:, VSW:STATE] = vdot_sw
    return S


def random_states(count: int, start: int = 0, seed: int = dogbot.DEFAULT_SEED + 77) -> np.ndarray:
    """Synthetic robot states (the draws of ``dogbot._Physics``), reproducible per instance index."""
    out = np.empty((count, STATE))
    done = 0
    while done < count:
        i = start + done
        chunk, off = divmod(i, dogbot.CHUNK)
        take = min(dogbot.CHUNK - off, count - done)
        rng = np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFFFFFFFFFF, chunk]))
        ph = dogbot._Physics(rng, dogbot.CHUNK)
        # (_Physics keeps the acceleration limits, not q and dq: fresh draws from the same stream)
        q = 0.5 * (dogbot.QMIN + dogbot.QMAX) + rng.uniform(-0.3, 0.3, size=(dogbot.CHUNK, 12))
        dq = rng.uniform(-0.5, 0.5, size=(dogbot.CHUNK, 12))
        S = pack_states(ph.Jc, ph.Jj, ph.M, ph.bias, ph.Jdqd, ph.W, q, dq, ph.vdot_sw)
        out[done:done + take] = S[off:off + take]
        done += take
    return out


def _dot6(S, a, b):
    """sum_k S[a+k] S[b+k], k = 0..5, from 0.0 in ascending k (what wbc_eval does)."""
    acc = np.zeros(S.shape[0])
    for k in range(6):
        acc = acc + S[:, a + k] * S[:, b + k]
    return acc


def formulate_reference(kind: str, S: np.ndarray):
    """Dense Q [B,30,30], A [B,p,30], G [B,m,30], c, h, b of ``kind`` for the state rows S, every entry by the same
    sequence of IEEE-754 operations as the device kernel.  Structure: main.cpp:1476-1647 (stance), 1744-1992 (trot),
    2960-3230 (crawl)."""
    S = np.ascontiguousarray(S, dtype=np.float64)
    B = S.shape[0]
    feet = STANCE_FEET[kind]
    swing = [f for f in range(4) if f not in feet]
    st = [3 * f + d for f in feet for d in range(3)]
    sw = [3 * f + d for f in swing for d in range(3)]
    nf, nslack, n = len(st), len(sw), 30
    p = 6 + nf
    m = 5 * len(feet) + 24 + 2 * nslack + 24
    Q = np.zeros((B, n, n))
    A = np.zeros((B, p, n))
    G = np.zeros((B, m, n))
    c = np.zeros((B, n))
    h = np.zeros((B, m))
    b = np.zeros((B, p))
    # cost: Q = R + 50 Ts'Ts with Ts = Jst_c' Sigma (1476-1481); R = I, slack weights on the last 3 nw variables (1751, 2976)
    for i in range(18 + nf):
        Q[:, i, i] = 1.0
    for i in range(18 + nf, 30):
        Q[:, i, i] = SLACK_WEIGHT[kind]
    for i in range(nf):
        for j in range(nf):
            Q[:, 18 + i, 18 + j] = (1.0 if i == j else 0.0) + 50.0 * _dot6(S, JC + 6 * st[i], JC + 6 * st[j])
        c[:, 18 + i] = -50.0 * _dot6(S, JC + 6 * st[i], W)  # c = -Ts' 50 W (1573)
    # equalities [Mcom 0 -Jst_c'; Jst_c Jst_j 0] (1579-1583), b = [-bias_com; -Jdqd_st] (1586-1588; never filled for trot, 1851)
    for r in range(6):
        A[:, r, 0:6] = S[:, MCOM + 6 * r:MCOM + 6 * r + 6]
        for a in range(nf):
            A[:, r, 18 + a] = -S[:, JC + 6 * st[a] + r]
    for a in range(nf):
        A[:, 6 + a, 0:6] = S[:, JC + 6 * st[a]:JC + 6 * st[a] + 6]
        A[:, 6 + a, 6:18] = S[:, JJ + 12 * st[a]:JJ + 12 * st[a] + 12]
    if kind != "trot":
        b[:, 0:6] = -S[:, BIAS:BIAS + 6]
        for a in range(nf):
            b[:, 6 + a] = -S[:, JDQD + st[a]]
    # inequalities
    r0 = 0
    cfr = dogbot._CFR
    for k in range(len(feet)):  # friction pyramids (1612-1632)
        G[:, r0:r0 + 5, 18 + 3 * k:18 + 3 * k + 3] = cfr
        r0 += 5
    for j in range(12):  # torque limits +-[0 Mjj -Jst_j'] (1594-1597), h = tau_max - bias, -(tau_min - bias) (1634-1635)
        G[:, r0 + j, 6:18] = S[:, MJJ + 12 * j:MJJ + 12 * j + 12]
        G[:, r0 + 12 + j, 6:18] = -S[:, MJJ + 12 * j:MJJ + 12 * j + 12]
        for a in range(nf):
            G[:, r0 + j, 18 + a] = -S[:, JJ + 12 * st[a] + j]
            G[:, r0 + 12 + j, 18 + a] = S[:, JJ + 12 * st[a] + j]
        h[:, r0 + j] = dogbot.TAU_MAX - S[:, BIAS + 6 + j]
        h[:, r0 + 12 + j] = dogbot.TAU_MAX + S[:, BIAS + 6 + j]
    r0 += 24
    for half, neg in ((0, False), (1, True)):  # swing-foot tracking with slack (1864-1875, 1991-1992)
        for k in range(nslack):
            r = r0 + half * nslack + k
            rowc = S[:, JC + 6 * sw[k]:JC + 6 * sw[k] + 6]
            rowj = S[:, JJ + 12 * sw[k]:JJ + 12 * sw[k] + 12]
            G[:, r, 0:6] = -rowc if neg else rowc
            G[:, r, 6:18] = -rowj if neg else rowj
            G[:, r, 18 + nf + k] = -1.0
            d = S[:, VSW + sw[k]] - S[:, JDQD + sw[k]]
            h[:, r] = -d if neg else d
    r0 += 2 * nslack
    kk = 2.0 / (dogbot.DT_LIMITS * dogbot.DT_LIMITS)
    for j in range(12):  # joint acceleration limits (1598-1599, 1642-1646)
        G[:, r0 + j, 6 + j] = 1.0
        G[:, r0 + 12 + j, 6 + j] = -1.0
        h[:, r0 + j] = kk * ((dogbot.QMAX[j] - S[:, QJ + j]) - dogbot.DT_LIMITS * S[:, DQJ + j])
        h[:, r0 + 12 + j] = -(kk * ((dogbot.QMIN[j] - S[:, QJ + j]) - dogbot.DT_LIMITS * S[:, DQJ + j]))
    assert r0 + 24 == m
    return Q, A, G, c, h, b


def value_rows(kind: str, S: np.ndarray) -> dogbot.Batch:
    """The plan-ordered value rows (``dogbot.Batch``) of the reference formulation."""
    Qd, A, G, c, h, b = formulate_reference(kind, S)
    pat = dogbot.pattern(kind)
    _, Qm, Am, Gm = dogbot._PATTERN_CACHE[kind]
    sel = [np.nonzero(M.T)[::-1] for M in (Qm, Am, Gm)]  # (rows, cols) in CCS order
    rows = [np.ascontiguousarray(M[:, rr, cc]) for M, (rr, cc) in zip((Qd, A, G), sel)]
    return dogbot.Batch(pat, rows[0], rows[1], rows[2], c, h, b)


def torques_reference(kind: str, S: np.ndarray, x: np.ndarray) -> np.ndarray:
    """tau = Mjj qdd + bias_j - Jst_j' f (main.cpp:1667 / 2023 / 3246), sums from 0.0 in ascending index order."""
    st = [3 * f + d for f in STANCE_FEET[kind] for d in range(3)]
    B = S.shape[0]
    tau = np.empty((B, 12))
    for j in range(12):
        a = np.zeros(B)
        for k in range(12):
            a = a + S[:, MJJ + 12 * j + k] * x[:, 6 + k]
        bb = np.zeros(B)
        for k in range(len(st)):
            bb = bb + S[:, JJ + 12 * st[k] + j] * x[:, 18 + k]
        tau[:, j] = (a + S[:, BIAS + 6 + j]) - bb
    return tau


def descriptors(plan, kind: str):
    """The library's formula table for ``kind`` on ``plan``'s pattern (host only): (op, flag, a, b) arrays + constants."""
    from . import qpswift
    lb = qpswift.lib()
    pat = plan.pattern
    cnt = pat.nnzP + pat.nnzA + pat.nnzG + pat.n + pat.m + pat.p
    raw = np.zeros(cnt, dtype=np.uint64)
    consts = np.zeros(32)
    rc = lb.QPB_WBC_DESCRIPTORS(plan.handle, KINDS[kind], raw.ctypes.data_as(C.POINTER(C.c_uint64)), cnt, consts.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != cnt:
        raise RuntimeError(f"QPB_WBC_DESCRIPTORS: {rc} ({lb.QPB_LAST_ERROR().decode()})")
    lo = (raw & np.uint64(0xFFFFFFFF)).astype(np.int64)
    return {"op": lo & 15, "flag": (lo >> 4) & 1, "a": lo >> 8, "b": (raw >> np.uint64(32)).astype(np.int64), "consts": consts}


def evaluate_descriptors(d, S: np.ndarray) -> np.ndarray:
    """What wbc_formulate_kernel computes from the table, in numpy with the kernel's operation order (csrc/wbc.cuh
    wbc_eval): [B, entries] in the order Ppr | Apr | Gpr | c | h | b."""
    B = S.shape[0]
    X = np.concatenate([S, np.broadcast_to(d["consts"], (B, 32))], axis=1)
    out = np.empty((B, len(d["op"])))
    fifty, kk, dt = X[:, 473], X[:, 474], X[:, 475]
    for e, (op, fl, a, b) in enumerate(zip(d["op"], d["flag"], d["a"], d["b"])):
        if op == 0:
            v = X[:, a]
            v = -v if fl else v
        elif op in (1, 2):
            acc = np.zeros(B)
            for k in range(6):
                acc = acc + X[:, a + k] * X[:, b + k]
            v = (1.0 if fl else 0.0) + fifty * acc if op == 1 else -fifty * acc
        elif op == 3:
            v = X[:, a] - X[:, b]
            v = -v if fl else v
        elif op == 4:
            v = X[:, a] + X[:, b]
        else:
            v = kk * ((X[:, a] - X[:, QJ + b]) - dt * X[:, DQJ + b])
            v = -v if fl else v
        out[:, e] = v
    return out


class Formulator:
    """``QPB_WBC_FORMULATE_DEVICE`` / ``QPB_WBC_TORQUES_DEVICE`` on a ``qpswift.BatchSolver`` of the kind's pattern."""

    def __init__(self, solver, kind: str):
        from . import qpswift
        self.lib, self.solver, self.kind = qpswift.lib(), solver, KINDS[kind]
        V = C.c_void_p
        self.lib.QPB_WBC_FORMULATE_DEVICE.restype = C.c_long
        self.lib.QPB_WBC_FORMULATE_DEVICE.argtypes = [V, C.c_int, C.c_long] + [V] * 10
        self.lib.QPB_WBC_TORQUES_DEVICE.restype = C.c_long
        self.lib.QPB_WBC_TORQUES_DEVICE.argtypes = [V, C.c_int, C.c_long] + [V] * 6

    def formulate(self, count, dstate, dPpr, dApr, dGpr, dc, dh, db, dindex=None, dcount=None, stream=None):
        rc = self.lib.QPB_WBC_FORMULATE_DEVICE(self.solver.handle, self.kind, int(count), dstate, dindex, dcount, dPpr, dApr, dGpr, dc, dh, db, stream)
        if rc < 0:
            raise RuntimeError(f"QPB_WBC_FORMULATE_DEVICE failed ({rc}): {self.lib.QPB_LAST_ERROR().decode()}")

    def torques(self, count, dstate, dx, dtau, dindex=None, dcount=None, stream=None):
        rc = self.lib.QPB_WBC_TORQUES_DEVICE(self.solver.handle, self.kind, int(count), dstate, dx, dtau, dindex, dcount, stream)
        if rc < 0:
            raise RuntimeError(f"QPB_WBC_TORQUES_DEVICE failed ({rc}): {self.lib.QPB_LAST_ERROR().decode()}")
