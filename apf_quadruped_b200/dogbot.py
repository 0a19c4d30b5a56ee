"""Synthetic DogBot whole-body-control QP batches (host side, numpy).

The reference builds one dense QP per 400 Hz control tick from iDynTree quantities
(``dogbot_controller/src/client/main.cpp``) and hands it to ``QP_SETUP_dense``.  iDynTree, ROS
and Gazebo are not available, so -- as SURVEY.md section 8d prescribes -- the batches used by the
tests and by ``bench.py`` are synthetic draws with the *shape* of those QPs:

=========  ===  ===  ===  =====================  ==========================================
pattern     n    m    p   nnz P / A / G          formulated at (main.cpp)
=========  ===  ===  ===  =====================  ==========================================
stance      30   68   18  162 / 324 / 636        1471-1649  (4 feet in contact)
crawl       30   69   15  102 / 252 / 669        2943-3232  (3 feet, 3 slack variables)
trot        30   70   12   60 / 180 / 702        1734-2005  (2 feet, 6 slack variables)
mpc        240  240  120  240 / 660 / 400        not in the reference (ours): horizon-10 SRB MPC
=========  ===  ===  ===  =====================  ==========================================

Every block that the controller fills from the mass matrix / Jacobians is structurally dense,
every ``Zero()`` / ``Identity()`` block is structural.  The pattern is therefore fixed per kind and
only the values vary across the batch, which is what a shared symbolic plan needs.

Instances are reproducible and independent of how a batch is sharded: instance ``i`` of a kind is
drawn from ``Philox(key=(seed, i // CHUNK))`` at offset ``i % CHUNK``.

This module is data generation only; it never solves anything.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

CHUNK = 1024
DEFAULT_SEED = 0xD06B07

ROBOT_MASS = 21.27  # DogBotV4 dogbot.urdf total mass
GRAVITY = 9.81
X_NOM, Y_NOM = 0.186571, 0.289186  # main.cpp:433-434
QMIN = np.array([-1.75, -1.75, -1.75, -1.75, -1.58, -2.62, -3.15, -0.02, -1.58, -2.62, -3.15, -0.02])  # main.cpp:493
QMAX = np.array([1.75, 1.75, 1.75, 1.75, 3.15, 0.02, 1.58, 2.62, 3.15, 0.02, 1.58, 2.62])  # main.cpp:494
DT_LIMITS = 0.025  # main.cpp:1638
TAU_MAX = 60.0  # main.cpp:1630
MU = 0.5  # main.cpp:1511


@dataclass
class Pattern:
    """CCS sparsity pattern shared by a whole batch (what ``QP_SETUP`` takes, minus values)."""

    kind: str
    n: int
    m: int
    p: int
    Pjc: np.ndarray
    Pir: np.ndarray
    Ajc: np.ndarray
    Air: np.ndarray
    Gjc: np.ndarray
    Gir: np.ndarray

    @property
    def nnzP(self) -> int:
        return int(self.Pjc[-1])

    @property
    def nnzA(self) -> int:
        return int(self.Ajc[-1])

    @property
    def nnzG(self) -> int:
        return int(self.Gjc[-1])

    @property
    def N(self) -> int:
        return self.n + self.m + self.p


@dataclass
class Batch:
    """Value rows of ``count`` QPs sharing ``pattern`` (C-contiguous float64, row per QP)."""

    pattern: Pattern
    Ppr: np.ndarray
    Apr: np.ndarray
    Gpr: np.ndarray
    c: np.ndarray
    h: np.ndarray
    b: np.ndarray

    @property
    def count(self) -> int:
        return self.c.shape[0]

    def slice(self, lo: int, hi: int) -> "Batch":
        return Batch(self.pattern, self.Ppr[lo:hi], self.Apr[lo:hi], self.Gpr[lo:hi], self.c[lo:hi], self.h[lo:hi], self.b[lo:hi])

    def dense(self, i: int):
        """Dense column-major copies of instance ``i`` (what ``QP_SETUP_dense`` would be given)."""
        pat = self.pattern

        def fill(rows, jc, ir, vals):
            M = np.zeros((rows, pat.n), order="F")
            for col in range(pat.n):
                M[ir[jc[col]:jc[col + 1]], col] = vals[jc[col]:jc[col + 1]]
            return M

        return (fill(pat.n, pat.Pjc, pat.Pir, self.Ppr[i]), fill(pat.p, pat.Ajc, pat.Air, self.Apr[i]),
                fill(pat.m, pat.Gjc, pat.Gir, self.Gpr[i]))

    def dense_batch(self, row_major: bool = False):
        """The whole batch as dense blocks P [count, n*n], A [count, p*n], G [count, m*n], each block column-major
        (or row-major): the input of ``QP_SOLVE_BATCH_DENSE``."""
        pat = self.pattern

        def fill(rows, jc, ir, vals):
            jc, ir = np.asarray(jc), np.asarray(ir)
            cols = np.repeat(np.arange(pat.n), np.diff(jc))
            cell = ir * pat.n + cols if row_major else cols * rows + ir
            M = np.zeros((self.count, rows * pat.n))
            if cell.size:
                M[:, cell] = vals
            return M

        return (fill(pat.n, pat.Pjc, pat.Pir, self.Ppr), fill(pat.p, pat.Ajc, pat.Air, self.Apr),
                fill(pat.m, pat.Gjc, pat.Gir, self.Gpr))


def _ccs_from_mask(mask: np.ndarray):
    cols, rows = np.nonzero(mask.T)  # column-major order of the structural entries
    jc = np.zeros(mask.shape[1] + 1, dtype=np.int64)
    np.add.at(jc, cols + 1, 1)
    return np.cumsum(jc), rows.astype(np.int64), cols.astype(np.int64)


def _skew(r: np.ndarray) -> np.ndarray:
    S = np.zeros(r.shape[:-1] + (3, 3))
    S[..., 0, 1], S[..., 0, 2] = -r[..., 2], r[..., 1]
    S[..., 1, 0], S[..., 1, 2] = r[..., 2], -r[..., 0]
    S[..., 2, 0], S[..., 2, 1] = -r[..., 1], r[..., 0]
    return S


_CFR = np.array([[1.0, 0.0, -MU], [0.0, 1.0, -MU], [-1.0, 0.0, -MU], [0.0, -1.0, -MU], [0.0, 0.0, -1.0]])  # main.cpp:1612-1625


class _Physics:
    """Random robot state for ``B`` instances (the quantities the controller gets from iDynTree)."""

    def __init__(self, rng: np.random.Generator, B: int):
        u = lambda lo, hi, *shape: rng.uniform(lo, hi, size=(B,) + shape)
        signs = np.array([[-1, -1], [-1, 1], [1, 1], [1, -1]], dtype=float)  # br, bl, fl, fr
        r = np.empty((B, 4, 3))
        r[:, :, 0] = signs[:, 0] * X_NOM
        r[:, :, 1] = signs[:, 1] * Y_NOM
        r[:, :, 2] = -0.40
        r += u(-0.03, 0.03, 4, 3)
        Jc = np.zeros((B, 12, 6))
        for f in range(4):
            Jc[:, 3 * f:3 * f + 3, 0:3] = np.eye(3)
            Jc[:, 3 * f:3 * f + 3, 3:6] = -_skew(r[:, f])
        Jc += u(-1e-3, 1e-3, 12, 6)
        Jj = u(-0.01, 0.01, 12, 12)
        for f in range(4):
            Jj[:, 3 * f:3 * f + 3, 3 * f:3 * f + 3] = u(-0.35, 0.35, 3, 3)
        Bm = u(-0.05, 0.05, 18, 18)
        diag = np.concatenate([[ROBOT_MASS] * 3, [0.3, 0.8, 0.9], [0.08] * 12])
        M = Bm @ np.swapaxes(Bm, 1, 2) + np.diag(diag)
        self.Jc, self.Jj, self.M = Jc, Jj, M
        self.bias = u(-0.5, 0.5, 18)
        self.Jdqd = u(-0.05, 0.05, 12)
        dx = u(-0.01, 0.01, 6)
        dv = u(-0.05, 0.05, 6)
        g = np.zeros(6)
        g[2] = GRAVITY
        self.W = 3000.0 * dx + 50.0 * dv + ROBOT_MASS * g  # main.cpp:1572 (desired CoM acceleration = 0)
        q = 0.5 * (QMIN + QMAX) + u(-0.3, 0.3, 12)
        dq = u(-0.5, 0.5, 12)
        k = 2.0 / DT_LIMITS ** 2
        self.ddqmin = k * (QMIN - q - DT_LIMITS * dq)  # main.cpp:1642-1643
        self.ddqmax = k * (QMAX - q - DT_LIMITS * dq)
        self.vdot_sw = u(-2.0, 2.0, 12)  # desired swing-foot accelerations (Kp, Kd terms folded in)


def _wbc(kind: str, ph: _Physics):
    """Dense Q, A, G, c, h, b and their structural masks for stance / crawl / trot."""
    B = ph.M.shape[0]
    stance_feet = {"stance": [0, 1, 2, 3], "crawl": [1, 2, 3], "trot": [1, 3]}[kind]
    swing_feet = [f for f in range(4) if f not in stance_feet]
    ns, nw = len(stance_feet), len(swing_feet)
    nf = 3 * ns                     # force variables
    nslack = 3 * nw                 # slack variables (swing-foot tracking)
    n = 30
    p = 6 + nf
    m = 5 * ns + 24 + 2 * nslack + 24
    st_rows = np.concatenate([np.arange(3 * f, 3 * f + 3) for f in stance_feet])
    Jst_c, Jst_j = ph.Jc[:, st_rows, :], ph.Jj[:, st_rows, :]
    Mcom, Mjj = ph.M[:, 0:6, 0:6], ph.M[:, 6:18, 6:18]

    # cost (main.cpp:1476-1481 / 1744-1754 / 2960-2966): Q = R + 50 Ts'Ts, Ts = Jst_c' Sigma
    Q = np.zeros((B, n, n))
    Qm = np.zeros((n, n), dtype=bool)
    idx = np.arange(n)
    Q[:, idx, idx] = 1.0
    Qm[idx, idx] = True
    if nslack:
        w = 1e8 if kind == "trot" else 1e4  # main.cpp:1751 / 2976
        sl = np.arange(18 + nf, 30)
        Q[:, sl, sl] = w
    Q[:, 18:18 + nf, 18:18 + nf] += 50.0 * (Jst_c @ np.swapaxes(Jst_c, 1, 2))
    Qm[18:18 + nf, 18:18 + nf] = True
    c = np.zeros((B, n))
    c[:, 18:18 + nf] = -50.0 * np.einsum("bij,bj->bi", Jst_c, ph.W)  # main.cpp:1573

    # equalities (main.cpp:1579-1588): [Mcom 0 -Jst_c'; Jst_c Jst_j 0]
    A = np.zeros((B, p, n))
    Am = np.zeros((p, n), dtype=bool)
    A[:, 0:6, 0:6] = Mcom
    Am[0:6, 0:6] = True
    A[:, 0:6, 18:18 + nf] = -np.swapaxes(Jst_c, 1, 2)
    Am[0:6, 18:18 + nf] = True
    A[:, 6:p, 0:6] = Jst_c
    Am[6:p, 0:6] = True
    A[:, 6:p, 6:18] = Jst_j
    Am[6:p, 6:18] = True
    b = np.zeros((B, p))
    if kind != "trot":  # the trot branch never fills eigenb (main.cpp:1851)
        b[:, 0:6] = -ph.bias[:, 0:6]
        b[:, 6:p] = -ph.Jdqd[:, st_rows]

    # inequalities
    G = np.zeros((B, m, n))
    Gm = np.zeros((m, n), dtype=bool)
    h = np.zeros((B, m))
    r0 = 0
    for k in range(ns):  # friction pyramids (main.cpp:1612-1632)
        G[:, r0:r0 + 5, 18 + 3 * k:18 + 3 * k + 3] = _CFR
        Gm[r0:r0 + 5, 18 + 3 * k:18 + 3 * k + 3] = _CFR != 0
        r0 += 5
    G[:, r0:r0 + 12, 6:18] = Mjj  # torque limits (main.cpp:1594-1597, 1634-1636)
    G[:, r0:r0 + 12, 18:18 + nf] = -np.swapaxes(Jst_j, 1, 2)
    G[:, r0 + 12:r0 + 24, 6:18] = -Mjj
    G[:, r0 + 12:r0 + 24, 18:18 + nf] = np.swapaxes(Jst_j, 1, 2)
    Gm[r0:r0 + 24, 6:18 + nf] = True
    h[:, r0:r0 + 12] = TAU_MAX - ph.bias[:, 6:18]
    h[:, r0 + 12:r0 + 24] = TAU_MAX + ph.bias[:, 6:18]
    r0 += 24
    if nslack:  # swing-foot tracking with slack (main.cpp:1864-1875, 1991-1992)
        sw_rows = np.concatenate([np.arange(3 * f, 3 * f + 3) for f in swing_feet])
        Jsw = np.concatenate([ph.Jc[:, sw_rows, :], ph.Jj[:, sw_rows, :]], axis=2)
        rhs = ph.vdot_sw[:, sw_rows] - ph.Jdqd[:, sw_rows]
        sl = np.arange(nslack)
        for sgn in (1.0, -1.0):
            G[:, r0:r0 + nslack, 0:18] = sgn * Jsw
            Gm[r0:r0 + nslack, 0:18] = True
            G[:, r0 + sl, 18 + nf + sl] = -1.0
            Gm[r0 + sl, 18 + nf + sl] = True
            h[:, r0:r0 + nslack] = sgn * rhs
            r0 += nslack
    j = np.arange(12)  # joint acceleration limits (main.cpp:1598-1599, 1645-1646)
    G[:, r0 + j, 6 + j] = 1.0
    Gm[r0 + j, 6 + j] = True
    h[:, r0:r0 + 12] = ph.ddqmax
    G[:, r0 + 12 + j, 6 + j] = -1.0
    Gm[r0 + 12 + j, 6 + j] = True
    h[:, r0 + 12:r0 + 24] = -ph.ddqmin
    r0 += 24
    assert r0 == m
    return (n, m, p), (Q, Qm), (A, Am), (G, Gm), c, h, b


def _mpc(rng: np.random.Generator, B: int):
    """Horizon-10 single-rigid-body MPC QP (ours).  z = [u0,x1,...,u9,x10]."""
    H, dt, fmax, mass = 10, 0.03, 180.0, ROBOT_MASS
    n, p, m = 24 * H, 12 * H, 24 * H
    u = lambda lo, hi, *shape: rng.uniform(lo, hi, size=(B,) + shape)
    Iinv = 1.0 / np.array([0.3, 0.8, 0.9])
    x0 = np.concatenate([u(-0.05, 0.05, 3), u(-0.02, 0.02, 2), 0.4 + u(-0.02, 0.02, 1), u(-0.1, 0.1, 3), u(-0.2, 0.2, 3)], axis=1)
    yaw = x0[:, 2:3] + np.cumsum(u(-0.02, 0.02, H), axis=1)
    signs = np.array([[-1, -1], [-1, 1], [1, 1], [1, -1]], dtype=float)
    r = np.empty((B, H, 4, 3))
    r[..., 0] = signs[:, 0] * X_NOM
    r[..., 1] = signs[:, 1] * Y_NOM
    r[..., 2] = -0.40
    r += u(-0.03, 0.03, H, 4, 3)
    phase = rng.integers(0, 2, size=(B, 1))
    swing = np.zeros((B, H, 4), dtype=bool)  # trot-like schedule: diagonal pairs alternate every 5 steps
    for k in range(H):
        first = ((k // 5) % 2 == 0)
        pair = np.where(phase == (0 if first else 1), 0, 1)[:, 0]
        swing[:, k, 0] = swing[:, k, 2] = pair == 0
        swing[:, k, 1] = swing[:, k, 3] = pair == 1

    P = np.zeros((B, n, n))
    Pm = np.zeros((n, n), dtype=bool)
    A = np.zeros((B, p, n))
    Am = np.zeros((p, n), dtype=bool)
    G = np.zeros((B, m, n))
    Gm = np.zeros((m, n), dtype=bool)
    c = np.zeros((B, n))
    h = np.zeros((B, m))
    b = np.zeros((B, p))
    Qw = np.array([50.0, 50.0, 20.0, 10.0, 10.0, 100.0, 1.0, 1.0, 1.0, 2.0, 2.0, 2.0])
    xref = np.zeros(12)
    xref[5] = 0.4
    e12 = np.arange(12)
    for k in range(H):
        uo, xo = 24 * k, 24 * k + 12  # u_k, x_{k+1}
        P[:, uo + e12, uo + e12] = 1e-4
        P[:, xo + e12, xo + e12] = Qw + 1e-6
        Pm[uo + e12, uo + e12] = Pm[xo + e12, xo + e12] = True
        c[:, xo:xo + 12] = -(Qw + 1e-6) * (xref + u(-0.01, 0.01, 12))
        ro = 12 * k
        A[:, ro + e12, xo + e12] = 1.0
        Am[ro + e12, xo + e12] = True
        cy, sy = np.cos(yaw[:, k]), np.sin(yaw[:, k])
        Rz = np.zeros((B, 3, 3))
        Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = cy, sy, -sy, cy, 1.0
        Rm = np.zeros((3, 3), dtype=bool)
        Rm[0:2, 0:2] = True
        Rm[2, 2] = True
        b[:, ro + 11] = -GRAVITY * dt
        if k == 0:  # x_0 is data: its terms move to the right-hand side
            b[:, ro:ro + 12] += x0
            b[:, ro:ro + 3] += dt * np.einsum("bij,bj->bi", Rz, x0[:, 6:9])
            b[:, ro + 3:ro + 6] += dt * x0[:, 9:12]
        else:
            xp = 24 * (k - 1) + 12
            A[:, ro + e12, xp + e12] = -1.0
            Am[ro + e12, xp + e12] = True
            A[:, ro:ro + 3, xp + 6:xp + 9] = -dt * Rz
            Am[ro:ro + 3, xp + 6:xp + 9] = Rm
            d3 = np.arange(3)
            A[:, ro + 3 + d3, xp + 9 + d3] = -dt
            Am[ro + 3 + d3, xp + 9 + d3] = True
        for f in range(4):
            fo = uo + 3 * f
            A[:, ro + 6:ro + 9, fo:fo + 3] = -dt * Iinv[:, None] * _skew(r[:, k, f])
            Am[ro + 6:ro + 9, fo:fo + 3] = ~np.eye(3, dtype=bool)
            d3 = np.arange(3)
            A[:, ro + 9 + d3, fo + d3] = -dt / mass
            Am[ro + 9 + d3, fo + d3] = True
            go = 24 * k + 6 * f
            pyr = np.array([[1.0, 0, -MU], [-1.0, 0, -MU], [0, 1.0, -MU], [0, -1.0, -MU], [0, 0, -1.0], [0, 0, 1.0]])
            G[:, go:go + 6, fo:fo + 3] = pyr
            Gm[go:go + 6, fo:fo + 3] = pyr != 0
            h[:, go + 5] = np.where(swing[:, k, f], 1e-3, fmax)
    return (n, m, p), (P, Pm), (A, Am), (G, Gm), c, h, b


_PATTERN_CACHE: dict = {}


def _build(kind: str, rng: np.random.Generator, B: int):
    if kind == "mpc":
        return _mpc(rng, B)
    return _wbc(kind, _Physics(rng, B))


def pattern(kind: str) -> Pattern:
    """The shared CCS pattern of ``kind`` ('stance' | 'crawl' | 'trot' | 'mpc')."""
    if kind not in _PATTERN_CACHE:
        (n, m, p), (_, Qm), (_, Am), (_, Gm), *_ = _build(kind, np.random.default_rng(0), 1)
        Pjc, Pir, _ = _ccs_from_mask(Qm)
        Ajc, Air, _ = _ccs_from_mask(Am)
        Gjc, Gir, _ = _ccs_from_mask(Gm)
        _PATTERN_CACHE[kind] = (Pattern(kind, n, m, p, Pjc, Pir, Ajc, Air, Gjc, Gir), Qm, Am, Gm)
    return _PATTERN_CACHE[kind][0]


def generate(kind: str, count: int, start: int = 0, seed: int = DEFAULT_SEED) -> Batch:
    """Instances ``start .. start+count`` of ``kind``; identical no matter how the range is split."""
    pat = pattern(kind)
    _, Qm, Am, Gm = _PATTERN_CACHE[kind]
    sel = [np.nonzero(M.T)[::-1] for M in (Qm, Am, Gm)]  # (rows, cols) in CCS order
    out = [np.empty((count, k)) for k in (pat.nnzP, pat.nnzA, pat.nnzG, pat.n, pat.m, pat.p)]
    done = 0
    while done < count:
        i = start + done
        chunk, off = divmod(i, CHUNK)
        take = min(CHUNK - off, count - done)
        rng = np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFFFFFFFFFF, chunk]))
        _, (Q, _), (A, _), (G, _), c, h, b = _build(kind, rng, CHUNK)
        lo, hi = off, off + take
        for dst, M, (rr, cc) in zip(out[:3], (Q, A, G), sel):
            dst[done:done + take] = M[lo:hi][:, rr, cc]
        out[3][done:done + take] = c[lo:hi]
        out[4][done:done + take] = h[lo:hi]
        out[5][done:done + take] = b[lo:hi]
        done += take
    return Batch(pat, *out)


def _fill_shared(kind, start, lo, hi, seed, names, count):
    """Worker of generate_parallel: instances [lo, hi) of the range into the parent's shared buffers."""
    from multiprocessing import resource_tracker, shared_memory
    pat = pattern(kind)
    widths = (pat.nnzP, pat.nnzA, pat.nnzG, pat.n, pat.m, pat.p)
    step = 8 * CHUNK
    views, keep = [], []
    for name, w in zip(names, widths):
        shm = shared_memory.SharedMemory(name=name)
        resource_tracker.unregister(shm._name, "shared_memory")  # the parent owns the block
        keep.append(shm)
        views.append(np.ndarray((count, max(w, 1)), dtype=np.float64, buffer=shm.buf) if w else None)
    at = lo
    while at < hi:  # (piece by piece: bounded memory; cuts on the generator's chunk boundaries)
        nxt = min(hi, ((start + at) // CHUNK + step // CHUNK) * CHUNK - start)
        part = generate(kind, nxt - at, start + at, seed)
        for v, a in zip(views, (part.Ppr, part.Apr, part.Gpr, part.c, part.h, part.b)):
            if v is not None:
                v[at:nxt] = a
        at = nxt
    del views
    for shm in keep:
        shm.close()


def generate_parallel(kind: str, count: int, start: int = 0, seed: int = DEFAULT_SEED, procs: int | None = None) -> Batch:
    """generate() on several host cores (large benchmark batches): the same instances, bit for bit -- instance i
    depends on (seed, i) only -- written by worker processes (``python -m apf_quadruped_b200.dogbot fill ...``)
    straight into shared buffers."""
    import os
    import subprocess
    import sys
    procs = procs or min(os.cpu_count() or 1, 32)
    if count < 4 * CHUNK or procs <= 1:
        return generate(kind, count, start, seed)
    from multiprocessing import shared_memory
    pat = pattern(kind)
    widths = (pat.nnzP, pat.nnzA, pat.nnzG, pat.n, pat.m, pat.p)
    blocks = [shared_memory.SharedMemory(create=True, size=max(8, 8 * count * max(w, 1))) for w in widths]
    try:
        names = [b.name for b in blocks]
        # one contiguous range per worker, cut on the generator's chunk boundaries so that no chunk is drawn twice
        first = start // CHUNK
        chunks = (start + count + CHUNK - 1) // CHUNK - first
        procs = max(1, min(procs, chunks))
        cuts = [0] + [min(count, max(0, (first + (chunks * (w + 1)) // procs) * CHUNK - start)) for w in range(procs)]
        cuts[-1] = count
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        workers = [subprocess.Popen([sys.executable, "-m", "apf_quadruped_b200.dogbot", "fill", kind, str(start), str(lo), str(hi), str(seed), str(count)] + names, cwd=root)
                   for lo, hi in zip(cuts, cuts[1:]) if hi > lo]
        for w in workers:
            if w.wait() != 0:
                raise RuntimeError("dogbot.generate_parallel: a worker failed")
        out = [np.ndarray((count, w), dtype=np.float64, buffer=b.buf).copy() if w else np.empty((count, 0)) for b, w in zip(blocks, widths)]
    finally:
        for b in blocks:
            b.close()
            b.unlink()
    return Batch(pat, *out)


def mixed(count: int, start: int = 0, seed: int = DEFAULT_SEED):
    """BASELINE.json configs[2]: thirds of stance / crawl / trot, one plan per pattern."""
    third = count // 3
    sizes = {"stance": count - 2 * third, "crawl": third, "trot": third}
    return {k: generate(k, v, start, seed + i) for i, (k, v) in enumerate(sizes.items())}


if __name__ == "__main__":
    import sys
    if len(sys.argv) >= 8 and sys.argv[1] == "fill":
        _fill_shared(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), sys.argv[8:], int(sys.argv[7]))
