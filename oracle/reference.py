"""ctypes access to the CPU oracles.  TEST INFRASTRUCTURE ONLY.

Nothing under ``apf_quadruped_b200/`` imports this module; only ``tests/``, ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs and ``__graft_entry__.smoke()`` do, and only as the
checker.  Two oracles share one interface:

* ``kind="reference"`` -- ``oracle/_ref/libqpswift_ref.so``: the UNMODIFIED reference qpSWIFT
  compiled by ``oracle/Makefile`` from ``/root/reference/dogbot_controller/src/qpSWIFT/*.c``.
* ``kind="port"`` -- ``oracle/liboracle_port.so``: our plain-C restatement
  (``oracle/qpswift_port.c``), pinned against the reference by ``tests/test_oracle.py``.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libqpswift_ref.so")
REF_O3_SO = os.path.join(HERE, "_ref", "libqpswift_ref_O3.so")  # timing variant only (-O3 -march=x86-64-v3)
PORT_SO = os.path.join(HERE, "liboracle_port.so")

L = C.c_long
D = C.c_double
PL = C.POINTER(C.c_long)
PD = C.POINTER(C.c_double)


class Smat(C.Structure):
    _fields_ = [("jc", PL), ("ir", PL), ("pr", PD), ("n", L), ("m", L), ("nnz", L)]


class Kkt(C.Structure):
    _fields_ = [("kktmatrix", C.POINTER(Smat)), ("b", PD), ("Parent", PL), ("Flag", PL), ("Lnz", PL), ("Li", PL),
                ("Lp", PL), ("Lti", PL), ("Ltp", PL), ("Pattern", PL), ("UPattern", PL), ("Y", PD), ("Lx", PD),
                ("D", PD), ("P", PL), ("Pinv", PL)]


class Stats(C.Structure):
    _fields_ = [("tsetup", D), ("tsolve", D), ("kkt_time", D), ("ldl_numeric", D), ("IterationCount", L),
                ("n_rx", D), ("n_ry", D), ("n_rz", D), ("n_mu", D), ("alpha_p", D), ("alpha_d", D), ("fval", D),
                ("Flag", L), ("AMD_RESULT", L), ("resolve_kkt", L)]


class Settings(C.Structure):
    _fields_ = [("maxit", L), ("reltol", D), ("abstol", D), ("sigma", D), ("verbose", L)]


class QP(C.Structure):
    _fields_ = [("n", L), ("m", L), ("p", L), ("sigma_d", D), ("mu", D), ("rho", D), ("x", PD), ("y", PD), ("z", PD),
                ("s", PD), ("rx", PD), ("ry", PD), ("rz", PD), ("delta", PD), ("delta_x", PD), ("delta_y", PD),
                ("delta_z", PD), ("delta_s", PD), ("ds", PD), ("lambda_", PD), ("temp", PD), ("P", C.POINTER(Smat)),
                ("c", PD), ("G", C.POINTER(Smat)), ("h", PD), ("A", C.POINTER(Smat)), ("b", PD),
                ("At", C.POINTER(Smat)), ("Gt", C.POINTER(Smat)), ("kkt", C.POINTER(Kkt)),
                ("options", C.POINTER(Settings)), ("stats", C.POINTER(Stats))]


def _pl(a):
    return a.ctypes.data_as(PL) if a is not None else None


def _pd(a):
    return a.ctypes.data_as(PD) if a is not None else None


def _bind(lib):
    lib.QP_SETUP.restype = C.POINTER(QP)
    lib.QP_SETUP.argtypes = [L, L, L, PL, PL, PD, PL, PL, PD, PL, PL, PD, PD, PD, PD, D, PL]
    lib.QP_SETUP_dense.restype = C.POINTER(QP)
    lib.QP_SETUP_dense.argtypes = [L, L, L, PD, PD, PD, PD, PD, PD, PL, C.c_int]
    lib.QP_SOLVE.restype = L
    lib.QP_SOLVE.argtypes = [C.POINTER(QP)]
    lib.QP_CLEANUP.argtypes = [C.POINTER(QP)]
    lib.QP_CLEANUP_dense.argtypes = [C.POINTER(QP)]
    return lib


_LIBS: dict = {}


def available(kind: str = "reference") -> bool:
    return os.path.exists(REF_SO if kind == "reference" else PORT_SO)


def lib(kind: str = "reference"):
    if kind not in _LIBS:
        path = REF_SO if kind == "reference" else PORT_SO
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (needs /root/reference for kind='reference')")
        _LIBS[kind] = _bind(C.CDLL(path, mode=C.RTLD_LOCAL))
    return _LIBS[kind]


def best_kind() -> str:
    return "reference" if available("reference") else "port"


@dataclass
class Solution:
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    s: np.ndarray
    flag: int
    iters: int
    fval: float
    stats: dict
    plan: dict
    init: dict


def _arr(ptr, count, dtype):
    if count == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True)


def _plan_of(qp) -> dict:
    k = qp.kkt.contents
    N = qp.n + qp.m + qp.p
    Lp = _arr(k.Lp, N + 1, np.int64)
    K = k.kktmatrix.contents
    return {
        "Kjc": _arr(K.jc, N + 1, np.int64), "Kir": _arr(K.ir, K.nnz, np.int64), "P": _arr(k.P, N, np.int64),
        "Pinv": _arr(k.Pinv, N, np.int64), "Parent": _arr(k.Parent, N, np.int64), "Lnz": _arr(k.Lnz, N, np.int64),
        "Lp": Lp, "Li": _arr(k.Li, int(Lp[N]), np.int64), "AMD_RESULT": int(qp.stats.contents.AMD_RESULT),
    }


def solve_sparse(pat, Ppr, Apr, Gpr, c, h, b, *, kind=None, sigma_d=0.0, permut=None, maxit=None, reltol=None,
                 abstol=None, want_factor=False, verbose=None) -> Solution:
    """``QP_SETUP`` (CCS, shared pattern) + ``QP_SOLVE`` + ``QP_CLEANUP`` for ONE instance."""
    kind = kind or best_kind()
    lb = lib(kind)
    if kind == "port" and permut is None:
        permut = golden_permutation(pat)  # the port does not restate AMD (see qpswift_port.c header)
    keep = [np.ascontiguousarray(a, dtype=np.int64) for a in (pat.Pjc, pat.Pir, pat.Ajc, pat.Air, pat.Gjc, pat.Gir)]
    vals = [np.ascontiguousarray(a, dtype=np.float64).copy() for a in (Ppr, Apr, Gpr, c, h, b)]
    perm = np.ascontiguousarray(permut, dtype=np.int64).copy() if permut is not None else None
    has_eq = pat.p > 0
    qp_p = lb.QP_SETUP(pat.n, pat.m, pat.p, _pl(keep[0]), _pl(keep[1]), _pd(vals[0]),
                       _pl(keep[2]) if has_eq else None, _pl(keep[3]) if has_eq else None,
                       _pd(vals[1]) if has_eq else None, _pl(keep[4]), _pl(keep[5]), _pd(vals[2]), _pd(vals[3]),
                       _pd(vals[4]), _pd(vals[5]) if has_eq else None, float(sigma_d), _pl(perm))
    qp = qp_p.contents
    init = {"x": _arr(qp.x, qp.n, float), "y": _arr(qp.y, qp.p, float) if has_eq else np.zeros(0),
            "z": _arr(qp.z, qp.m, float), "s": _arr(qp.s, qp.m, float)}
    N = qp.n + qp.m + qp.p
    if want_factor:
        lnz = int(_arr(qp.kkt.contents.Lp, N + 1, np.int64)[N])
        init["Lx"] = _arr(qp.kkt.contents.Lx, lnz, float)
        init["D"] = _arr(qp.kkt.contents.D, N, float)
    o = qp.options.contents
    if maxit is not None:
        o.maxit = int(maxit)
    if reltol is not None:
        o.reltol = float(reltol)
    if abstol is not None:
        o.abstol = float(abstol)
    if verbose is not None:
        o.verbose = int(verbose)
    flag = int(lb.QP_SOLVE(qp_p))
    st = qp.stats.contents
    stats = {f: getattr(st, f) for f, _ in Stats._fields_}
    stats["sigma"] = o.sigma
    stats["mu"] = qp.mu
    sol = Solution(_arr(qp.x, qp.n, float), _arr(qp.y, qp.p, float) if has_eq else np.zeros(0), _arr(qp.z, qp.m, float),
                   _arr(qp.s, qp.m, float), flag, int(st.IterationCount), float(st.fval), stats, _plan_of(qp), init)
    if want_factor:
        lnz = int(sol.plan["Lp"][N])
        sol.stats["Lx"] = _arr(qp.kkt.contents.Lx, lnz, float)
        sol.stats["D"] = _arr(qp.kkt.contents.D, N, float)
    lb.QP_CLEANUP(qp_p)
    return sol


def solve_dense(n, m, p, P, A, G, c, h, b, *, ordering=30, reltol=None, abstol=None, maxit=None) -> Solution:
    """``QP_SETUP_dense`` (dense blocks in the given ordering: 30 column-major, 20 row-major) + ``QP_SOLVE`` +
    ``QP_CLEANUP_dense`` of the compiled reference for ONE instance (dogbot_controller/src/qpSWIFT/qpSWIFT.c:260-456,
    densetosparse Auxilary.c:1154-1273).  ``plan`` additionally carries the CCS the reference made of the dense
    input: ``ccs = {"P": (jc, ir, pr), "A": ..., "G": ...}``."""
    lb = lib("reference")
    vals = [np.ascontiguousarray(a, dtype=np.float64).ravel().copy() for a in (P, A, G, c, h, b)]
    has_eq = p > 0
    qp_p = lb.QP_SETUP_dense(n, m, p, _pd(vals[0]), _pd(vals[1]) if has_eq else None, _pd(vals[2]), _pd(vals[3]),
                             _pd(vals[4]), _pd(vals[5]) if has_eq else None, None, int(ordering))
    qp = qp_p.contents
    o = qp.options.contents
    if maxit is not None:
        o.maxit = int(maxit)
    if reltol is not None:
        o.reltol = float(reltol)
    if abstol is not None:
        o.abstol = float(abstol)

    def ccs(mp, cols):
        M = mp.contents
        jc = _arr(M.jc, cols + 1, np.int64)
        return jc, _arr(M.ir, int(jc[cols]), np.int64), _arr(M.pr, int(jc[cols]), float)

    plan = _plan_of(qp)
    plan["ccs"] = {"P": ccs(qp.P, n), "G": ccs(qp.G, n)}
    if has_eq:
        plan["ccs"]["A"] = ccs(qp.A, n)
    flag = int(lb.QP_SOLVE(qp_p))
    st = qp.stats.contents
    stats = {f: getattr(st, f) for f, _ in Stats._fields_}
    sol = Solution(_arr(qp.x, qp.n, float), _arr(qp.y, qp.p, float) if has_eq else np.zeros(0), _arr(qp.z, qp.m, float),
                   _arr(qp.s, qp.m, float), flag, int(st.IterationCount), float(st.fval), stats, plan, {})
    lb.QP_CLEANUP_dense(qp_p)
    return sol


_GOLDEN_PLANS = None


def golden_permutation(pat):
    """AMD permutation of a DogBot pattern from tests/golden/plans.npz (made by the real reference)."""
    global _GOLDEN_PLANS
    if _GOLDEN_PLANS is None:
        _GOLDEN_PLANS = np.load(os.path.join(os.path.dirname(HERE), "tests", "golden", "plans.npz"))
    key = f"{pat.kind}_P"
    if key not in _GOLDEN_PLANS.files:
        raise KeyError(f"no golden AMD permutation for pattern {pat.kind!r}: pass permut= or build oracle/_ref")
    P = _GOLDEN_PLANS[key]
    assert len(P) == pat.N
    return P


def solve_batch(batch, *, kind=None, **kw):
    """Loop of :func:`solve_sparse` over a ``dogbot.Batch``; returns stacked arrays."""
    sols = [solve_sparse(batch.pattern, batch.Ppr[i], batch.Apr[i], batch.Gpr[i], batch.c[i], batch.h[i], batch.b[i],
                         kind=kind, **kw) for i in range(batch.count)]
    out = {k: np.stack([getattr(s_, k) for s_ in sols]) for k in ("x", "y", "z", "s")}
    out["flag"] = np.array([s_.flag for s_ in sols], dtype=np.int64)
    out["iters"] = np.array([s_.iters for s_ in sols], dtype=np.int64)
    out["fval"] = np.array([s_.fval for s_ in sols])
    out["plan"] = sols[0].plan if sols else {}
    return out


def amd_order(n, Ap, Ai, kind="reference"):
    """Reference ``amd_l_order`` with default controls (only the real reference exports it)."""
    lb = lib(kind)
    Ap = np.ascontiguousarray(Ap, dtype=np.int64)
    Ai = np.ascontiguousarray(Ai, dtype=np.int64)
    perm = np.zeros(max(n, 1), dtype=np.int64)
    ctrl = (C.c_double * 5)()
    info = (C.c_double * 20)()
    lb.amd_l_defaults(ctrl)
    lb.amd_l_order.restype = L
    st = lb.amd_l_order(L(n), _pl(Ap), _pl(Ai), _pl(perm), ctrl, info)
    return int(st), perm[:n]


def fnv1a(a) -> str:
    """Word-wise FNV-1a-64 of an int64 array (the hash SURVEY.md section 8c quotes)."""
    h = 1469598103934665603
    for v in np.asarray(a, dtype=np.int64).tolist():
        h ^= v & 0xFFFFFFFFFFFFFFFF
        h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


DRIVER_SO = os.path.join(HERE, "liboracle_driver.so")
_DRIVER = None


def run_batch(batch, *, kind=None, threads=None, permut="shared", sigma_d=0.0, maxit=None, reltol=None, abstol=None,
              outputs=True):
    """Whole batch through ``QP_SETUP`` + ``QP_SOLVE`` + ``QP_CLEANUP``, one instance per host thread
    (oracle/ref_driver.c).  ``permut="shared"``: the pattern's AMD permutation is computed once and
    supplied to every QP_SETUP (plan-reuse baseline); ``None``: every
    instance runs AMD itself (only the real reference can); an array: that permutation.
    Returns (result dict, wall seconds)."""
    global _DRIVER
    kind = kind or best_kind()
    if _DRIVER is None:
        if not os.path.exists(DRIVER_SO):
            raise FileNotFoundError(f"{DRIVER_SO} missing: run `make -C oracle driver`")
        _DRIVER = C.CDLL(DRIVER_SO, mode=C.RTLD_LOCAL)
        _DRIVER.oracle_run_batch.restype = D
    pat = batch.pattern
    cnt = batch.count
    if isinstance(permut, str):
        if kind in ("reference", "reference_O3"):
            i0 = 0
            permut = solve_sparse(pat, batch.Ppr[i0], batch.Apr[i0], batch.Gpr[i0], batch.c[i0], batch.h[i0], batch.b[i0],
                                  kind="reference", maxit=0).plan["P"]
        else:
            permut = golden_permutation(pat)
    if permut is None and kind == "port":
        permut = golden_permutation(pat)
    ints = [np.ascontiguousarray(a, dtype=np.int64) for a in (pat.Pjc, pat.Pir, pat.Ajc, pat.Air, pat.Gjc, pat.Gir)]
    perm = np.ascontiguousarray(permut, dtype=np.int64) if permut is not None else None
    vals = [np.ascontiguousarray(a, dtype=np.float64) for a in (batch.Ppr, batch.Apr, batch.Gpr, batch.c, batch.h, batch.b)]
    out = {"x": np.empty((cnt, pat.n)), "y": np.empty((cnt, pat.p)), "z": np.empty((cnt, pat.m)), "s": np.empty((cnt, pat.m)),
           "flag": np.empty(cnt, dtype=np.int64), "iters": np.empty(cnt, dtype=np.int64), "fval": np.empty(cnt)}
    vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None and a.size else None
    path = {"reference": REF_SO, "reference_O3": REF_O3_SO}.get(kind, PORT_SO).encode()
    threads = threads or (os.cpu_count() or 1)
    secs = _DRIVER.oracle_run_batch(
        C.c_char_p(path), C.c_int(int(threads)), C.c_long(cnt), L(pat.n), L(pat.m), L(pat.p),
        *[vp(a) for a in ints], vp(perm), *[vp(a) for a in vals], D(float(sigma_d)),
        L(-1 if maxit is None else int(maxit)), D(-1.0 if reltol is None else float(reltol)),
        D(-1.0 if abstol is None else float(abstol)),
        *([vp(out[k]) for k in ("x", "y", "z", "s", "flag", "iters", "fval")] if outputs else [None] * 7))
    if secs < 0:
        raise RuntimeError("oracle driver could not load " + path.decode())
    out["threads"] = int(min(threads, max(cnt, 1)))
    return out, float(secs)


def _driver():
    global _DRIVER
    if _DRIVER is None:
        if not os.path.exists(DRIVER_SO):
            raise FileNotFoundError(f"{DRIVER_SO} missing: run `make -C oracle driver`")
        _DRIVER = C.CDLL(DRIVER_SO, mode=C.RTLD_LOCAL)
        _DRIVER.oracle_run_batch.restype = D
    _DRIVER.oracle_run_batch_dense.restype = D
    _DRIVER.oracle_control_loop.restype = C.c_int
    return _DRIVER


def run_batch_dense(batch, *, lib_path=None, threads=None, reltol=None, abstol=None, outputs=True):
    """Controller style (main.cpp:1649-1663): every instance through ``QP_SETUP_dense``
    (dense column-major blocks: densetosparse + AMD + symbolic analysis per QP) + ``QP_SOLVE`` + ``QP_CLEANUP_dense``,
    one instance per pinned host thread.  ``lib_path``: any library exporting qpSWIFT's C API (default: the compiled
    reference).  Returns (dict with x, flag, iters, threads; wall seconds)."""
    drv = _driver()
    pat, cnt = batch.pattern, batch.count
    P, A, G = [np.ascontiguousarray(a) for a in batch.dense_batch()]
    vals = [P, A, G] + [np.ascontiguousarray(a, dtype=np.float64) for a in (batch.c, batch.h, batch.b)]
    out = {"x": np.empty((cnt, pat.n)), "flag": np.empty(cnt, dtype=np.int64), "iters": np.empty(cnt, dtype=np.int64)}
    vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None and a.size else None
    threads = threads or (os.cpu_count() or 1)
    secs = drv.oracle_run_batch_dense(C.c_char_p((lib_path or REF_SO).encode()), C.c_int(int(threads)), C.c_long(cnt), L(pat.n), L(pat.m), L(pat.p),
                                      *[vp(a) for a in vals], D(-1.0 if reltol is None else float(reltol)), D(-1.0 if abstol is None else float(abstol)),
                                      *([vp(out[k]) for k in ("x", "flag", "iters")] if outputs else [None] * 3))
    if secs < 0:
        raise RuntimeError(f"oracle driver could not load {lib_path or REF_SO}")
    out["threads"] = int(min(threads, max(cnt, 1)))
    return out, float(secs)


def control_loop(batch, ticks, *, lib_path=None, reltol=1e-2, abstol=1e-2, cleanup=True):
    """``ticks`` control ticks on the calling thread, one QP each: ``QP_SETUP_dense`` -> tolerances (the controller's
    1e-2, main.cpp:1651-1652) -> ``QP_SOLVE`` -> x -> ``QP_CLEANUP_dense``, cycling through ``batch``.  Works on the
    compiled reference and on libqpswift_b200.so alike (same C API).  Returns per-tick microseconds etc."""
    drv = _driver()
    pat, cnt = batch.pattern, batch.count
    P, A, G = [np.ascontiguousarray(a) for a in batch.dense_batch()]
    vals = [P, A, G] + [np.ascontiguousarray(a, dtype=np.float64) for a in (batch.c, batch.h, batch.b)]
    us, setup_us = np.empty(ticks), np.empty(ticks)
    x, flag, iters = np.empty((ticks, pat.n)), np.empty(ticks, dtype=np.int64), np.empty(ticks, dtype=np.int64)
    vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None and a.size else None
    rc = drv.oracle_control_loop(C.c_char_p((lib_path or REF_SO).encode()), C.c_long(int(ticks)), C.c_long(cnt), L(pat.n), L(pat.m), L(pat.p),
                                 *[vp(a) for a in vals], D(float(reltol)), D(float(abstol)), C.c_int(1 if cleanup else 0),
                                 vp(us), vp(setup_us), vp(x), vp(flag), vp(iters))
    if rc != 0:
        raise RuntimeError(f"oracle driver could not load {lib_path or REF_SO}")
    return {"us": us, "setup_us": setup_us, "x": x, "flag": flag, "iters": iters}
