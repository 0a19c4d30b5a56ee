"""Python host mirror of the qpSWIFT C API, bound to ``libqpswift_b200.so`` through ctypes.

The product is the C-ABI shared library (``include/qpSWIFT/qpSWIFT.h``, ``include/qpswift_b200.h``);
this module only forwards to it so that tests and ``bench.py`` read like calls into the
reference (same names, same argument meaning: ``QP_SETUP`` / ``QP_SOLVE`` / ``QP_CLEANUP`` and
the batched ``QP_SOLVE_BATCH``).  There is no Python or CPU implementation of the solve here: if
the library is missing, or no CUDA device is usable, calls raise / report ``QP_FATAL``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QPB_LIB") or os.path.join(HERE, "lib", "libqpswift_b200.so")  # QPB_LIB: experiments with library variants

QP_OPTIMAL, QP_KKTFAIL, QP_MAXIT, QP_FATAL = 0, 1, 2, 3
ROW_MAJOR_ORDERING, COLUMN_MAJOR_ORDERING = 20, 30
QPB_E_NODEVICE, QPB_E_CUDA, QPB_E_ARG, QPB_E_PATTERN = -101, -102, -103, -104

L = C.c_long
D = C.c_double
PL = C.POINTER(C.c_long)
PD = C.POINTER(C.c_double)


class smat(C.Structure):
    _fields_ = [("jc", PL), ("ir", PL), ("pr", PD), ("n", L), ("m", L), ("nnz", L)]


class kkt(C.Structure):
    _fields_ = [("kktmatrix", C.POINTER(smat)), ("b", PD), ("Parent", PL), ("Flag", PL), ("Lnz", PL), ("Li", PL),
                ("Lp", PL), ("Lti", PL), ("Ltp", PL), ("Pattern", PL), ("UPattern", PL), ("Y", PD), ("Lx", PD),
                ("D", PD), ("P", PL), ("Pinv", PL)]


class stats(C.Structure):
    _fields_ = [("tsetup", D), ("tsolve", D), ("kkt_time", D), ("ldl_numeric", D), ("IterationCount", L),
                ("n_rx", D), ("n_ry", D), ("n_rz", D), ("n_mu", D), ("alpha_p", D), ("alpha_d", D), ("fval", D),
                ("Flag", L), ("AMD_RESULT", L), ("resolve_kkt", L)]


class settings(C.Structure):
    _fields_ = [("maxit", L), ("reltol", D), ("abstol", D), ("sigma", D), ("verbose", L)]


class QP(C.Structure):
    _fields_ = [("n", L), ("m", L), ("p", L), ("sigma_d", D), ("mu", D), ("rho", D), ("x", PD), ("y", PD), ("z", PD),
                ("s", PD), ("rx", PD), ("ry", PD), ("rz", PD), ("delta", PD), ("delta_x", PD), ("delta_y", PD),
                ("delta_z", PD), ("delta_s", PD), ("ds", PD), ("lambda_", PD), ("temp", PD), ("P", C.POINTER(smat)),
                ("c", PD), ("G", C.POINTER(smat)), ("h", PD), ("A", C.POINTER(smat)), ("b", PD),
                ("At", C.POINTER(smat)), ("Gt", C.POINTER(smat)), ("kkt", C.POINTER(kkt)),
                ("options", C.POINTER(settings)), ("stats", C.POINTER(stats))]


#: every symbol include/qpSWIFT/qpSWIFT.h and include/qpswift_b200.h declare
EXPORTS = [
    "QP_SETUP", "QP_SETUP_dense", "QP_SOLVE", "QP_CLEANUP", "QP_CLEANUP_dense",
    "QPB_PLAN_CREATE", "QPB_PLAN_DESTROY", "QPB_PLAN_ARRAY", "QPB_PLAN_INFO", "QPB_AMD_ORDER", "QPB_PLAN_PROGRAM",
    "QPB_PLAN_LAYOUT", "QPB_PLAN_STREAM", "QPB_PLAN_STREAM_INFO", "QPB_PLAN_STREAM_KEY", "QPB_GENERATED_KERNELS",
    "QP_SETUP_BATCH", "QPB_BATCH_CREATE", "QP_SOLVE_BATCH", "QP_SOLVE_BATCH_DEVICE", "QP_CLEANUP_BATCH",
    "QP_SOLVE_BATCH_DENSE", "QPB_DENSE_TO_VALUES_DEVICE",
    "QPB_BATCH_READ_STATS", "QPB_BATCH_SET_SIGMA_D", "QPB_BATCH_ENABLE_TRACE", "QPB_BATCH_READ_TRACE", "QPB_BATCH_LAST_KERNEL_MS", "QPB_BATCH_LAST_LAUNCHES", "QPB_BATCH_LAST_KERNEL",
    "QPB_LDL_NUMERIC_BATCH", "QPB_LDL_SOLVE_BATCH", "QPB_LAST_ERROR", "QPB_DEVICE_COUNT", "QPB_PLAN_CACHE_INFO",
    "QPB_BATCH_ROW_WIDTHS", "QP_SETUP_BATCH_MULTI", "QPB_BATCH_CREATE_MULTI", "QPB_MULTI_DEVICES", "QPB_MULTI_SHARD",
    "QP_SOLVE_BATCH_MULTI", "QP_SOLVE_BATCH_MULTI_DEVICE", "QP_CLEANUP_BATCH_MULTI", "QPB_MULTI_LAST_ERROR", "QPB_FP64_PEAK",
    "QP_SOLVE_BATCH_DEVICE_INDEXED", "QPB_WBC_FORMULATE_DEVICE", "QPB_WBC_TORQUES_DEVICE", "QPB_WBC_DESCRIPTORS", "QPB_ROLLOUT_TICK_DEVICE",
]

_lib = None


def lib():
    """The loaded C-ABI library; raises if it has not been built (``make`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make` (there is no fallback implementation)")
        lb = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        V = C.c_void_p
        lb.QP_SETUP.restype = C.POINTER(QP)
        lb.QP_SETUP.argtypes = [L, L, L, PL, PL, PD, PL, PL, PD, PL, PL, PD, PD, PD, PD, D, PL]
        lb.QP_SETUP_dense.restype = C.POINTER(QP)
        lb.QP_SETUP_dense.argtypes = [L, L, L, PD, PD, PD, PD, PD, PD, PL, C.c_int]
        lb.QP_SOLVE.restype = L
        lb.QP_SOLVE.argtypes = [C.POINTER(QP)]
        lb.QP_CLEANUP.argtypes = [C.POINTER(QP)]
        lb.QP_CLEANUP_dense.argtypes = [C.POINTER(QP)]
        lb.QPB_PLAN_CREATE.restype = V
        lb.QPB_PLAN_CREATE.argtypes = [L, L, L, PL, PL, PL, PL, PL, PL, PL]
        lb.QPB_PLAN_DESTROY.argtypes = [V]
        lb.QPB_PLAN_ARRAY.restype = PL
        lb.QPB_PLAN_ARRAY.argtypes = [V, C.c_int, PL]
        lb.QPB_PLAN_INFO.restype = L
        lb.QPB_PLAN_INFO.argtypes = [V, C.c_int]
        lb.QPB_PLAN_PROGRAM.restype = C.POINTER(C.c_uint64)
        lb.QPB_PLAN_PROGRAM.argtypes = [V, PL]
        lb.QPB_PLAN_LAYOUT.restype = L
        lb.QPB_PLAN_LAYOUT.argtypes = [V, C.POINTER(C.c_int), C.c_int]
        lb.QPB_PLAN_STREAM.restype = C.POINTER(C.c_uint64)
        lb.QPB_PLAN_STREAM.argtypes = [V, PL]
        lb.QPB_PLAN_STREAM_INFO.restype = L
        lb.QPB_PLAN_STREAM_INFO.argtypes = [V, C.POINTER(C.c_longlong), C.c_int]
        lb.QPB_PLAN_STREAM_KEY.restype = C.c_uint64
        lb.QPB_PLAN_STREAM_KEY.argtypes = [V]
        lb.QPB_GENERATED_KERNELS.restype = L
        lb.QPB_GENERATED_KERNELS.argtypes = [C.POINTER(C.c_uint64), C.c_int]
        lb.QPB_AMD_ORDER.restype = L
        lb.QPB_AMD_ORDER.argtypes = [L, PL, PL, PL]
        lb.QP_SETUP_BATCH.restype = V
        lb.QP_SETUP_BATCH.argtypes = [C.POINTER(QP), L, C.c_int]
        lb.QPB_BATCH_CREATE.restype = V
        lb.QPB_BATCH_CREATE.argtypes = [V, D, L, C.c_int]
        lb.QP_SOLVE_BATCH.restype = L
        lb.QP_SOLVE_BATCH.argtypes = [V, L, C.POINTER(settings)] + [V] * 6 + [V] * 4 + [V, V]
        lb.QP_SOLVE_BATCH_DEVICE.restype = L
        lb.QP_SOLVE_BATCH_DEVICE.argtypes = [V, L, C.POINTER(settings)] + [V] * 6 + [V] * 4 + [V, V, V, V, V]
        lb.QP_SOLVE_BATCH_DENSE.restype = L
        lb.QP_SOLVE_BATCH_DENSE.argtypes = [V, L, C.POINTER(settings), C.c_int] + [V] * 6 + [V] * 4 + [V, V, V]
        lb.QPB_DENSE_TO_VALUES_DEVICE.restype = L
        lb.QPB_DENSE_TO_VALUES_DEVICE.argtypes = [V, L, C.c_int] + [V] * 8
        lb.QP_CLEANUP_BATCH.argtypes = [V]
        lb.QPB_BATCH_READ_STATS.restype = L
        lb.QPB_BATCH_READ_STATS.argtypes = [V, L, V, V]
        lb.QPB_BATCH_ENABLE_TRACE.restype = L
        lb.QPB_BATCH_ENABLE_TRACE.argtypes = [V, C.c_int]
        lb.QPB_BATCH_READ_TRACE.restype = L
        lb.QPB_BATCH_READ_TRACE.argtypes = [V, L, L, V]
        lb.QPB_BATCH_SET_SIGMA_D.restype = L
        lb.QPB_BATCH_SET_SIGMA_D.argtypes = [V, D]
        lb.QPB_BATCH_LAST_KERNEL_MS.restype = D
        lb.QPB_BATCH_LAST_KERNEL_MS.argtypes = [V]
        lb.QPB_BATCH_LAST_LAUNCHES.restype = L
        lb.QPB_BATCH_LAST_LAUNCHES.argtypes = [V]
        lb.QPB_BATCH_LAST_KERNEL.restype = C.c_int
        lb.QPB_BATCH_LAST_KERNEL.argtypes = [V]
        lb.QPB_LDL_NUMERIC_BATCH.restype = L
        lb.QPB_LDL_NUMERIC_BATCH.argtypes = [V, L, V, V, V, V, V]
        lb.QPB_LDL_SOLVE_BATCH.restype = L
        lb.QPB_LDL_SOLVE_BATCH.argtypes = [V, L, V, V, V, V, V]
        lb.QPB_PLAN_CACHE_INFO.restype = L
        lb.QPB_PLAN_CACHE_INFO.argtypes = [C.c_int]
        lb.QPB_BATCH_CREATE_MULTI.restype = V
        lb.QPB_BATCH_CREATE_MULTI.argtypes = [V, L, L, L, D, L, C.POINTER(C.c_int), C.c_int]
        lb.QP_SETUP_BATCH_MULTI.restype = V
        lb.QP_SETUP_BATCH_MULTI.argtypes = [C.POINTER(QP), L, C.POINTER(C.c_int), C.c_int]
        lb.QPB_MULTI_DEVICES.restype = C.c_int
        lb.QPB_MULTI_DEVICES.argtypes = [V]
        lb.QPB_MULTI_SHARD.restype = None
        lb.QPB_MULTI_SHARD.argtypes = [V, L, C.c_int, PL, PL]
        lb.QP_SOLVE_BATCH_MULTI.restype = L
        lb.QP_SOLVE_BATCH_MULTI.argtypes = [V, L, C.POINTER(settings)] + [V] * 6 + [V] * 4 + [V, V]
        lb.QP_SOLVE_BATCH_MULTI_DEVICE.restype = L
        lb.QP_SOLVE_BATCH_MULTI_DEVICE.argtypes = [V, L, C.POINTER(settings)] + [V] * 15
        lb.QP_CLEANUP_BATCH_MULTI.argtypes = [V]
        lb.QPB_MULTI_LAST_ERROR.restype = C.c_char_p
        lb.QPB_FP64_PEAK.restype = D
        lb.QP_SOLVE_BATCH_DEVICE_INDEXED.restype = L
        lb.QP_SOLVE_BATCH_DEVICE_INDEXED.argtypes = [V, L, C.POINTER(settings), V, V] + [V] * 6 + [V] * 4 + [V, V, V]
        lb.QPB_ROLLOUT_TICK_DEVICE.restype = L
        lb.QPB_ROLLOUT_TICK_DEVICE.argtypes = [C.c_int, L, L, C.c_uint64] + [V] * 8
        lb.QPB_WBC_DESCRIPTORS.restype = L
        lb.QPB_WBC_DESCRIPTORS.argtypes = [V, C.c_int, C.POINTER(C.c_uint64), L, PD]
        lb.QPB_FP64_PEAK.argtypes = [C.c_int, C.c_int]
        lb.QPB_LAST_ERROR.restype = C.c_char_p
        lb.QPB_DEVICE_COUNT.restype = C.c_int
        _lib = lb
    return _lib


def last_error() -> str:
    return lib().QPB_LAST_ERROR().decode()


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _pl(a):
    return a.ctypes.data_as(PL) if a is not None else None


def _pd(a):
    return a.ctypes.data_as(PD) if a is not None else None


def default_settings(maxit=100, reltol=1e-6, abstol=1e-6, sigma=100.0, verbose=0) -> settings:
    """qpSWIFT defaults (GlobalOptions.h:46-50)."""
    return settings(int(maxit), float(reltol), float(abstol), float(sigma), int(verbose))


def amd_order(n: int, Ap, Ai):
    """``amd_l_order`` with default controls (amd_order.c:21-199): returns (status, permutation)."""
    Ap, Ai = _i64(Ap), _i64(Ai)
    perm = np.zeros(max(int(n), 1), dtype=np.int64)
    st = lib().QPB_AMD_ORDER(int(n), _pl(Ap), _pl(Ai), _pl(perm))
    return int(st), perm[:n]


class Plan:
    """Symbolic plan of one sparsity pattern (host only; no GPU needed)."""

    ARRAYS = {"Kjc": 0, "Kir": 1, "P": 2, "Pinv": 3, "Parent": 4, "Lnz": 5, "Lp": 6, "Li": 7}
    INFO = {"N": 0, "nnzK": 1, "lnz": 2, "amd_status": 3, "factor_flops": 4, "nnzU": 5, "valid": 6, "plan_words": 7,
            "program_ok": 8, "slice_doubles": 9, "program_words": 10}
    SEGMENT = {"factor": 32, "solve": 48}
    SEGMENT_STATS = ["waves", "mul_waves", "div_waves", "terms", "divs", "chains", "lane_slots", "temps", "sum_waves",
                     "sums", "suspensions", "first", "width"]
    LAYOUT = ["zero", "V", "c", "h", "b", "zd", "Lx", "D", "T", "rhs", "x", "y", "z", "s", "r", "lam", "ds", "dz", "dsl",
              "total", "tsize"]

    def __init__(self, pattern, permut=None):
        self.pattern = pattern
        self._keep = [_i64(a) for a in (pattern.Pjc, pattern.Pir, pattern.Ajc, pattern.Air, pattern.Gjc, pattern.Gir)]
        perm = _i64(permut) if permut is not None else None
        has_eq = pattern.p > 0
        self.handle = lib().QPB_PLAN_CREATE(pattern.n, pattern.m, pattern.p, _pl(self._keep[0]), _pl(self._keep[1]),
                                            _pl(self._keep[2]) if has_eq else None,
                                            _pl(self._keep[3]) if has_eq else None, _pl(self._keep[4]),
                                            _pl(self._keep[5]), _pl(perm))

    def array(self, name: str) -> np.ndarray:
        cnt = C.c_long()
        ptr = lib().QPB_PLAN_ARRAY(self.handle, self.ARRAYS[name], C.byref(cnt))
        if cnt.value == 0:
            return np.zeros(0, dtype=np.int64)
        return np.ctypeslib.as_array(ptr, shape=(cnt.value,)).copy()

    def info(self, name: str) -> int:
        return int(lib().QPB_PLAN_INFO(self.handle, self.INFO[name]))

    def segment(self, which: str) -> dict:
        """Statistics of the compiled ``factor`` / ``solve`` schedule (csrc/program.h)."""
        base = self.SEGMENT[which]
        return {n: int(lib().QPB_PLAN_INFO(self.handle, base + i)) for i, n in enumerate(self.SEGMENT_STATS)}

    def layout(self) -> dict:
        buf = (C.c_int * 32)()
        cnt = int(lib().QPB_PLAN_LAYOUT(self.handle, buf, 32))
        assert cnt == len(self.LAYOUT)
        return {n: int(buf[i]) for i, n in enumerate(self.LAYOUT)}

    def program(self, which: str | None = None) -> np.ndarray:
        """The instruction words (uint64): all of them, flat, or one schedule as an array of shape
        (super-waves, 32 lanes, width)."""
        cnt = C.c_long()
        ptr = lib().QPB_PLAN_PROGRAM(self.handle, C.byref(cnt))
        words = np.ctypeslib.as_array(ptr, shape=(cnt.value,)).copy() if cnt.value else np.zeros(0, dtype=np.uint64)
        if which is None:
            return words
        seg = self.segment(which)
        lo = seg["first"] * 32
        return words[lo:lo + seg["waves"] * 32 * seg["width"]].reshape(seg["waves"], 32, seg["width"])

    STREAM_LAYOUT = ["zero", "one", "mone", "V", "c", "h", "b", "zd", "Lx", "D", "rhs", "x", "y", "z", "s", "r", "lam", "ds", "dz",
                     "dsl", "px", "hs", "nb", "total", "scratch"]
    STREAM_SEGMENTS = ["factor", "solve", "resid", "resz"]
    STREAM_STATS = ["first", "groups", "ops", "terms", "divs", "late", "hbm_loads", "hbm_stores", "nops", "slow_groups"]

    def stream_info(self) -> dict:
        """Layout, group size and per-segment statistics of the throughput kernel's micro-programs."""
        buf = (C.c_longlong * 128)()
        cnt = int(lib().QPB_PLAN_STREAM_INFO(self.handle, buf, 128))
        nl = len(self.STREAM_LAYOUT)
        assert cnt == nl + 2 + 10 * len(self.STREAM_SEGMENTS)
        info = {"layout": {n: int(buf[i]) for i, n in enumerate(self.STREAM_LAYOUT)}, "ok": int(buf[nl]),
                "group": int(buf[nl + 1]), "segments": {}}
        for s, name in enumerate(self.STREAM_SEGMENTS):
            base = nl + 2 + 10 * s
            info["segments"][name] = {k: int(buf[base + i]) for i, k in enumerate(self.STREAM_STATS)}
        return info

    def stream_key(self) -> int:
        """Key of the micro-programs: a kernel generated from them (codegen.py) registers under it."""
        return int(lib().QPB_PLAN_STREAM_KEY(self.handle))

    def stream_words(self, which: str | None = None) -> np.ndarray:
        cnt = C.c_long()
        ptr = lib().QPB_PLAN_STREAM(self.handle, C.byref(cnt))
        words = np.ctypeslib.as_array(ptr, shape=(cnt.value,)).copy() if cnt.value else np.zeros(0, dtype=np.uint64)
        if which is None:
            return words
        info = self.stream_info()
        seg = info["segments"][which]
        return words[seg["first"]:seg["first"] + seg["groups"] * info["group"]]

    def close(self):
        if self.handle:
            lib().QPB_PLAN_DESTROY(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchSolver:
    """``QP_SETUP_BATCH`` / ``QP_SOLVE_BATCH`` / ``QP_CLEANUP_BATCH`` for one pattern on one device."""

    def __init__(self, pattern, capacity: int, device: int = -1, sigma_d: float = 0.0, permut=None):
        self.pattern = pattern
        self.plan = Plan(pattern, permut)
        self.capacity = int(capacity)
        self.handle = lib().QPB_BATCH_CREATE(self.plan.handle, float(sigma_d), self.capacity, int(device))
        if not self.handle:
            raise RuntimeError(f"QPB_BATCH_CREATE failed: {last_error()}")

    # -- host arrays in, host arrays out (H2D and D2H inside the call) -----------------------------
    def solve(self, batch, opts: settings | None = None, out=None):
        pat = self.pattern
        cnt = batch.count
        opts = opts or default_settings()
        if out is None:
            out = {"x": np.empty((cnt, pat.n)), "y": np.empty((cnt, pat.p)), "z": np.empty((cnt, pat.m)),
                   "s": np.empty((cnt, pat.m)), "flag": np.empty(cnt, dtype=np.int64),
                   "iters": np.empty(cnt, dtype=np.int64)}
        arrs = [_f64(a) for a in (batch.Ppr, batch.Apr, batch.Gpr, batch.c, batch.h, batch.b)]
        ptr = lambda a: a.ctypes.data if a.size else None
        rc = lib().QP_SOLVE_BATCH(self.handle, cnt, C.byref(opts), *[ptr(a) for a in arrs], ptr(out["x"]), ptr(out["y"]),
                                  ptr(out["z"]), ptr(out["s"]), ptr(out["flag"]), ptr(out["iters"]))
        if rc < 0:
            raise RuntimeError(f"QP_SOLVE_BATCH failed ({rc}): {last_error()}")
        out["not_optimal"] = int(rc)
        return out

    def solve_dense(self, P, A, G, c, h, b, opts: settings | None = None, ordering: int = COLUMN_MAJOR_ORDERING):
        """``QP_SOLVE_BATCH_DENSE``: dense P [count, n*n], A [count, p*n], G [count, m*n] (each block in the given
        ordering, as ``QP_SETUP_dense`` takes them), host arrays in and out; also returns the per-QP pattern
        violation bits (1: nonzero outside the plan's pattern, 2: exact zero inside it)."""
        pat = self.pattern
        cnt = int(np.asarray(c).shape[0])
        opts = opts or default_settings()
        out = {"x": np.empty((cnt, pat.n)), "y": np.empty((cnt, pat.p)), "z": np.empty((cnt, pat.m)),
               "s": np.empty((cnt, pat.m)), "flag": np.empty(cnt, dtype=np.int64),
               "iters": np.empty(cnt, dtype=np.int64), "violation": np.empty(cnt, dtype=np.int64)}
        arrs = [_f64(a) for a in (P, A, G, c, h, b)]
        ptr = lambda a: a.ctypes.data if a.size else None
        rc = lib().QP_SOLVE_BATCH_DENSE(self.handle, cnt, C.byref(opts), int(ordering), *[ptr(a) for a in arrs],
                                        ptr(out["x"]), ptr(out["y"]), ptr(out["z"]), ptr(out["s"]), ptr(out["flag"]),
                                        ptr(out["iters"]), ptr(out["violation"]))
        if rc < 0:
            raise RuntimeError(f"QP_SOLVE_BATCH_DENSE failed ({rc}): {last_error()}")
        out["not_optimal"] = int(rc)
        return out

    def dense_to_values_device(self, count, ordering, dP, dA, dG, dPpr, dApr, dGpr, dviolation=None, stream=None):
        rc = lib().QPB_DENSE_TO_VALUES_DEVICE(self.handle, int(count), int(ordering), dP, dA, dG, dPpr, dApr, dGpr,
                                              dviolation, stream)
        if rc < 0:
            raise RuntimeError(f"QPB_DENSE_TO_VALUES_DEVICE failed ({rc}): {last_error()}")

    def read_stats(self, count: int):
        fval = np.empty(count)
        stat = np.empty((count, 8))
        rc = lib().QPB_BATCH_READ_STATS(self.handle, count, fval.ctypes.data, stat.ctypes.data)
        if rc < 0:
            raise RuntimeError(f"QPB_BATCH_READ_STATS failed ({rc}): {last_error()}")
        return fval, stat

    # -- raw device pointers (ints), asynchronous on `stream` -------------------------------------
    def solve_device(self, count, opts, dPpr, dApr, dGpr, dc, dh, db, dx, dy, dz, ds, dflag, diters, dfval=None,
                     dstat=None, stream=None):
        rc = lib().QP_SOLVE_BATCH_DEVICE(self.handle, int(count), C.byref(opts), dPpr, dApr, dGpr, dc, dh, db, dx, dy,
                                         dz, ds, dflag, diters, dfval, dstat, stream)
        if rc < 0:
            raise RuntimeError(f"QP_SOLVE_BATCH_DEVICE failed ({rc}): {last_error()}")

    def solve_device_indexed(self, rows, opts, dindex, dcount, dPpr, dApr, dGpr, dc, dh, db, dx, dy, dz, ds, dflag, diters, stream=None):
        """``QP_SOLVE_BATCH_DEVICE_INDEXED``: the QPs listed in dindex[0 .. *dcount) (device memory)."""
        rc = lib().QP_SOLVE_BATCH_DEVICE_INDEXED(self.handle, int(rows), C.byref(opts), dindex, dcount, dPpr, dApr, dGpr, dc, dh, db, dx, dy, dz, ds, dflag, diters, stream)
        if rc < 0:
            raise RuntimeError(f"QP_SOLVE_BATCH_DEVICE_INDEXED failed ({rc}): {last_error()}")

    def ldl_numeric_device(self, count, dvals, dzdiag, dLx, dD, stream=None):
        rc = lib().QPB_LDL_NUMERIC_BATCH(self.handle, int(count), dvals, dzdiag, dLx, dD, stream)
        if rc < 0:
            raise RuntimeError(f"QPB_LDL_NUMERIC_BATCH failed ({rc}): {last_error()}")

    def ldl_solve_device(self, count, dLx, dD, drhs, dsol, stream=None):
        rc = lib().QPB_LDL_SOLVE_BATCH(self.handle, int(count), dLx, dD, drhs, dsol, stream)
        if rc < 0:
            raise RuntimeError(f"QPB_LDL_SOLVE_BATCH failed ({rc}): {last_error()}")

    @property
    def last_kernel_ms(self) -> float:
        return float(lib().QPB_BATCH_LAST_KERNEL_MS(self.handle))

    @property
    def last_kernel(self) -> str:
        return {0: "warp", 1: "stream", 2: "stream", 3: "warp"}.get(int(lib().QPB_BATCH_LAST_KERNEL(self.handle)), "?")

    @property
    def last_kernel_generated(self) -> bool:
        """The kernel that ran was a pattern-specialised one (codegen.py / codegen_warp.py), not an interpreter."""
        return int(lib().QPB_BATCH_LAST_KERNEL(self.handle)) in (2, 3)

    @property
    def last_launches(self) -> int:
        return int(lib().QPB_BATCH_LAST_LAUNCHES(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            lib().QP_CLEANUP_BATCH(self.handle)
            self.handle = None
        self.plan.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiSolver:
    """``QP_SETUP_BATCH_MULTI`` / ``QP_SOLVE_BATCH_MULTI``: one batch cut into contiguous shards over several
    devices inside the library (one host thread and one copy of the plan per device, no collective on the
    host-pointer path)."""

    def __init__(self, pattern, capacity: int, devices, sigma_d: float = 0.0, permut=None):
        self.pattern = pattern
        self.plan = Plan(pattern, permut)
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        self.handle = lib().QPB_BATCH_CREATE_MULTI(self.plan.handle, pattern.n, pattern.m, pattern.p, float(sigma_d), int(capacity), arr, len(self.devices))
        if not self.handle:
            raise RuntimeError(f"QPB_BATCH_CREATE_MULTI failed: {lib().QPB_MULTI_LAST_ERROR().decode()}")

    def shard(self, count: int, d: int):
        lo, hi = C.c_long(), C.c_long()
        lib().QPB_MULTI_SHARD(self.handle, int(count), int(d), C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    def solve(self, batch, opts: settings | None = None, out=None):
        pat = self.pattern
        cnt = batch.count
        opts = opts or default_settings()
        if out is None:
            out = {"x": np.empty((cnt, pat.n)), "y": np.empty((cnt, pat.p)), "z": np.empty((cnt, pat.m)),
                   "s": np.empty((cnt, pat.m)), "flag": np.empty(cnt, dtype=np.int64),
                   "iters": np.empty(cnt, dtype=np.int64)}
        arrs = [_f64(a) for a in (batch.Ppr, batch.Apr, batch.Gpr, batch.c, batch.h, batch.b)]
        ptr = lambda a: a.ctypes.data if a.size else None
        rc = lib().QP_SOLVE_BATCH_MULTI(self.handle, cnt, C.byref(opts), *[ptr(a) for a in arrs], ptr(out["x"]), ptr(out["y"]),
                                        ptr(out["z"]), ptr(out["s"]), ptr(out["flag"]), ptr(out["iters"]))
        if rc < 0:
            raise RuntimeError(f"QP_SOLVE_BATCH_MULTI failed ({rc}): {lib().QPB_MULTI_LAST_ERROR().decode()}")
        out["not_optimal"] = int(rc)
        return out

    def close(self):
        if getattr(self, "handle", None):
            lib().QP_CLEANUP_BATCH_MULTI(self.handle)
            self.handle = None
        self.plan.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def solve_single(pat, Ppr, Apr, Gpr, c, h, b, *, sigma_d=0.0, permut=None, maxit=None, reltol=None, abstol=None, verbose=None, cleanup=True):
    """``QP_SETUP`` + ``QP_SOLVE`` + ``QP_CLEANUP`` for one instance, exactly as a C caller would."""
    lb = lib()
    keep = [_i64(a) for a in (pat.Pjc, pat.Pir, pat.Ajc, pat.Air, pat.Gjc, pat.Gir)]
    vals = [_f64(a).copy() for a in (Ppr, Apr, Gpr, c, h, b)]
    perm = _i64(permut).copy() if permut is not None else None
    eq = pat.p > 0
    qp_p = lb.QP_SETUP(pat.n, pat.m, pat.p, _pl(keep[0]), _pl(keep[1]), _pd(vals[0]), _pl(keep[2]) if eq else None,
                       _pl(keep[3]) if eq else None, _pd(vals[1]) if eq else None, _pl(keep[4]), _pl(keep[5]),
                       _pd(vals[2]), _pd(vals[3]), _pd(vals[4]), _pd(vals[5]) if eq else None, float(sigma_d), _pl(perm))
    qp = qp_p.contents
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
    grab = lambda ptr, k: np.ctypeslib.as_array(ptr, shape=(k,)).copy() if k else np.zeros(0)
    st = qp.stats.contents
    out = {"x": grab(qp.x, qp.n), "y": grab(qp.y, qp.p) if eq else np.zeros(0), "z": grab(qp.z, qp.m),
           "s": grab(qp.s, qp.m), "flag": flag, "iters": int(st.IterationCount), "fval": float(st.fval),
           "stats": {f: getattr(st, f) for f, _ in stats._fields_}, "sigma": o.sigma,
           "P": np.ctypeslib.as_array(qp.kkt.contents.P, shape=(pat.N,)).copy()}
    if cleanup:
        lb.QP_CLEANUP(qp_p)
    return out
