"""Monte-Carlo rollout of the DogBot controller loop (BASELINE.json configs[3]): robots x sequential
control ticks, every tick = state update -> WBC formulation -> QP solve -> joint torques, per contact pattern, with
everything resident on the GPU and a whole tick replayed from a CUDA graph.

The reference runs ONE robot: ``DOGCTRL::ctrl_loop`` at 400 Hz (``dogbot_controller/src/client/main.cpp:1107``) against
Gazebo, iDynTree and towr, with the artificial-potential-field step planner (``main.cpp:1261-1421``, ``2745-2845``)
choosing the footholds.  None of those exist here; the loop is closed with the fixed model documented in
``csrc/rollout.cuh`` (semi-implicit Euler on the joints, a PD-driven CoM error, a fixed stance / trot / stance / crawl
schedule staggered over the robots, APF-perturbed footholds for the swing feet from a counter-based generator).

* ``Rollout`` -- the device loop: ``QPB_ROLLOUT_TICK_DEVICE`` + per pattern ``QPB_WBC_FORMULATE_DEVICE`` /
  ``QP_SOLVE_BATCH_DEVICE_INDEXED`` / ``QPB_WBC_TORQUES_DEVICE`` on the pattern's device-side robot list; one tick is
  captured once in a CUDA graph (three branches, one per contact pattern) and replayed.
* ``tick_reference`` -- the SAME model in numpy, operation for operation.  ``oracle/rollout_reference.py`` (test
  infrastructure) steps it with the CPU oracle solving every QP (cold start, as the reference always does:
  ``Auxilary.c:992-1089``); the device loop must match that bit for bit, tick by tick (``tests/test_rollout_gpu.py``).
* ``bench`` -- the ``rollout`` record of ``bench.py``.

Warm start has no reference semantics (no API for it): the rollout is cold-start, which is what keeps it
comparable with the oracle.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import dogbot, wbc

DT = 0.0025
KP_COM, KD_COM = 3000.0, 50.0
G = 9.81
MASS_G = 21.27 * 9.81
GOAL_SPREAD, SAT, STEP, HALF, KPA, KP_SW = 0.03, 0.05, 0.025, 0.5, 1.5, 300.0
PHASE_TICKS, ROBOT_STRIDE = 25, 7
QDD_SAT, KQ, KDQ, A_SAT, REPLAN = 40.0, 400.0, 40.0, 30.0, 0.25
ROW = 104
RR_R, RR_GOAL, RR_DX, RR_DV, RR_NOISE = 0, 12, 20, 26, 32
KIND_NAMES = ("stance", "trot", "crawl")
NOMINAL = np.array([[-dogbot.X_NOM, -dogbot.Y_NOM], [-dogbot.X_NOM, dogbot.Y_NOM], [dogbot.X_NOM, dogbot.Y_NOM], [dogbot.X_NOM, -dogbot.Y_NOM]])
DEFAULT_SEED = 0xA9F0D06
M64 = (1 << 64) - 1


def initial(robots: int, first: int = 0, seed: int = DEFAULT_SEED):
    """State rows [robots, 468] and rollout rows [robots, 104] of robots first .. first+robots-1 (reproducible per robot)."""
    S = np.empty((robots, wbc.STATE))
    Rw = np.empty((robots, ROW))
    for k in range(robots):
        rng = np.random.Generator(np.random.Philox(key=[seed & M64, first + k]))
        u = lambda lo, hi, *shape: rng.uniform(lo, hi, size=shape)
        signs = np.array([[-1, -1], [-1, 1], [1, 1], [1, -1]], dtype=float)
        r = np.empty((4, 3))
        r[:, 0], r[:, 1], r[:, 2] = signs[:, 0] * dogbot.X_NOM, signs[:, 1] * dogbot.Y_NOM, -0.40
        r += u(-0.03, 0.03, 4, 3)
        noise = u(-1e-3, 1e-3, 12, 6)
        Jj = u(-0.01, 0.01, 12, 12)
        for f in range(4):
            Jj[3 * f:3 * f + 3, 3 * f:3 * f + 3] = u(-0.35, 0.35, 3, 3)
        Bm = u(-0.05, 0.05, 18, 18)
        M = Bm @ Bm.T + np.diag(np.concatenate([[dogbot.ROBOT_MASS] * 3, [0.3, 0.8, 0.9], [0.08] * 12]))
        bias, Jdqd = u(-0.5, 0.5, 18), u(-0.05, 0.05, 12)
        q = 0.5 * (dogbot.QMIN + dogbot.QMAX) + u(-0.3, 0.3, 12)
        dq = u(-0.5, 0.5, 12)
        S[k] = wbc.pack_states(np.zeros((1, 12, 6)), Jj[None], M[None], bias[None], Jdqd[None], np.zeros((1, 6)), q[None], dq[None], np.zeros((1, 12)))[0]
        Rw[k, RR_R:RR_GOAL] = r.ravel()
        Rw[k, RR_GOAL:RR_DX] = (r[:, 0:2]).ravel()
        Rw[k, RR_DX:RR_DV] = u(-0.01, 0.01, 6)
        Rw[k, RR_DV:RR_NOISE] = u(-0.05, 0.05, 6)
        Rw[k, RR_NOISE:ROW] = noise.ravel()
    return S, Rw


def kinds_at(tick: int, robots: np.ndarray) -> np.ndarray:
    ph = ((tick + ROBOT_STRIDE * robots) // PHASE_TICKS) & 3
    return np.where(ph == 1, 1, np.where(ph == 3, 2, 0)).astype(np.int32)


def _hash(seed, robot, tick, foot, axis):
    """rollout_hash of csrc/rollout.cuh (splitmix64 finaliser), vectorised over robots, in Python integers."""
    out = np.empty(len(robot), dtype=np.uint64)
    for i, rb in enumerate(robot):
        z = (seed + 0x9E3779B97F4A7C15 * (int(rb) + 1) + 0xBF58476D1CE4E5B9 * (tick + 1) + 0x94D049BB133111EB * (4 * foot + axis + 1)) & M64
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        out[i] = z ^ (z >> 31)
    return out


def tick_reference(S, Rw, x, tick: int, first: int = 0, seed: int = DEFAULT_SEED):
    """rollout_tick_kernel in numpy (same operations in the same order).  Updates S, Rw in place; returns the kinds."""
    R = S.shape[0]
    gi = first + np.arange(R)
    if tick > 0:
        for j in range(12):
            qdd = np.minimum(np.maximum(x[:, 6 + j], -QDD_SAT), QDD_SAT)
            acc = (qdd - KQ * (S[:, wbc.QJ + j] - 0.5 * (dogbot.QMIN[j] + dogbot.QMAX[j]))) - KDQ * S[:, wbc.DQJ + j]
            dq = S[:, wbc.DQJ + j] + DT * acc
            S[:, wbc.DQJ + j] = dq
            S[:, wbc.QJ + j] = S[:, wbc.QJ + j] + DT * dq
        for k in range(6):
            a = np.minimum(np.maximum(x[:, k] - (G if k == 2 else 0.0), -A_SAT), A_SAT)
            dv = Rw[:, RR_DV + k] - DT * a
            Rw[:, RR_DV + k] = dv
            Rw[:, RR_DX + k] = Rw[:, RR_DX + k] + DT * dv
    kind = kinds_at(tick, gi)
    entering = np.ones(R, dtype=bool) if tick == 0 else kinds_at(tick - 1, gi) != kind
    if tick > 0:
        for k in range(6):
            Rw[:, RR_DX + k] = np.where(entering, REPLAN * Rw[:, RR_DX + k], Rw[:, RR_DX + k])
            Rw[:, RR_DV + k] = np.where(entering, REPLAN * Rw[:, RR_DV + k], Rw[:, RR_DV + k])
    swing = [kind != 0, np.zeros(R, dtype=bool), kind == 1, np.zeros(R, dtype=bool)]
    for f in range(4):
        for a in range(2):
            sw = swing[f]
            new = sw & entering
            if new.any():
                u = (_hash(seed, gi[new], tick, f, a) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
                Rw[new, RR_GOAL + 2 * f + a] = NOMINAL[f, a] + (u * (2.0 * GOAL_SPREAD) - GOAL_SPREAD)
            goal = Rw[:, RR_GOAL + 2 * f + a]
            e = np.minimum(np.maximum(Rw[:, RR_R + 3 * f + a] - goal, -SAT), SAT)
            r = Rw[:, RR_R + 3 * f + a] + STEP * (HALF * (-KPA * e))
            Rw[:, RR_R + 3 * f + a] = np.where(sw, r, Rw[:, RR_R + 3 * f + a])
            S[:, wbc.VSW + 3 * f + a] = np.where(sw, KP_SW * (goal - r), 0.0)
        S[:, wbc.VSW + 3 * f + 2] = 0.0
        rx, ry, rz = Rw[:, RR_R + 3 * f], Rw[:, RR_R + 3 * f + 1], Rw[:, RR_R + 3 * f + 2]
        zero, one = np.zeros(R), np.ones(R)
        base = [[one, zero, zero, zero, rz, -ry], [zero, one, zero, -rz, zero, rx], [zero, zero, one, ry, -rx, zero]]
        for a in range(3):
            for c in range(6):
                S[:, wbc.JC + 6 * (3 * f + a) + c] = base[a][c] + Rw[:, RR_NOISE + 6 * (3 * f + a) + c]
    for k in range(6):
        S[:, wbc.W + k] = (KP_COM * Rw[:, RR_DX + k] + KD_COM * Rw[:, RR_DV + k]) + (MASS_G if k == 2 else 0.0)
    return kind


class Rollout:
    """The device loop for ``robots`` robots on one GPU (robots first .. first+robots-1 of the whole rollout)."""

    def __init__(self, torch, robots: int, device: int = 0, first: int = 0, seed: int = DEFAULT_SEED, reltol=1e-2, abstol=1e-2, graph: bool = True):
        from . import qpswift
        self.torch, self.qpswift, self.lib = torch, qpswift, qpswift.lib()
        self.R, self.first, self.seed, self.device = int(robots), int(first), int(seed), int(device)
        self.dev = torch.device("cuda", device)
        S, Rw = initial(robots, first, seed)
        f64 = torch.float64
        self.S = torch.from_numpy(S).to(self.dev)
        self.Rw = torch.from_numpy(Rw).to(self.dev)
        self.x = torch.zeros((robots, 30), dtype=f64, device=self.dev)
        self.tau = torch.zeros((robots, 12), dtype=f64, device=self.dev)
        self.kind = torch.zeros(robots, dtype=torch.int32, device=self.dev)
        self.lists = torch.zeros((3, robots), dtype=torch.int32, device=self.dev)
        self.counts = torch.zeros(3, dtype=torch.int32, device=self.dev)
        self.tick = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.flag = torch.zeros(robots, dtype=torch.int32, device=self.dev)
        self.iters = torch.zeros(robots, dtype=torch.int32, device=self.dev)
        self.opts = qpswift.default_settings(reltol=reltol, abstol=abstol)  # the controller's tolerances (main.cpp:1651-1652)
        self.legs = []
        for name in KIND_NAMES:
            pat = dogbot.pattern(name)
            solver = qpswift.BatchSolver(pat, robots, device=device)
            rows = {k: torch.zeros((robots, w), dtype=f64, device=self.dev) for k, w in
                    (("Ppr", pat.nnzP), ("Apr", pat.nnzA), ("Gpr", pat.nnzG), ("c", pat.n), ("h", pat.m), ("b", pat.p))}
            duals = {"y": torch.zeros((robots, pat.p), dtype=f64, device=self.dev), "z": torch.zeros((robots, pat.m), dtype=f64, device=self.dev),
                     "s": torch.zeros((robots, pat.m), dtype=f64, device=self.dev)}
            self.legs.append({"name": name, "solver": solver, "form": wbc.Formulator(solver, name), "rows": rows, "duals": duals,
                              "stream": torch.cuda.Stream(device=self.dev)})
        # (the formula tables are built on first use -- not something to do inside a graph capture: an empty list now)
        names = ("Ppr", "Apr", "Gpr", "c", "h", "b")
        for k, leg in enumerate(self.legs):
            leg["form"].formulate(self.R, self.S.data_ptr(), *[leg["rows"][n].data_ptr() for n in names], dindex=self.lists[k].data_ptr(), dcount=self.counts[k:k + 1].data_ptr(),
                                  stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        self.use_graph = graph
        self.graph = None
        self.launches_per_tick = 2 + 3 * 4  # tick + lists, and per pattern formulate + counter reset + solve + torques

    def _enqueue_tick(self):
        torch = self.torch
        cur = torch.cuda.current_stream()
        rc = self.lib.QPB_ROLLOUT_TICK_DEVICE(self.device, self.R, self.first, C.c_uint64(self.seed & M64), self.S.data_ptr(), self.Rw.data_ptr(), self.x.data_ptr(),
                                              self.kind.data_ptr(), self.lists.data_ptr(), self.counts.data_ptr(), self.tick.data_ptr(), cur.cuda_stream)
        if rc < 0:
            raise RuntimeError(f"QPB_ROLLOUT_TICK_DEVICE failed ({rc}): {self.qpswift.last_error()}")
        names = ("Ppr", "Apr", "Gpr", "c", "h", "b")
        for k, leg in enumerate(self.legs):  # the three contact patterns are independent: one branch each
            st = leg["stream"]
            st.wait_stream(cur)
            sp = st.cuda_stream
            lst, cnt = self.lists[k].data_ptr(), self.counts[k:k + 1].data_ptr()
            rows, d = leg["rows"], leg["duals"]
            leg["form"].formulate(self.R, self.S.data_ptr(), *[rows[n].data_ptr() for n in names], dindex=lst, dcount=cnt, stream=sp)
            leg["solver"].solve_device_indexed(self.R, self.opts, lst, cnt, *[rows[n].data_ptr() for n in names], self.x.data_ptr(),
                                               d["y"].data_ptr() if d["y"].numel() else None, d["z"].data_ptr(), d["s"].data_ptr(), self.flag.data_ptr(), self.iters.data_ptr(), sp)
            leg["form"].torques(self.R, self.S.data_ptr(), self.x.data_ptr(), self.tau.data_ptr(), dindex=lst, dcount=cnt, stream=sp)
        for leg in self.legs:
            cur.wait_stream(leg["stream"])

    def step(self, ticks: int = 1):
        """Run `ticks` control ticks (asynchronously on the current stream)."""
        torch = self.torch
        if not self.use_graph:
            for _ in range(ticks):
                self._enqueue_tick()
            return
        if self.graph is None:
            # one untimed eager tick first would advance the state; capture only records, so capture directly
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream())
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(self.graph, stream=side, capture_error_mode="relaxed"):
                    self._enqueue_tick()
            torch.cuda.current_stream().wait_stream(side)
        for _ in range(ticks):
            self.graph.replay()

    def snapshot(self):
        self.torch.cuda.synchronize()
        g = lambda t: t.cpu().numpy()
        return {"S": g(self.S), "Rw": g(self.Rw), "x": g(self.x), "tau": g(self.tau), "kind": g(self.kind), "flag": g(self.flag), "iters": g(self.iters), "tick": int(self.tick.item())}

    def close(self):
        for leg in self.legs:
            leg["solver"].close()
        self.legs = []


def bench(args, env, torch, dist, opts, check=None):
    """bench.py's `rollout` record: robots-per-GPU x ticks sequential control ticks on every GPU (weak scaling over the
    robots; nothing crosses GPUs during the ticks), device time by CUDA events, max over ranks; robot-ticks per second
    = QP solves per second.  `check` (bench.py passes the CPU-oracle comparison of a 64-robot x 50-tick rollout) runs
    on rank 0 inside the run."""
    rank, world, local, dev = env
    R, T = int(args.robots_per_gpu), int(args.ticks)
    first = rank * R
    checked = check(torch, local) if (check is not None and rank == 0) else None
    ro = Rollout(torch, R, device=local, first=first)
    ro.step(max(3, min(25, T // 10)))  # warm-up ticks (graph capture included)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ro.step(T)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    snap = ro.snapshot()
    rec = {"workload": f"configs[3]: APF-perturbed Monte-Carlo rollout, {R} robots per GPU x {T} sequential control ticks (cold-start QPs, controller tolerances 1e-2), "
                       "stance / trot / stance / crawl schedule staggered over the robots, formulation + solve + torques + integration on the device, one CUDA graph per tick",
           "scaling": "weak", "n_gpus": world, "robots_per_gpu": R, "ticks": T, "value": world * R * T / (ms * 1e-3), "unit": "robot-ticks/s (= QP solves/s)",
           "ticks_per_s": T / (ms * 1e-3), "ms_per_tick": ms / T, "gpu_launches": T * ro.launches_per_tick, "cuda_graph": bool(ro.use_graph),
           "not_optimal_last_tick": int((snap["flag"] != 0).sum()), "iterations_mean_last_tick": float(snap["iters"].mean()),
           "warm_start": "none: the reference has no warm-start semantics (Auxilary.c:992-1089 always cold-starts); cold start keeps the loop comparable with the oracle",
           "oracle_check": checked}
    ro.close()
    return rec
