"""Where the end-to-end time of QP_SOLVE_BATCH (host buffers) goes: wall time per call, the kernel's own time inside it
(events around the launch) and the device-resident time of the same batch.  Development tool (GPU box)."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="trot")
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
import torch
from apf_quadruped_b200 import dogbot, qpswift
B = a.batch
batch = dogbot.generate(a.kind, B)
pat = batch.pattern
opts = qpswift.default_settings()
host = {k: torch.from_numpy(getattr(batch, k)).pin_memory() for k in ("Ppr", "Apr", "Gpr", "c", "h", "b")}
hb = dogbot.Batch(pat, *[host[k].numpy() for k in ("Ppr", "Apr", "Gpr", "c", "h", "b")])
for pinned_out in (True,):
    for env in ({}, {"QPB_NO_OVERLAP": "1"}):
        os.environ.pop("QPB_NO_OVERLAP", None)
        os.environ.update(env)
        solver = qpswift.BatchSolver(pat, B, device=0)
        mk = (lambda *s: torch.empty(s, dtype=torch.float64).pin_memory().numpy()) if pinned_out else (lambda *s: np.empty(s))
        out = {"x": mk(B, pat.n), "y": mk(B, max(pat.p, 1)), "z": mk(B, pat.m), "s": mk(B, pat.m), "flag": np.empty(B, dtype=np.int64), "iters": np.empty(B, dtype=np.int64)}
        for _ in range(2):
            solver.solve(hb, opts, out=out)
        wall, kern = [], []
        for _ in range(a.steps):
            t0 = time.perf_counter()
            solver.solve(hb, opts, out=out)
            wall.append((time.perf_counter() - t0) * 1e3)
            kern.append(solver.last_kernel_ms)
        print(f"pinned_out={pinned_out} env={env}: wall ms {np.median(wall):.2f} (min {min(wall):.2f}), kernel ms inside {np.median(kern):.2f}, e2e {B / np.median(wall) * 1e3:.0f} solves/s", flush=True)
        solver.close()
