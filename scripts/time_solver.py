"""Times QP_SOLVE_BATCH_DEVICE for kernel variants selected by environment switches, and checks the results
against each other and against an oracle sample.  Development tool (GPU box):

    python scripts/time_solver.py --kind trot --batch 65536 --variants stream warp warp_interp
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "stream": {},                                           # default dispatch (throughput kernel above 2048 QPs)
    "warp": {"QPB_STREAM_MIN": "1000000000"},               # generated warp-per-QP kernel for everything
    "warp_interp": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_INTERPRET": "1"},
    "warp_free": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_FREE": "1"},
    "warp4": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_WARPS": "4"},
    "warp5": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_WARPS": "5"},
    "warp6": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_WARPS": "6"},
    "warp4_free": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_WARPS": "4", "QPB_WARP_FREE": "1"},
    "warp2": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_WARPS": "2"},
    "warp3": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_WARPS": "3"},
    "warp7": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_WARPS": "7"},
    "warp8": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_WARPS": "8"},
    "warp9": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_WARPS": "9"},
    "warp10": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_WARPS": "10"},
    "warp11": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_WARPS": "11"},
    # the warps of a CTA in G groups that run 1/G of a cycle apart (QPB_WARP_GROUP_DELAY: ns between two groups' starts)
    "warp_g2": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_GROUPS": "2", "QPB_WARP_GROUP_DELAY": "25000"},
    "warp_g3": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_GROUPS": "3", "QPB_WARP_GROUP_DELAY": "17000"},
    "warp_g4": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_GROUPS": "4", "QPB_WARP_GROUP_DELAY": "13000"},
    "warp_g6": {"QPB_STREAM_MIN": "1000000000", "QPB_WARP_GROUPS": "6", "QPB_WARP_GROUP_DELAY": "9000"},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="trot")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--variants", nargs="*", default=["stream", "warp"])
    ap.add_argument("--oracle", type=int, default=256, help="instances checked against the CPU oracle")
    ap.add_argument("--env", nargs="*", default=[], help="extra KEY=VALUE switches for every variant")
    a = ap.parse_args()
    import torch
    from apf_quadruped_b200 import dogbot, qpswift
    dev = torch.device("cuda", 0)
    B = a.batch
    batch = dogbot.generate(a.kind, B)
    pat = batch.pattern
    opts = qpswift.default_settings()
    devin = {k: torch.from_numpy(getattr(batch, k)).to(dev) for k in ("Ppr", "Apr", "Gpr", "c", "h", "b")}
    ptr = lambda t: t.data_ptr() if t.numel() else None
    results = {}
    for name in a.variants:
        for k in ("QPB_STREAM_MIN", "QPB_WARP_INTERPRET", "QPB_WARP_FREE", "QPB_WARP_WARPS", "QPB_WARP_GROUPS", "QPB_WARP_GROUP_DELAY"):
            os.environ.pop(k, None)
        os.environ.update(VARIANTS[name])
        os.environ.update(dict(e.split("=", 1) for e in a.env))
        solver = qpswift.BatchSolver(pat, B, device=0)
        out = {"x": torch.empty((B, pat.n), dtype=torch.float64, device=dev), "y": torch.empty((B, max(pat.p, 1)), dtype=torch.float64, device=dev),
               "z": torch.empty((B, pat.m), dtype=torch.float64, device=dev), "s": torch.empty((B, pat.m), dtype=torch.float64, device=dev),
               "flag": torch.empty(B, dtype=torch.int32, device=dev), "iters": torch.empty(B, dtype=torch.int32, device=dev)}
        st = torch.cuda.current_stream().cuda_stream

        def step():
            solver.solve_device(B, opts, ptr(devin["Ppr"]), ptr(devin["Apr"]), ptr(devin["Gpr"]), ptr(devin["c"]), ptr(devin["h"]), ptr(devin["b"]),
                                ptr(out["x"]), ptr(out["y"]) if pat.p else None, ptr(out["z"]), ptr(out["s"]), ptr(out["flag"]), ptr(out["iters"]), None, None, st)

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        times = []
        for _ in range(a.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        res = {k: v.cpu().numpy() for k, v in out.items()}
        results[name] = res
        print(json.dumps({"variant": name, "kind": a.kind, "batch": B, "ms": ms, "min_ms": min(times), "solves_per_s": B / (ms * 1e-3),
                          "kernel": solver.last_kernel, "generated": solver.last_kernel_generated, "launches": solver.last_launches,
                          "iters_mean": float(res["iters"].mean()), "iters_max": int(res["iters"].max()),
                          "not_optimal": int((res["flag"] != 0).sum())}), flush=True)
        solver.close()
    names = list(results)
    for other in names[1:]:
        same = all(np.array_equal(results[names[0]][k], results[other][k]) for k in ("x", "z", "s", "flag", "iters"))
        print(f"{names[0]} == {other}: {same}", flush=True)
    if a.oracle:
        from oracle import reference
        cnt = min(a.oracle, B)
        ref, _ = reference.run_batch(batch.slice(0, cnt))
        for name, res in results.items():
            ok = all(np.array_equal(res[k][:cnt], ref[k]) for k in ("x", "z", "s", "flag", "iters")) and (pat.p == 0 or np.array_equal(res["y"][:cnt], ref["y"]))
            print(f"{name} == oracle on the first {cnt}: {ok}", flush=True)


if __name__ == "__main__":
    main()
