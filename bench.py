#!/usr/bin/env python
"""bench.py -- batched DogBot WBC QP solves/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: BASELINE.json configs[1], 65,536 identical-
pattern DogBot trot-gait whole-body-control QPs with randomised states, per GPU (weak scaling: every
rank solves its own contiguous shard of instance indices; QPs are independent, so the only collective
is the gather of solutions and exit flags the north_star names).

  value      whole-job solves/s, value rows already resident in HBM, CUDA events around K steps,
             max over ranks.
  e2e        the same metric through the reference-facing C ABI call QP_SOLVE_BATCH with HOST
             buffers (pinned): H2D of the value rows and D2H of x, y, z, s, flags, iterations inside
             the timed region (the library sends the rows up in chunks while the kernel already runs and
             lets the kernel write finished rows straight into the pinned result buffers; the bytes are
             the same, and the results are checked against the device-pointer path).
  roofline   dominant kernel (the throughput kernel qp_solve_stream_kernel: one QP per thread, working
             sets interleaved in HBM) against the measured HBM peak: "achieved" uses SURVEY.md section
             8d's ALGORITHMIC bytes per solve (see DESIGN.md "roofline"), "traffic" is the kernel's real
             DRAM traffic per launch from the committed ncu capture (profiles/ncu_traffic.json), and
             "traffic_gbs" that traffic over the live launch time; plus the stand-alone LDL factor /
             solve kernels and the dense front-end's gather kernel under "kernels".
  cpu_baseline  the reference qpSWIFT (oracle/_ref, kind "reference"; else our port) on the box's
             host cores, one instance per thread, on a bounded sample of the same workload.

--impl reference times only that CPU arm (rank 0; other ranks exit 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched QP solves/sec (DogBot WBC QP)"
UNIT = "solves/s"
KIND = "trot"
BATCH_PER_GPU = 65536
WORKLOAD = "configs[1]: 65,536 identical-pattern DogBot trot-gait WBC QPs (n30 m70 p12, 2 feet in contact), randomised states, tol 1e-6"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="QPs per GPU per step")
    ap.add_argument("--kind", default=KIND)
    ap.add_argument("--cpu-sample", type=int, default=16384, help="instances per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernels", action="store_true", help="skip the stand-alone factor/solve kernel timings")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.5)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(self.rows)}


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kind, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the solve kernel on this workload,
    from the ncu --set full capture summarised under profiles/ (null when no capture matches)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, None
    rec = json.load(open(path)).get(f"{kind}:{batch}")
    return (float(rec["dram_bytes"]), rec.get("source")) if rec else (None, None)


def algorithmic_bytes(plan, pat, iters):
    """per solve (it+1) B_factor + (2 it+1) B_solve + it B_vec."""
    N, lnz, nnzU = plan.info("N"), plan.info("lnz"), plan.info("nnzU")
    b_factor = 8 * (nnzU + lnz + N)
    b_solve = 8 * (lnz + 3 * N)
    b_vec = 8 * (pat.nnzP + pat.nnzA + pat.nnzG + 14 * N)
    it = np.asarray(iters, dtype=np.float64)
    total = float(np.sum((it + 1) * b_factor + (2 * it + 1) * b_solve + it * b_vec))
    return total, b_factor, b_solve, b_vec


def cpu_arm(args, steps, warmup):
    from apf_quadruped_b200 import dogbot
    from oracle import reference as R
    kind = R.best_kind()
    sample = dogbot.generate(args.kind, min(args.cpu_sample, args.batch))
    threads = os.cpu_count() or 1
    for _ in range(warmup):
        R.run_batch(sample.slice(0, min(sample.count, 2048)), kind=kind, threads=threads, outputs=False)
    t = 0.0
    for _ in range(steps):
        _, secs = R.run_batch(sample, kind=kind, threads=threads, outputs=False)
        t += secs
    value = sample.count * steps / t
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "reference" if kind == "reference" else "port",
            "sample": f"first {sample.count} instances of the workload per step, {steps} steps; QP_SETUP(shared AMD permutation)+QP_SOLVE+QP_CLEANUP per instance, one instance per host thread"}, t / steps * 1e3


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        base, ms = cpu_arm(args, max(1, args.steps), args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "sample_per_step": min(args.cpu_sample, args.batch)},
                "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    from apf_quadruped_b200 import dogbot, qpswift

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout at any NCCL_DEBUG level; stdout carries exactly one JSON line
        os.environ.pop("NCCL_DEBUG", None)
        if os.environ.get("BENCH_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = os.environ["BENCH_NCCL_DEBUG"]
            os.environ["NCCL_DEBUG_FILE"] = os.path.join(ROOT, "gpurun_out", "nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    batch = dogbot.generate(args.kind, B, start=rank * B)
    pat = batch.pattern
    solver = qpswift.BatchSolver(pat, B, device=local)
    opts = qpswift.default_settings()

    # pinned host copies (e2e arm) and HBM-resident copies (value arm)
    host = {k: torch.from_numpy(getattr(batch, k)).pin_memory() for k in ("Ppr", "Apr", "Gpr", "c", "h", "b")}
    devin = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    dout = {"x": torch.empty((B, pat.n), dtype=torch.float64, device=dev), "y": torch.empty((B, max(pat.p, 1)), dtype=torch.float64, device=dev),
            "z": torch.empty((B, pat.m), dtype=torch.float64, device=dev), "s": torch.empty((B, pat.m), dtype=torch.float64, device=dev),
            "flag": torch.empty(B, dtype=torch.int32, device=dev), "iters": torch.empty(B, dtype=torch.int32, device=dev)}
    gathered = None
    # what the north star gathers: the primal solutions x (what the controller reads, main.cpp:1663), exit
    # flags and iteration counts; the duals y, z, s stay on the rank that computed them (shard.gather moves
    # all of x | y | z | s | flag | iterations when a caller wants them)
    width = pat.n + 2
    if world > 1:
        gathered = torch.empty((world * B, width), dtype=torch.float64, device=dev)

    def ptr(t):
        return t.data_ptr() if t.numel() else None

    def device_step():
        st = torch.cuda.current_stream().cuda_stream
        solver.solve_device(B, opts, ptr(devin["Ppr"]), ptr(devin["Apr"]), ptr(devin["Gpr"]), ptr(devin["c"]), ptr(devin["h"]), ptr(devin["b"]),
                            ptr(dout["x"]), ptr(dout["y"]) if pat.p else None, ptr(dout["z"]), ptr(dout["s"]), ptr(dout["flag"]), ptr(dout["iters"]),
                            None, None, st)
        if world > 1:  # the only collective on the path: gather solutions + exit flags + iteration counts
            packed = torch.cat([dout["x"], dout["flag"].double()[:, None], dout["iters"].double()[:, None]], dim=1)
            dist.all_gather_into_tensor(gathered, packed)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        device_step()
    sync_all()
    # nvidia-smi takes a driver-wide lock: eight ranks polling it five times a second stall each other's
    # launches (measured: 193 ms instead of 75 ms per step), so only rank 0 samples, twice a second
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * B / (ms_step * 1e-3)
    iters = dout["iters"].cpu().numpy()
    flags = dout["flag"].cpu().numpy()

    # ---- e2e through the C ABI with host buffers ----------------------------------------------------
    hout = {"x": torch.empty((B, pat.n), dtype=torch.float64).pin_memory(), "y": torch.empty((B, pat.p), dtype=torch.float64).pin_memory(),
            "z": torch.empty((B, pat.m), dtype=torch.float64).pin_memory(), "s": torch.empty((B, pat.m), dtype=torch.float64).pin_memory()}
    out_np = {k: v.numpy() for k, v in hout.items()}
    out_np["flag"] = np.empty(B, dtype=np.int64)
    out_np["iters"] = np.empty(B, dtype=np.int64)
    hbatch = dogbot.Batch(pat, *[host[k].numpy() for k in ("Ppr", "Apr", "Gpr", "c", "h", "b")])
    for _ in range(2):
        solver.solve(hbatch, opts, out=out_np)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solver.solve(hbatch, opts, out=out_np)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())
    clocks = sampler.summary() if rank == 0 else None
    h2d = 8 * B * (pat.nnzP + pat.nnzA + pat.nnzG + pat.n + pat.m + pat.p)
    d2h = 8 * B * (pat.n + pat.p + 2 * pat.m) + 8 * B

    assert np.array_equal(out_np["iters"], iters.astype(np.int64)), "host-pointer and device-pointer paths disagree"
    for k in ("x", "y", "z", "s") if pat.p else ("x", "z", "s"):  # (the host path writes finished rows straight into the pinned buffers)
        assert np.array_equal(out_np[k], dout[k].cpu().numpy()), f"host-pointer and device-pointer paths disagree on {k}"

    if rank == 0:
        peak, peak_src = peak_hbm()
        total_bytes, b_factor, b_solve, b_vec = algorithmic_bytes(solver.plan, pat, iters)
        achieved = total_bytes / (ms_step * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(args.kind, B)
        stream = solver.last_kernel == "stream"
        roof = {"bound": "hbm", "kernel": "qp_solve_stream_kernel" if stream else "qp_solve_cta_kernel", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_solve": total_bytes / B,
                "traffic_gbs": traffic / (ms_step * 1e-3) / 1e9 if traffic else None, "traffic_source": traffic_src,
                "note": ("one QP per thread, working sets interleaved in HBM: 'achieved' = SURVEY 8d algorithmic bytes (factor written once, read once per "
                         "solve) / launch time; the up-looking factorisation re-reads L row by row, which is the gap between 'traffic' and the algorithmic bytes"
                         if stream else "one QP per warp, working set in shared memory: DRAM traffic is only the value rows in and the solution out")}
        kernels = {}
        if not args.no_kernels:
            plan = solver.plan
            N, lnz = plan.info("N"), plan.info("lnz")
            vals = torch.cat([devin["Ppr"], devin["Apr"], devin["Gpr"]], dim=1).contiguous()
            zd = torch.full((B, pat.m), -1.0, dtype=torch.float64, device=dev)
            Lx = torch.empty((B, lnz), dtype=torch.float64, device=dev)
            Dd = torch.empty((B, N), dtype=torch.float64, device=dev)
            rhs = torch.randn((B, N), dtype=torch.float64, device=dev)
            sol = torch.empty_like(rhs)
            st = torch.cuda.current_stream().cuda_stream
            for name, fn, nbytes in (("ldl_numeric_kernel", lambda: solver.ldl_numeric_device(B, vals.data_ptr(), zd.data_ptr(), Lx.data_ptr(), Dd.data_ptr(), st), b_factor),
                                     ("ldl_solve_kernel", lambda: solver.ldl_solve_device(B, Lx.data_ptr(), Dd.data_ptr(), rhs.data_ptr(), sol.data_ptr(), st), b_solve)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k0.record()
                for _ in range(args.steps):
                    fn()
                k1.record()
                torch.cuda.synchronize()
                kms = k0.elapsed_time(k1) / args.steps
                gbs = nbytes * B / (kms * 1e-3) / 1e9
                kernels[name] = {"ms": kms, "bytes_per_qp": nbytes, "achieved": gbs, "frac": gbs / peak, "per_s": B / (kms * 1e-3)}
            # dense front-end: dense column-major P | A | G -> plan-ordered value rows.  Bytes: every dense cell read,
            # every value written once.
            Bd = B
            dense = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in batch.slice(0, Bd).dense_batch()]
            rows_out = [torch.empty((Bd, w), dtype=torch.float64, device=dev) for w in (pat.nnzP, pat.nnzA, pat.nnzG)]
            viol = torch.empty((Bd,), dtype=torch.int32, device=dev)
            ptr = lambda t: t.data_ptr() if t.numel() else None
            fn = lambda: solver.dense_to_values_device(Bd, 30, *[ptr(t) for t in dense], *[ptr(t) for t in rows_out], viol.data_ptr(), st)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            for _ in range(args.steps):
                fn()
            k1.record()
            torch.cuda.synchronize()
            kms = k0.elapsed_time(k1) / args.steps
            nbytes = 8 * (pat.n * pat.n + pat.p * pat.n + pat.m * pat.n + pat.nnzP + pat.nnzA + pat.nnzG) + 4
            gbs = nbytes * Bd / (kms * 1e-3) / 1e9
            assert all(torch.equal(r, d) for r, d in zip(rows_out, (devin["Ppr"][:Bd], devin["Apr"][:Bd], devin["Gpr"][:Bd]))) and not bool(viol.any())
            kernels["dense_gather_kernel"] = {"ms": kms, "qps": Bd, "bytes_per_qp": nbytes, "achieved": gbs, "frac": gbs / peak, "per_s": Bd / (kms * 1e-3)}
            del dense, rows_out
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "kind": args.kind, "batch_per_gpu": B, "global_batch": world * B, "sharding": f"dp{world} contiguous shards, gather x+flags+iters" if world > 1 else "single GPU",
                           "l2": "inputs (value rows, %.0f MB per GPU) exceed the 126 MB L2, no explicit flush" % (h2d / 1e6),
                           "iterations_mean": float(iters.mean()), "iterations_max": int(iters.max()), "not_optimal": int((flags != 0).sum())},
                "clocks": clocks, "roofline": roof, "kernels": kernels,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": args.steps * solver.last_launches}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_arm(args, 2, 1)
        print(json.dumps(line), flush=True)
    solver.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
