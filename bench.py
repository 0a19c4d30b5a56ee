#!/usr/bin/env python
"""bench.py -- batched DogBot WBC QP solves/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config all|trot65k|mixed1m|mpc131k|rollout]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch.  The HEADLINE (top-level keys of the line) is always
BASELINE.json configs[1]: 65,536 identical-pattern DogBot trot-gait whole-body-control QPs with randomised
states per GPU (weak scaling: every rank solves its own contiguous shard of instance indices; QPs are
independent, so the only collective is the gather of solutions and exit flags the north_star names).

  value      whole-job solves/s, value rows already resident in HBM, CUDA events around K steps, max over ranks.
  e2e        the same metric through the reference-facing C ABI call QP_SOLVE_BATCH with HOST buffers (pinned):
             H2D of the value rows and D2H of x, y, z, s, flags, iterations inside the timed region.
  roofline   the dominant kernel against the measured HBM peak: "achieved" = SURVEY.md 8d's ALGORITHMIC bytes per
             solve / launch time; "traffic" = the kernel's real DRAM bytes per launch from the committed ncu capture
             (profiles/ncu_traffic.json).  "fp64": the same launch against the FP64 peak MEASURED in this run
             (QPB_FP64_PEAK, fused and unfused), with SURVEY.md 8d's algorithmic flops.
  latency    BASELINE.json configs[0]: ONE stance QP per control tick through QP_SETUP_dense -> QP_SOLVE ->
             QP_CLEANUP_dense (what dogbot_controller does at 400 Hz, main.cpp:1649-1663), p50 / p99 over 1,000
             ticks, this library next to the compiled reference on the same box.
  configs    the other BASELINE.json configurations, each with value / e2e / roofline / iteration histogram and a
             1,024-instance oracle sample checked inside the run:
               mixed1m   configs[2]: 1,048,576 QPs, thirds of stance / crawl / trot, three plans on three streams
                         per GPU, every pattern split across all ranks (STRONG scaling);
               mpc131k   configs[4]: 131,072 horizon-10 MPC QPs per GPU (weak scaling);
               rollout   configs[3]: robots x sequential control ticks on the device (formulation -> solve ->
                         torques -> integrate), ticks/s.
  cpu_baseline  the reference qpSWIFT (oracle/_ref, kind "reference"; else our port) on the box's host cores, one
             instance per pinned thread, on the same workload; extras: controller style (QP_SETUP_dense with AMD per
             QP) and the -O3 build.

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
MIXED_TOTAL = 3 * 349528  # 1,048,584 >= 2^20: thirds that split evenly over 1, 2, 4 and 8 GPUs
MPC_PER_GPU = 131072
ORACLE_SAMPLE = 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="all", help="all | trot65k (headline only) | mixed1m | mpc131k | rollout (comma separated: the extras run next to the headline)")
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="QPs per GPU per step (headline)")
    ap.add_argument("--kind", default=KIND)
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances per CPU-baseline step (0: the whole batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernels", action="store_true", help="skip the stand-alone factor/solve kernel timings")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--mixed-total", type=int, default=MIXED_TOTAL)
    ap.add_argument("--mpc-per-gpu", type=int, default=MPC_PER_GPU)
    ap.add_argument("--robots-per-gpu", type=int, default=512)
    ap.add_argument("--ticks", type=int, default=1000)
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


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the solve kernel on this workload,
    from the ncu --set full capture summarised under profiles/ (null when no capture matches)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, None
    rec = json.load(open(path)).get(key)
    return (float(rec["dram_bytes"]), rec.get("source")) if rec else (None, None)


def algorithmic(plan, pat, iters):
    """SURVEY.md section 8d.  Bytes per solve (it+1) B_factor + (2 it+1) B_solve + it B_vec; flops per solve
    (it+1) F_factor + (2 it+1)(4 Lnz + N) + 2 it (nnzP + 2 nnzA + 2 nnzG + 12 N).  Summed over the batch."""
    N, lnz, nnzU = plan.info("N"), plan.info("lnz"), plan.info("nnzU")
    b_factor = 8 * (nnzU + lnz + N)
    b_solve = 8 * (lnz + 3 * N)
    b_vec = 8 * (pat.nnzP + pat.nnzA + pat.nnzG + 14 * N)
    it = np.asarray(iters, dtype=np.float64)
    total = float(np.sum((it + 1) * b_factor + (2 * it + 1) * b_solve + it * b_vec))
    flops = float(np.sum((it + 1) * plan.info("factor_flops") + (2 * it + 1) * (4 * lnz + N) + 2 * it * (pat.nnzP + 2 * pat.nnzA + 2 * pat.nnzG + 12 * N)))
    return total, flops, b_factor, b_solve, b_vec


def histogram(iters):
    v, c = np.unique(np.asarray(iters), return_counts=True)
    return {str(int(a)): int(b) for a, b in zip(v, c)}


def cpu_arm(kind_name, batch, steps, warmup, oracle_kind=None, note=""):
    """QP_SETUP(shared AMD permutation) + QP_SOLVE + QP_CLEANUP per instance, one instance per pinned host thread."""
    from oracle import reference as R
    kind = oracle_kind or R.best_kind()
    threads = os.cpu_count() or 1
    for _ in range(warmup):
        R.run_batch(batch.slice(0, min(batch.count, 2048)), kind=kind, threads=threads, outputs=False)
    t = 0.0
    for _ in range(steps):
        _, secs = R.run_batch(batch, kind=kind, threads=threads, outputs=False)
        t += secs
    value = batch.count * steps / t
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "reference" if kind.startswith("reference") else "port",
            "sample": f"{batch.count} {kind_name} instances per step, {steps} steps, threads pinned; QP_SETUP(shared AMD permutation)+QP_SOLVE+QP_CLEANUP per instance, one instance per host thread{note}"}, t / steps * 1e3


def cpu_extras(batch):
    """Labelled extras of BASELINE.md section 4: controller style (QP_SETUP_dense: densetosparse + AMD + symbolic
    analysis per QP, main.cpp:1649) and the -O3 -march=x86-64-v3 build; bounded samples."""
    from oracle import reference as R
    out = {}
    threads = os.cpu_count() or 1
    if R.available("reference"):
        sample = batch.slice(0, min(batch.count, 8192))
        R.run_batch_dense(sample.slice(0, 512), threads=threads, outputs=False)
        _, secs = R.run_batch_dense(sample, threads=threads, outputs=False)
        out["controller_style"] = {"value": sample.count / secs, "unit": UNIT, "cores": threads,
                                   "sample": f"{sample.count} instances, QP_SETUP_dense (AMD per QP) + QP_SOLVE + QP_CLEANUP_dense, tol 1e-6"}
    if os.path.exists(R.REF_O3_SO):
        sample = batch.slice(0, min(batch.count, 16384))
        R.run_batch(sample.slice(0, 2048), kind="reference_O3", threads=threads, outputs=False)
        _, secs = R.run_batch(sample, kind="reference_O3", threads=threads, outputs=False)
        out["O3_x86-64-v3"] = {"value": sample.count / secs, "unit": UNIT, "cores": threads,
                               "sample": f"{sample.count} instances, plan reuse; reference compiled -O3 -march=x86-64-v3 (the portable stand-in for -march=native: the binary is built off the box)"}
    return out


def oracle_check(batch, out, count=ORACLE_SAMPLE, what=""):
    """A spread sample of the batch against the CPU oracle, bit for bit (inside the run: a wrong fast answer is no answer)."""
    from oracle import reference as R
    from apf_quadruped_b200 import dogbot
    n = batch.count
    idx = np.unique(np.linspace(0, n - 1, min(count, n)).astype(np.int64))
    sub = dogbot.Batch(batch.pattern, *[np.ascontiguousarray(getattr(batch, k)[idx]) for k in ("Ppr", "Apr", "Gpr", "c", "h", "b")])
    ref, _ = R.run_batch(sub)
    keys = ("x", "y", "z", "s") if batch.pattern.p else ("x", "z", "s")
    same = bool(np.array_equal(ref["flag"], np.asarray(out["flag"])[idx].astype(np.int64)) and np.array_equal(ref["iters"], np.asarray(out["iters"])[idx].astype(np.int64))
                and all(np.array_equal(ref[k], np.asarray(out[k])[idx]) for k in keys))
    if not same:
        raise AssertionError(f"{what}: results differ from the {R.best_kind()} oracle on a {len(idx)}-instance sample")
    return {"instances": int(len(idx)), "oracle": R.best_kind(), "bit_identical": same}


def latency_leg(qpswift, dogbot, ticks=1000):
    """BASELINE.json configs[0] as the controller runs it: one stance QP per tick through QP_SETUP_dense -> QP_SOLVE
    -> x -> QP_CLEANUP_dense with the controller's tolerances (main.cpp:1649-1663), this library (called through
    ctypes from here) next to the compiled reference (oracle/ref_driver.c) on the same box, same inputs."""
    import ctypes as C
    from oracle import reference as R
    batch = dogbot.generate("stance", 256)
    pat = batch.pattern
    P, A, G = [np.ascontiguousarray(a) for a in batch.dense_batch()]
    lb = qpswift.lib()
    us, xs, its = np.empty(ticks), np.empty((ticks, pat.n)), np.empty(ticks, dtype=np.int64)
    kernel_us, setup_us = np.empty(ticks), np.empty(ticks)
    PD = C.POINTER(C.c_double)
    for t in range(-20, ticks):  # (20 untimed ticks: plan cache, first launch)
        i = t % batch.count
        bufs = [np.array(a[i], dtype=np.float64) for a in (P, A, G, batch.c, batch.h, batch.b)]
        ptrs = [b_.ctypes.data_as(PD) for b_ in bufs]
        t0 = time.perf_counter_ns()
        qp = lb.QP_SETUP_dense(pat.n, pat.m, pat.p, ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4], ptrs[5], None, 30)
        qp.contents.options.contents.reltol = 1e-2
        qp.contents.options.contents.abstol = 1e-2
        lb.QP_SOLVE(qp)
        x = np.ctypeslib.as_array(qp.contents.x, shape=(pat.n,)).copy()
        t1 = time.perf_counter_ns()
        if t >= 0:
            us[t], xs[t], its[t] = (t1 - t0) * 1e-3, x, qp.contents.stats.contents.IterationCount
            kernel_us[t], setup_us[t] = qp.contents.stats.contents.kkt_time * 1e6, qp.contents.stats.contents.tsetup * 1e6
        lb.QP_CLEANUP_dense(qp)
    out = {"workload": "configs[0]: one DogBot stance WBC QP per control tick (n30 m68 p18), QP_SETUP_dense -> QP_SOLVE -> x -> QP_CLEANUP_dense, reltol = abstol = 1e-2 (main.cpp:1649-1663)",
           "ticks": ticks, "unit": "us", "b200": {"p50": float(np.percentile(us, 50)), "p99": float(np.percentile(us, 99)), "mean": float(us.mean()),
                    "kernel_p50": float(np.percentile(kernel_us, 50)), "setup_p50": float(np.percentile(setup_us, 50)), "iterations_mean": float(its.mean())},
           "budget_us": 2500.0}
    if R.available("reference"):
        ref = R.control_loop(batch, ticks)
        out["reference_cpu"] = {"p50": float(np.percentile(ref["us"], 50)), "p99": float(np.percentile(ref["us"], 99)), "mean": float(ref["us"].mean()),
                                "setup_p50": float(np.percentile(ref["setup_us"], 50))}
        out["bit_identical_with_reference"] = bool(np.array_equal(ref["x"], xs) and np.array_equal(ref["iters"], its))
        if not out["bit_identical_with_reference"]:
            raise AssertionError("latency leg: QP_SETUP_dense + QP_SOLVE differs from the compiled reference")
    return out


class QuietStdout:
    """NCCL prints its version banner (and, at INFO, everything else) on the C-level stdout; this run's stdout carries
    exactly one JSON line.  While active, file descriptor 1 points at `path`; restore() puts the real stdout back."""

    def __init__(self, path):
        sys.stdout.flush()
        self.saved = os.dup(1)
        fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_APPEND, 0o644)
        os.dup2(fd, 1)
        os.close(fd)

    def restore(self):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def nccl_logging(rank):
    """NCCL's own evidence of the communicator (ranks, transports) goes to files next to the other records.  A caller
    who has asked for NCCL_DEBUG=INFO / TRACE themselves keeps exactly what they asked for (on their stdout).
    Returns (log file pattern or None, QuietStdout or None)."""
    if os.environ.get("NCCL_DEBUG", "").upper() in ("INFO", "TRACE"):
        return os.environ.get("NCCL_DEBUG_FILE"), None
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
    except OSError:
        out = "/tmp"
    os.environ["NCCL_DEBUG"] = "INFO"
    os.environ["NCCL_DEBUG_SUBSYS"] = "INIT,ENV"
    os.environ["NCCL_DEBUG_FILE"] = os.path.join(out, "nccl_bench_%h_%p.log")
    return os.environ["NCCL_DEBUG_FILE"], QuietStdout(os.path.join(out, f"nccl_bench_stdout_rank{rank}.log"))


def nccl_summary(pattern, world):
    """What rank 0's NCCL log says about the communicator."""
    import glob
    import re
    if not pattern:
        return None
    files = sorted(glob.glob(pattern.replace("%h", "*").replace("%p", "*")), key=os.path.getmtime)[-world:]
    ranks, via = set(), set()
    for f in files:
        try:
            txt = open(f, errors="replace").read()
        except OSError:
            continue
        for m in re.finditer(r"nranks (\d+)", txt):
            ranks.add(int(m.group(1)))
        for key in ("NVLS", "P2P/CUMEM", "P2P/IPC", "via SHM", "NET/"):
            if key in txt:
                via.add(key)
    return {"nranks_seen": sorted(ranks), "comm_nranks_ok": world in ranks, "transports": sorted(via), "log_files": len(files)}


class Leg:
    """One pattern on this rank: plan + solver, value rows pinned on the host and resident in HBM, outputs."""

    def __init__(self, torch, qpswift, batch, dev, local, stream=None):
        self.torch, self.batch, self.pat, self.B = torch, batch, batch.pattern, batch.count
        pat, B = self.pat, self.B
        self.solver = qpswift.BatchSolver(pat, max(B, 1), device=local)
        self.host = {k: torch.from_numpy(getattr(batch, k)).pin_memory() for k in ("Ppr", "Apr", "Gpr", "c", "h", "b")}
        self.dev = {k: v.to(dev, non_blocking=True) for k, v in self.host.items()}
        self.out = {"x": torch.empty((B, pat.n), dtype=torch.float64, device=dev), "y": torch.empty((B, max(pat.p, 1)), dtype=torch.float64, device=dev),
                    "z": torch.empty((B, pat.m), dtype=torch.float64, device=dev), "s": torch.empty((B, pat.m), dtype=torch.float64, device=dev),
                    "flag": torch.empty(B, dtype=torch.int32, device=dev), "iters": torch.empty(B, dtype=torch.int32, device=dev)}
        self.stream = stream
        self.hout = None

    def launch(self, opts, st=None):
        if self.B == 0:
            return
        p = lambda t: t.data_ptr() if t.numel() else None
        d, o = self.dev, self.out
        st = st if st is not None else (self.stream.cuda_stream if self.stream is not None else self.torch.cuda.current_stream().cuda_stream)
        self.solver.solve_device(self.B, opts, p(d["Ppr"]), p(d["Apr"]), p(d["Gpr"]), p(d["c"]), p(d["h"]), p(d["b"]),
                                 p(o["x"]), p(o["y"]) if self.pat.p else None, p(o["z"]), p(o["s"]), p(o["flag"]), p(o["iters"]), None, None, st)

    def host_solve(self, opts, dogbot):
        torch, pat, B = self.torch, self.pat, self.B
        if self.hout is None:
            self.hout = {"x": torch.empty((B, pat.n), dtype=torch.float64).pin_memory(), "y": torch.empty((B, pat.p), dtype=torch.float64).pin_memory(),
                         "z": torch.empty((B, pat.m), dtype=torch.float64).pin_memory(), "s": torch.empty((B, pat.m), dtype=torch.float64).pin_memory()}
            self.out_np = {k: v.numpy() for k, v in self.hout.items()}
            self.out_np["flag"] = np.empty(B, dtype=np.int64)
            self.out_np["iters"] = np.empty(B, dtype=np.int64)
            self.hbatch = dogbot.Batch(pat, *[self.host[k].numpy() for k in ("Ppr", "Apr", "Gpr", "c", "h", "b")])
        if B:
            self.solver.solve(self.hbatch, opts, out=self.out_np)

    def results(self):
        return {k: v.cpu().numpy() for k, v in self.out.items()}

    def bytes_h2d(self):
        pat = self.pat
        return 8 * self.B * (pat.nnzP + pat.nnzA + pat.nnzG + pat.n + pat.m + pat.p)

    def bytes_d2h(self):
        pat = self.pat
        return 8 * self.B * (pat.n + pat.p + 2 * pat.m) + 8 * self.B

    def check_host_against_device(self):
        res = self.results()
        assert np.array_equal(self.out_np["iters"], res["iters"].astype(np.int64)), "host-pointer and device-pointer paths disagree"
        for k in ("x", "y", "z", "s") if self.pat.p else ("x", "z", "s"):
            assert np.array_equal(self.out_np[k], res[k]), f"host-pointer and device-pointer paths disagree on {k}"
        return res

    def close(self):
        self.solver.close()


def timed(torch, dist, world, dev, steps, warmup, step_fn):
    """W untimed + K timed steps, CUDA events on the current stream, barrier + synchronize on both sides, max over ranks."""
    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    for _ in range(max(warmup, 3)):
        step_fn()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_fn()
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


def timed_host(torch, dist, world, dev, steps, step_fn):
    for _ in range(2):
        step_fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_fn()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


def gather_step(dist, gathered, leg):
    """The only collective on the path (north_star): solutions x (what the controller reads, main.cpp:1663), exit flags
    and iteration counts, straight from the solver's output arrays; the duals stay on the rank that computed them."""
    dist.all_gather_into_tensor(gathered["x"], leg.out["x"])
    dist.all_gather_into_tensor(gathered["flag"], leg.out["flag"])
    dist.all_gather_into_tensor(gathered["iters"], leg.out["iters"])


def run_mixed(args, env, torch, dist, qpswift, dogbot, opts):
    """configs[2]: thirds of stance / crawl / trot, one plan per pattern, every pattern split across all ranks (so that
    every GPU sees the same mix), three streams per GPU.  STRONG scaling: the total is fixed."""
    from apf_quadruped_b200 import shard
    rank, world, local, dev = env
    total = args.mixed_total
    third = total // 3
    counts = {"stance": total - 2 * third, "crawl": third, "trot": third}
    assert all(c % world == 0 for c in counts.values()), "every pattern's share must split evenly over the ranks (the gather takes equal shards)"
    mine = shard.mixed_shards(counts, world, rank)
    legs = {}
    for i, (kind, (lo, hi)) in enumerate(mine.items()):
        batch = dogbot.generate_parallel(kind, hi - lo, start=lo, seed=dogbot.DEFAULT_SEED + i)
        legs[kind] = Leg(torch, qpswift, batch, dev, local, stream=torch.cuda.Stream(device=dev))
    gathered = None
    if world > 1:
        gathered = {k: {"x": torch.empty((world * l.B, l.pat.n), dtype=torch.float64, device=dev), "flag": torch.empty(world * l.B, dtype=torch.int32, device=dev),
                        "iters": torch.empty(world * l.B, dtype=torch.int32, device=dev)} for k, l in legs.items()}

    def step():
        cur = torch.cuda.current_stream()
        for l in legs.values():
            l.stream.wait_stream(cur)
            l.launch(opts)
        for k, l in legs.items():
            cur.wait_stream(l.stream)
        if world > 1:
            for k, l in legs.items():
                gather_step(dist, gathered[k], l)

    ms = timed(torch, dist, world, dev, args.steps, args.warmup, step)
    value = total / (ms * 1e-3)

    def host_step():
        ths = [threading.Thread(target=l.host_solve, args=(opts, dogbot)) for l in legs.values()]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    e2e_s = timed_host(torch, dist, world, dev, args.steps, host_step)
    rec = {"workload": f"configs[2]: {total} QPs, thirds of stance / crawl / trot (one symbolic plan per pattern), every pattern split across {world} GPU(s), three streams per GPU",
           "scaling": "strong", "value": value, "unit": UNIT, "ms_per_step": ms, "n_gpus": world,
           "e2e": {"value": total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": sum(l.bytes_h2d() for l in legs.values()), "d2h_bytes_per_step": sum(l.bytes_d2h() for l in legs.values())},
           "per_pattern": {}, "gpu_launches": args.steps * sum(l.solver.last_launches for l in legs.values())}
    tot_bytes = tot_flops = 0.0
    for k, l in legs.items():
        res = l.check_host_against_device()
        b, f, *_ = algorithmic(l.solver.plan, l.pat, res["iters"])
        tot_bytes += b
        tot_flops += f
        rec["per_pattern"][k] = {"qps_this_rank": l.B, "kernel": l.solver.last_kernel, "generated": l.solver.last_kernel_generated, "iterations_mean": float(res["iters"].mean()),
                                 "iterations_max": int(res["iters"].max()), "not_optimal": int((res["flag"] != 0).sum()), "iterations": histogram(res["iters"]),
                                 "oracle_sample": oracle_check(l.batch, res, ORACLE_SAMPLE // 3, f"mixed/{k}") if rank == 0 else None}
    t = torch.tensor([tot_bytes, tot_flops], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    peak, src = peak_hbm()
    ach = float(t[0].item()) / (ms * 1e-3) / 1e9
    rec["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / (peak * world), "peak_source": src, "traffic": None,
                       "note": "three solve kernels run concurrently; algorithmic bytes of all three over the step time, against the HBM peak of all GPUs"}
    rec["fp64_tflops_algorithmic"] = float(t[1].item()) / (ms * 1e-3) / 1e12
    for l in legs.values():
        l.close()
    return rec


def run_mpc(args, env, torch, dist, qpswift, dogbot, opts):
    """configs[4]: horizon-10 single-rigid-body MPC QP (n240 m240 p120; ours), 131,072 per GPU, weak scaling."""
    rank, world, local, dev = env
    B = args.mpc_per_gpu
    batch = dogbot.generate_parallel("mpc", B, start=rank * B)
    leg = Leg(torch, qpswift, batch, dev, local)
    ms = timed(torch, dist, world, dev, args.steps, args.warmup, lambda: leg.launch(opts))
    e2e_s = timed_host(torch, dist, world, dev, args.steps, lambda: leg.host_solve(opts, dogbot))
    res = leg.check_host_against_device()
    b, f, *_ = algorithmic(leg.solver.plan, leg.pat, res["iters"])
    peak, src = peak_hbm()
    ach = b / (ms * 1e-3) / 1e9
    traffic, tsrc = ncu_traffic(f"mpc:{B}")
    rec = {"workload": f"configs[4]: horizon-10 sparse MPC QP for the DogBot centroidal model (n240 m240 p120, N = 600), {B} per GPU",
           "scaling": "weak", "value": world * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "n_gpus": world,
           "e2e": {"value": world * B / e2e_s, "unit": UNIT, "h2d_bytes_per_step": leg.bytes_h2d(), "d2h_bytes_per_step": leg.bytes_d2h()},
           "kernel": leg.solver.last_kernel, "generated": leg.solver.last_kernel_generated, "gpu_launches": args.steps * leg.solver.last_launches,
           "iterations_mean": float(res["iters"].mean()), "iterations_max": int(res["iters"].max()), "not_optimal": int((res["flag"] != 0).sum()), "iterations": histogram(res["iters"]),
           "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_source": tsrc, "peak_source": src},
           "fp64_tflops_algorithmic": f / (ms * 1e-3) / 1e12,
           "oracle_sample": oracle_check(batch, res, 256, "mpc") if rank == 0 else None}
    leg.close()
    return rec


def run_rollout(args, env, torch, dist, qpswift, dogbot, opts):
    """configs[3]: robots x sequential control ticks, everything on the device (apf_quadruped_b200/rollout.py)."""
    from apf_quadruped_b200 import rollout
    from oracle import rollout_reference
    return rollout.bench(args, env, torch, dist, opts, check=rollout_reference.check_device_rollout)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    want = {c.strip() for c in args.config.split(",") if c.strip()}
    if "all" in want:
        want = {"trot65k", "mixed1m", "mpc131k"} | ({"rollout"} if os.path.exists(os.path.join(ROOT, "apf_quadruped_b200", "rollout.py")) else set())

    from apf_quadruped_b200 import dogbot

    if args.impl == "reference":
        if rank != 0:
            return
        sample = args.cpu_sample or args.batch
        batch = dogbot.generate_parallel(args.kind, min(sample, args.batch))
        base, ms = cpu_arm(args.kind, batch, max(1, args.steps), args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "kind": args.kind, "batch_per_gpu": args.batch, "global_batch": args.batch, "sample_per_step": batch.count,
                                                                 "note": "the reference's own CPU implementation solves ONE box's worth of the workload (the whole 65,536-QP batch) per step on all host cores, whatever N"},
                "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        try:
            line["cpu_baseline_extras"] = cpu_extras(batch)
        except Exception as e:  # the extras never cost the line
            line["cpu_baseline_extras"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
        return

    B = args.batch
    batch = dogbot.generate_parallel(args.kind, B, start=rank * B)  # (before CUDA is touched: the generator forks workers)

    import torch
    import torch.distributed as dist
    from apf_quadruped_b200 import qpswift

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nccl_log = quiet = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        nccl_log, quiet = nccl_logging(rank)
        dist.init_process_group("nccl", device_id=dev)
    env = (rank, world, local, dev)

    pat = batch.pattern
    opts = qpswift.default_settings()
    leg = Leg(torch, qpswift, batch, dev, local)
    gathered = None
    if world > 1:
        gathered = {"x": torch.empty((world * B, pat.n), dtype=torch.float64, device=dev), "flag": torch.empty(world * B, dtype=torch.int32, device=dev),
                    "iters": torch.empty(world * B, dtype=torch.int32, device=dev)}

    def device_step():
        leg.launch(opts)
        if world > 1:
            gather_step(dist, gathered, leg)

    # nvidia-smi takes a driver-wide lock: eight ranks polling it stall each other's launches, so only rank 0 samples
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step = timed(torch, dist, world, dev, args.steps, args.warmup, device_step)
    value = world * B / (ms_step * 1e-3)
    # ---- e2e through the C ABI with host buffers ----------------------------------------------------
    e2e_s = timed_host(torch, dist, world, dev, args.steps, lambda: leg.host_solve(opts, dogbot))
    e2e_value = world * B / e2e_s
    clocks = sampler.summary() if rank == 0 else None
    res = leg.check_host_against_device()
    iters, flags = res["iters"], res["flag"]
    if world > 1:  # the gathered rows of this rank are this rank's rows
        g = gathered["x"][rank * B:(rank + 1) * B].cpu().numpy()
        assert np.array_equal(g, res["x"]), "gathered solutions differ from the rank's own"

    extras = {}
    line = None
    if rank == 0:
        peak, peak_src = peak_hbm()
        total_bytes, total_flops, b_factor, b_solve, b_vec = algorithmic(leg.solver.plan, pat, iters)
        achieved = total_bytes / (ms_step * 1e-3) / 1e9
        kname = {"stream": "qp_solve_stream_kernel", "warp": "qp_solve_warp_kernel"}[leg.solver.last_kernel] + ("<generated %s>" % args.kind if leg.solver.last_kernel_generated else "<interpreted>")
        traffic, traffic_src = ncu_traffic(f"{args.kind}:{B}:{leg.solver.last_kernel}")
        roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_solve": total_bytes / B, "traffic_gbs": traffic / (ms_step * 1e-3) / 1e9 if traffic else None, "traffic_source": traffic_src,
                "note": ("one QP per warp, the QP's working set resident on chip from load to store (factor L and iterate in shared memory, value row in an L2-resident scratch "
                         "row): 'achieved' counts SURVEY 8d's algorithmic bytes (factor written once and read once per solve) although those bytes never reach HBM here -- 'traffic' "
                         "(value rows in, solutions out, scratch rows that L2 evicted) is what does"
                         if leg.solver.last_kernel == "warp" else
                         "one QP per thread, working sets interleaved in HBM: 'achieved' = SURVEY 8d algorithmic bytes / launch time; the up-looking factorisation re-reads L row by row, "
                         "which is the gap between 'traffic' and the algorithmic bytes")}
        try:
            fused, unfused = qpswift.lib().QPB_FP64_PEAK(local, 1), qpswift.lib().QPB_FP64_PEAK(local, 0)
            ach = total_flops / (ms_step * 1e-3) / 1e12
            roof["fp64"] = {"achieved": ach, "unit": "TFLOP/s", "peak_fused": fused, "peak_unfused": unfused, "frac": ach / unfused, "frac_of_fused_peak": ach / fused,
                            "algorithmic_flops_per_solve": total_flops / B,
                            "note": "peaks measured in this run (QPB_FP64_PEAK: 8 independent chains per thread, DFMA vs DMUL+DADD); the kernels run unfused arithmetic (-fmad=false: bit-exactness), so the usable peak is the unfused one"}
        except Exception as e:
            roof["fp64"] = {"error": repr(e)}
        kernels = {}
        if not args.no_kernels:
            plan = leg.solver.plan
            N, lnz = plan.info("N"), plan.info("lnz")
            devin = leg.dev
            vals = torch.cat([devin["Ppr"], devin["Apr"], devin["Gpr"]], dim=1).contiguous()
            zd = torch.full((B, pat.m), -1.0, dtype=torch.float64, device=dev)
            Lx = torch.empty((B, lnz), dtype=torch.float64, device=dev)
            Dd = torch.empty((B, N), dtype=torch.float64, device=dev)
            rhs = torch.randn((B, N), dtype=torch.float64, device=dev)
            sol = torch.empty_like(rhs)
            st = torch.cuda.current_stream().cuda_stream
            solver = leg.solver
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
            # dense front-end: dense column-major P | A | G -> plan-ordered value rows
            dense = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in batch.dense_batch()]
            rows_out = [torch.empty((B, w), dtype=torch.float64, device=dev) for w in (pat.nnzP, pat.nnzA, pat.nnzG)]
            viol = torch.empty((B,), dtype=torch.int32, device=dev)
            ptr = lambda t: t.data_ptr() if t.numel() else None
            fn = lambda: solver.dense_to_values_device(B, 30, *[ptr(t) for t in dense], *[ptr(t) for t in rows_out], viol.data_ptr(), st)
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
            gbs = nbytes * B / (kms * 1e-3) / 1e9
            assert all(torch.equal(r, d) for r, d in zip(rows_out, (devin["Ppr"], devin["Apr"], devin["Gpr"]))) and not bool(viol.any())
            kernels["dense_gather_kernel"] = {"ms": kms, "qps": B, "bytes_per_qp": nbytes, "achieved": gbs, "frac": gbs / peak, "per_s": B / (kms * 1e-3)}
            # WBC formulation on the device: packed robot states -> plan-ordered value rows | c | h | b.  Bytes:
            # the state row read, every value written once.
            from apf_quadruped_b200 import wbc
            if args.kind in wbc.KINDS:
                states = torch.from_numpy(wbc.random_states(B)).to(dev)
                wr = [torch.empty((B, w), dtype=torch.float64, device=dev) for w in (pat.nnzP, pat.nnzA, pat.nnzG, pat.n, pat.m, pat.p)]
                form = wbc.Formulator(solver, args.kind)
                fn = lambda: form.formulate(B, states.data_ptr(), *[t.data_ptr() for t in wr], stream=st)
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
                nbytes = 8 * (wbc.STATE + pat.nnzP + pat.nnzA + pat.nnzG + pat.n + pat.m + pat.p)
                gbs = nbytes * B / (kms * 1e-3) / 1e9
                ref_rows = wbc.value_rows(args.kind, states[:256].cpu().numpy())
                assert all(np.array_equal(t[:256].cpu().numpy(), getattr(ref_rows, k)) for t, k in zip(wr, ("Ppr", "Apr", "Gpr", "c", "h", "b"))), "formulation kernel differs from the numpy restatement"
                kernels["wbc_formulate_kernel"] = {"ms": kms, "qps": B, "bytes_per_qp": nbytes, "achieved": gbs, "frac": gbs / peak, "per_s": B / (kms * 1e-3)}
                del states, wr
            del dense, rows_out, vals, Lx, Dd, rhs, sol
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "kind": args.kind, "batch_per_gpu": B, "global_batch": world * B,
                           "sharding": f"dp{world} contiguous shards, all-gather of x + flags + iterations" if world > 1 else "single GPU",
                           "l2": "inputs (value rows, %.0f MB per GPU) exceed the 126 MB L2, no explicit flush" % (leg.bytes_h2d() / 1e6),
                           "iterations_mean": float(iters.mean()), "iterations_max": int(iters.max()), "not_optimal": int((flags != 0).sum()), "iterations": histogram(iters)},
                "clocks": clocks, "roofline": roof, "kernels": kernels,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": leg.bytes_h2d(), "d2h_bytes_per_step": leg.bytes_d2h()},
                "gpu_launches": args.steps * leg.solver.last_launches,
                "oracle_sample": oracle_check(batch, res, ORACLE_SAMPLE, "headline")}
        if world > 1:
            line["nccl"] = nccl_summary(nccl_log, world)
    leg.close()
    del leg
    torch.cuda.empty_cache()

    # ---- the other configurations (every rank takes part; a failure costs the extra, never the headline) -----------
    for name, fn in (("mixed1m", run_mixed), ("mpc131k", run_mpc), ("rollout", run_rollout)):
        if name not in want:
            continue
        try:
            rec = fn(args, env, torch, dist, qpswift, dogbot, opts)
        except Exception as e:
            if world > 1:
                raise  # (ranks must not diverge in a collective)
            rec = {"error": repr(e)}
        extras[name] = rec
        torch.cuda.empty_cache()

    if rank == 0:
        line["configs"] = extras
        if world == 1 and not args.no_latency:
            try:
                line["latency"] = latency_leg(qpswift, dogbot)
            except Exception as e:
                line["latency"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            cb = batch if not args.cpu_sample else batch.slice(0, min(args.cpu_sample, B))
            line["cpu_baseline"], _ = cpu_arm(args.kind, cb, 2, 1)
            try:
                line["cpu_baseline_extras"] = cpu_extras(batch)
            except Exception as e:
                line["cpu_baseline_extras"] = {"error": repr(e)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if quiet is not None:
        quiet.restore()
    if rank == 0:
        print(json.dumps(line), flush=True)  # (the last thing on stdout, after NCCL has said all it has to say)


if __name__ == "__main__":
    main()
