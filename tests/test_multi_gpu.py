"""The multi-GPU split behind the C ABI (QP_SETUP_BATCH_MULTI / QP_SOLVE_BATCH_MULTI, csrc/multi.cpp), driven by a
compiled C program that links libqpswift_b200.so -- the way the reference's C++ node would call it
(dogbot_controller/src/client/main.cpp:1649) -- and through the ctypes mirror."""
import os
import subprocess

import numpy as np
import pytest

from apf_quadruped_b200 import dogbot, qpswift

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "c", "multi_gpu_test")


def build_c_test():
    subprocess.run(["make", "-C", ROOT, "ctest"], check=True, capture_output=True)
    return EXE


def write_problem(path, batch):
    pat = batch.pattern
    with open(path, "wb") as f:
        np.array([pat.n, pat.m, pat.p, batch.count], dtype=np.int64).tofile(f)
        for a in (pat.Pjc, pat.Pir, pat.Ajc, pat.Air, pat.Gjc, pat.Gir):
            np.ascontiguousarray(a, dtype=np.int64).tofile(f)
        for a in (batch.Ppr, batch.Apr, batch.Gpr, batch.c, batch.h, batch.b):
            np.ascontiguousarray(a, dtype=np.float64).tofile(f)


def read_result(path, pat, count):
    raw = np.fromfile(path, dtype=np.uint8)
    at, out = 0, {}
    for k, w in (("x", pat.n), ("y", pat.p), ("z", pat.m), ("s", pat.m)):
        out[k] = raw[at:at + 8 * count * w].view(np.float64).reshape(count, w)
        at += 8 * count * w
    out["flag"] = raw[at:at + 8 * count].view(np.int64)
    out["iters"] = raw[at + 8 * count:at + 16 * count].view(np.int64)
    return out


def test_c_test_program_compiles_and_links():
    """(CPU) the C caller compiles against include/ and links against the library: every multi-GPU symbol resolves."""
    exe = build_c_test()
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("kind,count,shards", [("trot", 1500, 0), ("stance", 777, 3), ("crawl", 64, 8)])
def test_c_caller_splits_a_batch_over_devices(tmp_path, kind, count, shards):
    from oracle import reference
    exe = build_c_test()
    batch = dogbot.generate(kind, count)
    prob, res = str(tmp_path / "problem.bin"), str(tmp_path / "result.bin")
    write_problem(prob, batch)
    ndev = qpswift.lib().QPB_DEVICE_COUNT()
    shards = shards or max(ndev, 2)  # 0: one shard per device of the box (two shards on one device on a 1-GPU box)
    r = subprocess.run([exe, prob, res, str(shards)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bit-identical" in r.stdout
    out = read_result(res, batch.pattern, count)
    ref, _ = reference.run_batch(batch)
    assert np.array_equal(out["flag"], ref["flag"]) and np.array_equal(out["iters"], ref["iters"])
    for k in ("x", "y", "z", "s"):
        assert np.array_equal(out[k], ref[k]), f"{k} differs from the oracle"


@pytest.mark.gpu
def test_ctypes_mirror_of_the_multi_entries():
    from oracle import reference
    ndev = qpswift.lib().QPB_DEVICE_COUNT()
    devices = list(range(ndev)) if ndev > 1 else [0, 0, 0]
    batch = dogbot.generate("trot", 1001)
    ms = qpswift.MultiSolver(batch.pattern, 1001, devices)
    bounds = [ms.shard(1001, d) for d in range(len(devices))]
    assert bounds[0][0] == 0 and bounds[-1][1] == 1001 and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
    out = ms.solve(batch)
    ref, _ = reference.run_batch(batch)
    for k in ("x", "y", "z", "s", "flag", "iters"):
        assert np.array_equal(out[k], ref[k]), k
    with pytest.raises(RuntimeError):
        ms.solve(dogbot.generate("trot", 1002))
    ms.close()
