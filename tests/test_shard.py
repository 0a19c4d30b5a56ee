"""The N > 1 path on CPU: contiguous shards + one all-gather (gloo, world_size 2 and 3).

Each rank "solves" its shard with the CPU oracle (test infrastructure) and the gathered result must
be identical, bit for bit and in instance order, to the oracle run on the whole batch in one
process: shard boundaries must not influence any QP, and the packing of flags / iteration counts
into the float64 payload must be exact.  The GPU variant of the same call is what bench.py runs
under torchrun with NCCL.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kind, total, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from apf_quadruped_b200 import dogbot, shard
    from oracle import reference
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_bounds(total, world)[rank]
    batch = dogbot.generate(kind, hi - lo, start=lo)
    local, _ = reference.run_batch(batch, threads=2)
    pat = batch.pattern
    got = shard.gather(local, total, {"x": pat.n, "y": pat.p, "z": pat.m, "s": pat.m}, dist, torch)
    if rank == 0:
        q.put(got)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 64), (3, 50)])
def test_sharded_solve_equals_single_process(oracle, world, total):
    import torch.multiprocessing as mp
    from apf_quadruped_b200 import dogbot
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + world) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, "stance", total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref, _ = oracle.run_batch(dogbot.generate("stance", total), threads=2)
    for k in ("x", "y", "z", "s", "flag", "iters"):
        assert np.array_equal(got[k], ref[k]), k


def test_shard_bounds_and_mixed_split():
    from apf_quadruped_b200 import shard
    assert shard.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard.shard_bounds(8, 8) == [(i, i + 1) for i in range(8)]
    assert shard.shard_bounds(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    for world in (1, 2, 4, 8):
        seen = {k: [] for k in ("stance", "crawl", "trot")}
        for r in range(world):
            for k, (lo, hi) in shard.mixed_shards({"stance": 349526, "crawl": 349525, "trot": 349525}, world, r).items():
                seen[k].append((lo, hi))
        for k, v in seen.items():  # every kind is split across ALL ranks, contiguously, without gaps
            assert v[0][0] == 0 and all(a[1] == b[0] for a, b in zip(v, v[1:])) and max(h - l for l, h in v) - min(h - l for l, h in v) <= 1


def test_pack_roundtrip_is_exact():
    import torch
    from apf_quadruped_b200 import shard
    rng = np.random.default_rng(1)
    out = {"x": rng.normal(size=(5, 3)), "y": np.zeros((5, 0)), "z": rng.normal(size=(5, 4)), "s": rng.normal(size=(5, 4)),
           "flag": np.array([0, 2, 0, 3, 1]), "iters": np.array([7, 100, 12, 0, 3])}
    back = shard.unpack(shard.pack(out, torch), {"x": 3, "y": 0, "z": 4, "s": 4})
    for k in out:
        assert np.array_equal(back[k], out[k])
