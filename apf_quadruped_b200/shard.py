"""Multi-GPU plumbing of the batched solve: contiguous shards in, one gather out.

QPs of a batch are independent (the reference solves one per control tick, qpSWIFT.c:473-644), so
the path shards with NO data-path collective: rank r of W owns a contiguous range of instance
indices, builds the same symbolic plan (integers, value independent) and runs ``QP_SOLVE_BATCH``
on its own GPU.  The only collective is the one BASELINE.json's north_star names: gathering the
solutions, exit flags and iteration counts (``torch.distributed`` all-gather: NCCL over NVLink on
GPUs, gloo on CPU for the tests).  Per-QP arithmetic does not depend on the shard boundaries, so
results are bit-identical for 1, 2, 4 or 8 ranks.

Mixed contact-set batches (BASELINE.json configs[2]: stance / crawl / trot, one plan per pattern)
are split PER PATTERN across all ranks, so that every GPU sees the same mix and the shards stay
balanced although a stance factorisation costs more than twice a trot one.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(total: int, world: int) -> list[tuple[int, int]]:
    """Contiguous ranges [lo, hi) of `total` instances for `world` ranks; the first total % world
    ranks hold one more."""
    base, extra = divmod(int(total), int(world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def mixed_shards(counts: dict, world: int, rank: int) -> dict:
    """For a mixed batch {pattern kind: instances}, the range of each kind this rank solves."""
    return {kind: shard_bounds(cnt, world)[rank] for kind, cnt in counts.items()}


FIELDS = ("x", "y", "z", "s")


def pack(out: dict, torch, device=None):
    """[count, n+p+2m+2] float64: x | y | z | s | flag | iterations (flags and counts are small
    integers, exact in binary64), one row per QP -- one collective instead of six."""
    cols = [torch.as_tensor(out[k]).reshape(len(out["flag"]), -1).to(torch.float64) for k in FIELDS]
    cols += [torch.as_tensor(out["flag"]).to(torch.float64)[:, None], torch.as_tensor(out["iters"]).to(torch.float64)[:, None]]
    t = torch.cat(cols, dim=1).contiguous()
    return t.to(device) if device is not None else t


def unpack(t, widths: dict) -> dict:
    """Inverse of pack(); widths = {"x": n, "y": p, "z": m, "s": m}."""
    t = t.cpu().numpy()
    out, at = {}, 0
    for k in FIELDS:
        out[k] = np.ascontiguousarray(t[:, at:at + widths[k]])
        at += widths[k]
    out["flag"] = t[:, at].astype(np.int64)
    out["iters"] = t[:, at + 1].astype(np.int64)
    return out


def gather(local: dict, total: int, widths: dict, dist, torch, device=None, group=None) -> dict:
    """All-gather the shards' results into instance order.  Shards may differ by one row: every rank
    pads to the largest shard (the collective needs equal sizes) and the padding is dropped."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(total, world)
    most = max(hi - lo for lo, hi in bounds)
    mine = pack(local, torch, device)
    if mine.shape[0] < most:
        mine = torch.cat([mine, torch.zeros((most - mine.shape[0], mine.shape[1]), dtype=mine.dtype, device=mine.device)])
    everything = torch.empty((world * most, mine.shape[1]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(everything, mine, group=group)
    rows = torch.cat([everything[r * most:r * most + (hi - lo)] for r, (lo, hi) in enumerate(bounds)])
    assert rows.shape[0] == total and bounds[rank][1] - bounds[rank][0] == len(local["flag"])
    return unpack(rows, widths)
