"""Sharding a sweep over ranks (one process per GPU) and the single gather at the end.

Points are independent: the flattened point index is dealt to ranks in
cyclic tiles (pass counts vary smoothly over the (alpha, Re) plane, so contiguous
blocks would imbalance), every rank solves its own tiles with no communication,
and the per-point results {c, passes, qend, status} are gathered to rank 0 once.
The solve itself is injected (``solve(alpha, Re, guess) -> dict``): the product
passes ``Engine.shoot_temporal``; the CPU tests pass a stand-in so that the
partition / gather / re-assembly logic runs under gloo without a GPU.
"""
from __future__ import annotations

import numpy as np


def shard_indices(npts: int, world: int, rank: int, tile: int = 4096) -> np.ndarray:
    ntile = (npts + tile - 1) // tile
    mine = np.arange(rank, ntile, world, dtype=np.int64)
    idx = (mine[:, None] * tile + np.arange(tile, dtype=np.int64)[None, :]).reshape(-1)
    return idx[idx < npts]


def sharded_sweep(solve, alpha, Re, guess, rank=0, world=1, tile=4096, dist=None, device="cpu"):
    """Solve rank's tiles and gather to rank 0.  Returns the full-order result dict on
    rank 0 and None elsewhere.  ``dist`` is ``torch.distributed`` (already initialised)
    when world > 1."""
    alpha = np.asarray(alpha, dtype=np.complex128)
    Re = np.asarray(Re, dtype=np.float64)
    guess = np.asarray(guess, dtype=np.complex128)
    npts = Re.size
    idx = shard_indices(npts, world, rank, tile)
    local = solve(alpha[idx], Re[idx], guess[idx])
    if world == 1:
        out = {k: np.empty(npts, dtype=np.asarray(v).dtype) for k, v in local.items() if k in ("c", "passes", "qend", "status")}
        for k in out:
            out[k][idx] = local[k]
        return out
    import torch

    # fixed-size payload per rank: pad to the largest shard
    nmax = max(shard_indices(npts, world, r, tile).size for r in range(world))
    f = torch.zeros(nmax, 2, dtype=torch.float64)
    i = torch.zeros(nmax, 3, dtype=torch.int32)
    f[: idx.size] = torch.from_numpy(np.ascontiguousarray(local["c"]).view(np.float64).reshape(-1, 2))
    i[: idx.size, 0] = torch.from_numpy(np.ascontiguousarray(local["passes"], dtype=np.int32))
    i[: idx.size, 1] = torch.from_numpy(np.ascontiguousarray(local["qend"], dtype=np.int32))
    i[: idx.size, 2] = torch.from_numpy(np.ascontiguousarray(local["status"], dtype=np.int32))
    f, i = f.to(device), i.to(device)
    gf = [torch.empty_like(f) for _ in range(world)] if rank == 0 else None
    gi = [torch.empty_like(i) for _ in range(world)] if rank == 0 else None
    dist.gather(f, gf, dst=0)
    dist.gather(i, gi, dst=0)
    if rank != 0:
        return None
    out = dict(c=np.empty(npts, dtype=np.complex128), passes=np.empty(npts, dtype=np.int32),
               qend=np.empty(npts, dtype=np.int32), status=np.empty(npts, dtype=np.int32))
    for r in range(world):
        ridx = shard_indices(npts, world, r, tile)
        fr = gf[r].cpu().numpy()[: ridx.size]
        ir = gi[r].cpu().numpy()[: ridx.size]
        out["c"][ridx] = fr[:, 0] + 1j * fr[:, 1]
        out["passes"][ridx] = ir[:, 0]
        out["qend"][ridx] = ir[:, 1]
        out["status"][ridx] = ir[:, 2]
    return out
