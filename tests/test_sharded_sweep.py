"""The N>1 path on CPU: world_size-2 gloo, sharding by cyclic tiles + one gather.
The per-point solve is the CPU oracle on a small batch (test stand-in for the GPU
engine); the assertion is that the sharded, gathered result equals the 1-rank one
bit for bit (points are independent)."""
import math
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
S2 = math.sqrt(2.0)


def _inputs(n):
    alpha = (0.179 + 0.00002 * np.arange(n)) * S2 + 0j
    Re = np.full(n, 580.0 * S2) * (1 + 1e-4 * (np.arange(n) % 7))
    guess = np.full(n, complex(0.36412287, 0.0079597206))
    return alpha, Re, guess


def _solve_factory():
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_binding as ob

    o = ob.load(rebuild=False)
    prof, _, _ = o.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, 200, 0.0, 17.0, 0.5, 0.0)

    def solve(alpha, Re, guess):
        return o.shoot_temporal_batch(ob.FLOW_CONTEBL, ob.INT_CK45, 200, 10.0, 0.0, 17.0, prof, alpha, Re, guess)

    return solve


def _worker(rank, world, port, n, tile, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist

    import os_stab_b200.sweep as sweep

    dist.init_process_group("gloo", rank=rank, world_size=world)
    alpha, Re, guess = _inputs(n)
    out = sweep.sharded_sweep(_solve_factory(), alpha, Re, guess, rank, world, tile, dist)
    if rank == 0:
        q.put({k: v.tolist() if k != "c" else [(z.real, z.imag) for z in v] for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_shard_indices_cover_every_point_once(osb):
    import os_stab_b200.sweep as sweep

    for npts, world, tile in ((37, 2, 8), (4096 * 5 + 3, 4, 4096), (5, 8, 4096)):
        parts = [sweep.shard_indices(npts, world, r, tile) for r in range(world)]
        assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(npts))
        sizes = [p.size for p in parts]
        assert max(sizes) - min(sizes) <= tile


@pytest.mark.timeout(300)
def test_world2_gloo_equals_single_rank(oracle, osb):
    import torch.multiprocessing as mp

    import os_stab_b200.sweep as sweep

    n, tile = 37, 8  # ragged: last tile is partial, ranks get 21 and 16 points
    alpha, Re, guess = _inputs(n)
    single = sweep.sharded_sweep(_solve_factory(), alpha, Re, guess, 0, 1, tile)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, tile, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = np.array([complex(a, b) for a, b in got["c"]])
    assert np.array_equal(c, single["c"])
    for k in ("passes", "qend", "status"):
        assert np.array_equal(np.array(got[k]), single[k])
    assert (single["status"] == 0).all()
