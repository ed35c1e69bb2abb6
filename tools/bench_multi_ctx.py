#!/usr/bin/env python
"""One host process driving all visible GPUs through osb_create_multi (the deployment of a single-process
Fortran host): e2e throughput of osb_shoot_temporal on a config-5 slab, against one device."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import os_stab_b200 as osb  # noqa: E402

ng = torch.cuda.device_count()
npts = 4096 * 1024 * ng
# npts points spread evenly over the 4096 x 4096 grid of config 5
idx = np.arange(npts, dtype=np.int64)
alpha, Re, guess = bench.build_inputs(idx * (4096 * 4096 // npts), 4096)
opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
for devs in ([0], list(range(ng))):
    eng = osb.Engine(devices=devs) if len(devs) > 1 else osb.Engine(0)
    prof = eng.solve_bl()
    n = npts if len(devs) > 1 else npts // ng
    eng.shoot_temporal(opts, prof, alpha[:n], Re[:n], guess[:n])
    t0 = time.perf_counter()
    r = eng.shoot_temporal(opts, prof, alpha[:n], Re[:n], guess[:n])
    dt = time.perf_counter() - t0
    print(json.dumps({"devices": len(devs), "points": n, "seconds": dt, "eig_per_s": float((r["status"] == 0).sum()) / dt,
                      "passes_per_point": float(r["passes"].mean())}), flush=True)
    prof.free()
    eng.close()
