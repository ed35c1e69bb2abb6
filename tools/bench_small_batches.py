#!/usr/bin/env python
"""contebl batches too small to fill the GPU: which shooting kernel wins?  queue (default for >= 512 points),
lane pair (OSB_NO_QUEUE=1, <= 9472 points), one thread per point (OSB_NO_QUEUE=1 OSB_NO_PAIR=1)."""
import json
import math
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os_stab_b200 as osb  # noqa: E402

eng = osb.Engine(0)
S2 = math.sqrt(2.0)
bl = eng.solve_bl()
opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
modes = {"queue": {}, "pair": {"OSB_NO_QUEUE": "1"}, "thread": {"OSB_NO_QUEUE": "1", "OSB_NO_PAIR": "1"}}
import bench  # noqa: E402  (config-5 inputs: pass counts vary from point to point)

for n in (512, 2048, 8192, 16384, 32768, 65536, 131072, 262144, 524288):
    idx = np.arange(n, dtype=np.int64) * (4096 * 4096 // n)
    al, Re, g = bench.build_inputs(idx, 4096)
    out = {"points": n}
    for name, env in modes.items():
        for k in ("OSB_NO_QUEUE", "OSB_NO_PAIR"):
            os.environ.pop(k, None)
        os.environ.update(env)
        eng.shoot_temporal(opts, bl, al, Re, g)
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            r = eng.shoot_temporal(opts, bl, al, Re, g)
            best = min(best, time.perf_counter() - t0)
        out[name + "_eig_per_s"] = round(n / best)
    print(json.dumps(out), flush=True)
