#!/usr/bin/env python
"""Lane-pair kernel vs one-thread-per-point kernel on sweeps too small to fill the GPU:
orrspace continuation lines (config 3) and small contebl batches.  One JSON line each."""
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


def timed(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def ab(label, units, fn):
    res = {}
    for mode in ("pair", "plain"):
        if mode == "plain":
            os.environ["OSB_NO_PAIR"] = "1"
        else:
            os.environ.pop("OSB_NO_PAIR", None)
        dt, out = timed(fn)
        res[mode] = dt
    os.environ.pop("OSB_NO_PAIR", None)
    print(json.dumps({"case": label, "units": units, "pair_s": res["pair"], "plain_s": res["plain"],
                      "pair_units_per_s": units / res["pair"], "plain_units_per_s": units / res["plain"],
                      "speedup": res["plain"] / res["pair"]}), flush=True)


fact = S2 / 1.7207876
ymin, ymax = 0.0, 20.0 * fact
prof = eng.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
opts = osb.shoot_defaults(osb.FLOW_ORRSPACE)
opts.nstep, opts.ymin, opts.ymax, opts.testalpha = 1000, ymin, ymax, 0.9
for nl in (1, 64, 1024, 4096, 16384):
    Re3 = (1000.0 + 0.01 * np.arange(nl)) * fact
    om = np.full(nl, complex(0.042, 0.0)) * fact
    dom = np.full(nl, 0.002) * fact
    a0 = np.full(nl, complex(0.13809592, 0.0051958399)) * fact
    ab(f"orrspace {nl} lines x 50 omega", nl * 50,
       lambda: eng.shoot_spatial_lines(opts, prof, Re3, om, dom, a0, 50))
prof.free()

bl = eng.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
for n in (1, 256, 4096, 16384):
    al = (0.179 + 0.00001 * np.arange(n)) * S2
    ab(f"contebl {n} points", n,
       lambda: eng.shoot_temporal(opts, bl, al, np.full(n, 580.0 * S2), np.full(n, complex(0.36412287, 0.0079597206))))
bl.free()
eng.close()
