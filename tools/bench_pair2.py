#!/usr/bin/env python
"""Lane-pair kernel, previous form (OSB_PAIR_V1=1: loop nest, lock-step exchange + test after every step) against the
current one (one pass per trip of a single loop, step k+1 started while step k is tested), at several lines-per-warp
settings: config 3 (sweep.inp as 4096 continuation lines x 50 omega) and small contebl batches.  Every variant must
return the same bits.  One JSON line per case."""
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
KEYS = ("OSB_PAIR_V1", "OSB_PAIR_LPW", "OSB_NO_PAIR")


def timed(fn, reps=2):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def variants(label, units, fn, fields, lpws):
    ref = None
    row = {"case": label, "units": units}
    for name, env in [("v1", {"OSB_PAIR_V1": "1"})] + [(f"lpw{l}", {"OSB_PAIR_V1": "0", "OSB_PAIR_LPW": str(l)}) for l in lpws] + [("auto", {})]:
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        dt, out = timed(fn)
        if ref is None:
            ref = out
        same = all(np.array_equal(np.asarray(out[f]), np.asarray(ref[f])) for f in fields)
        row[name] = {"s": round(dt, 5), "units_per_s": round(units / dt, 1), "bits_equal_v1": bool(same)}
    for k in KEYS:
        os.environ.pop(k, None)
    row["speedup_auto_vs_v1"] = round(row["v1"]["s"] / row["auto"]["s"], 3)
    print(json.dumps(row), flush=True)
    return ref


fact = S2 / 1.7207876
ymin, ymax = 0.0, 20.0 * fact
prof = eng.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
opts = osb.shoot_defaults(osb.FLOW_ORRSPACE)
opts.nstep, opts.ymin, opts.ymax, opts.testalpha = 1000, ymin, ymax, 0.9
nl = 4096
Re3 = (600.0 + 0.5 * np.arange(nl)) * fact
om = np.full(nl, complex(0.042, 0.0)) * fact
dom = np.full(nl, 0.002) * fact
a_seed = complex(0.13809592, 0.0051958399) * fact
i1000 = int(round((1000.0 - 600.0) / 0.5))
a0 = np.zeros(nl, dtype=complex)
for rng in (range(i1000, nl), range(i1000 - 1, -1, -1)):
    cur = a_seed if rng.start == i1000 else a0[i1000]
    for lo in range(rng.start, rng.stop, rng.step * 64):
        idx = list(range(lo, max(min(lo + rng.step * 64, nl), -1), rng.step))
        res = eng.shoot_spatial_lines(opts, prof, Re3[idx], om[idx], dom[idx], np.full(len(idx), cur), 1)
        a0[idx] = res["alpha"][:, 0]
        cur = res["alpha"][-1, 0]
FIELDS3 = ("alpha", "passes", "qend", "status")
r = variants("config 3: 4096 lines x 50 omega", nl * 50,
             lambda: eng.shoot_spatial_lines(opts, prof, Re3, om, dom, a0, 50), FIELDS3, (16, 8, 4, 2))
p = np.asarray(r["passes"]).reshape(nl, 50).astype(np.int64)
model = {"case": "pass-count model of config 3", "mean_total_per_line": float(p.sum(1).mean())}
for L in (16, 8, 4, 2, 1):
    g = p.reshape(nl // L, L, 50)
    model[f"L{L}"] = {"nest_sum_of_max": int(g.max(1).sum(1).max()), "flat_max_of_sum": int(g.sum(2).max())}
print(json.dumps(model), flush=True)
for n3 in (1, 64, 1024):
    idx = np.arange(n3) + i1000
    variants(f"orrspace {n3} lines x 50 omega", n3 * 50,
             lambda: eng.shoot_spatial_lines(opts, prof, Re3[idx], om[idx], dom[idx], a0[idx], 50), FIELDS3, (16, 4, 1))
prof.free()

bl = eng.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
for n in (1, 256, 4096, 9472):
    al = (0.179 + 0.00001 * np.arange(n)) * S2
    variants(f"contebl {n} points", n,
             lambda: eng.shoot_temporal(opts, bl, al, np.full(n, 580.0 * S2), np.full(n, complex(0.36412287, 0.0079597206))),
             ("c", "passes", "qend", "status"), (16, 4, 1))
bl.free()
eng.close()
