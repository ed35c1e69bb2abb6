#!/usr/bin/env python
"""Throughput of the other BASELINE configs (2, 3, 4, 5b) on one GPU -- supplementary
numbers; the headline metric is bench.py.  One JSON line per config.

  2  conte channel temporal: 262144 (alpha, Re) points around (1, 1e4), guesses by a coarse pass
  3  orrspace sweep.inp as 4096 independent continuation lines, Re_delta* = 600 + 0.5 i, 50 omega each
  4  orrfdchan / orrcolchan: see tools/bench_chan.py
  5b orrncbl neutral points: 4096 Re values 6000 (1 + i/1024), N = 64, sfac = .01, alpha0 = 1
"""
import json
import math
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os_stab_b200 as osb  # noqa: E402

eng = osb.Engine(0)


def timed(fn, reps=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


# ---- config 2: conte channel, CK45, n = 2000
prof = eng.channel_profile()
opts = osb.shoot_defaults(osb.FLOW_CONTE)
na = nr = 512
alpha = 0.95 + 0.1 * np.arange(na) / (na - 1)
Re = 9000.0 + 2000.0 * np.arange(nr) / (nr - 1)
A, R = np.meshgrid(alpha, Re)
# guesses: solve a 9x9 coarse grid from the reference default, then bilinear interpolation
ca = np.linspace(alpha[0], alpha[-1], 9); cr = np.linspace(Re[0], Re[-1], 9)
CA, CR = np.meshgrid(ca, cr)
coarse = eng.shoot_temporal(opts, prof, CA.ravel(), CR.ravel(), np.full(81, complex(0.23753, 0.00374)))
ok = coarse["status"] == 0
cc = coarse["c"].reshape(9, 9)
x = (A - alpha[0]) / (alpha[-1] - alpha[0]) * 8; y = (R - Re[0]) / (Re[-1] - Re[0]) * 8
x0 = np.minimum(x.astype(int), 7); y0 = np.minimum(y.astype(int), 7); fx = x - x0; fy = y - y0
guess = (1 - fy) * ((1 - fx) * cc[y0, x0] + fx * cc[y0, x0 + 1]) + fy * ((1 - fx) * cc[y0 + 1, x0] + fx * cc[y0 + 1, x0 + 1])
dt, r = timed(lambda: eng.shoot_temporal(opts, prof, A.ravel(), R.ravel(), guess.ravel()))
conv = int((r["status"] == 0).sum())
flops = float(r["passes"].sum()) * 2000 * 990.0  # conte step = 990 flop
print(json.dumps({"config": "2: conte channel temporal grid 512x512 around (alpha=1, Re=1e4), n=2000, CK45",
                  "points": A.size, "coarse_ok": int(ok.sum()), "converged": conv, "seconds": dt, "eig_per_s": conv / dt,
                  "passes_per_point": float(r["passes"].mean()), "algorithmic_tflops": flops / dt / 1e12}), flush=True)
prof.free()

# ---- config 3: orrspace continuation lines
fact = math.sqrt(2.0) / 1.7207876
ymin, ymax = 0.0, 20.0 * fact
prof = eng.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
opts = osb.shoot_defaults(osb.FLOW_ORRSPACE)
opts.nstep, opts.ymin, opts.ymax, opts.testalpha = 1000, ymin, ymax, 0.9
nl = 4096
Re3 = (600.0 + 0.5 * np.arange(nl)) * fact
om = np.full(nl, complex(0.042, 0.0)) * fact
dom = np.full(nl, 0.002) * fact
# first-point guesses: march a single line in Re from the deck's point (Re=1000) outwards
a_seed = complex(0.13809592, 0.0051958399) * fact
i1000 = int(round((1000.0 - 600.0) / 0.5))
a0 = np.zeros(nl, dtype=complex)
cur = a_seed
for rng in (range(i1000, nl), range(i1000 - 1, -1, -1)):
    cur = a_seed if rng.start == i1000 else a0[i1000]
    for lo in range(rng.start, rng.stop, rng.step * 64):
        idx = list(range(lo, max(min(lo + rng.step * 64, nl), -1), rng.step))
        res = eng.shoot_spatial_lines(opts, prof, Re3[idx], om[idx], dom[idx], np.full(len(idx), cur), 1)
        a0[idx] = res["alpha"][:, 0]
        cur = res["alpha"][-1, 0]
dt, r = timed(lambda: eng.shoot_spatial_lines(opts, prof, Re3, om, dom, a0, 50), reps=1)
conv = int((r["status"] == 0).sum())
flops = float(r["passes"].sum()) * 1000 * 802.0
print(json.dumps({"config": "3: orrspace sweep.inp as 4096 continuation lines (Re_delta* = 600 + 0.5 i), 50 omega each, RK4",
                  "points": nl * 50, "converged": conv, "seconds": dt, "eig_per_s": conv / dt,
                  "passes_per_point": float(r["passes"].mean()), "algorithmic_tflops": flops / dt / 1e12}), flush=True)
prof.free()

# ---- config 5b: orrncbl neutral points
ch = eng.chan(osb.DISC_CHEB_NUM, 64)
Re5 = 6000.0 * (1.0 + np.arange(4096) / 1024.0)
dt, r = timed(lambda: ch.neutral(Re5, 1.0, sfac=0.01), reps=1)
conv = int((r["status"] == 0).sum())
print(json.dumps({"config": "5b: orrncbl neutral points, N=64 collocation, 4096 Re in [6000, 30000), sfac=.01, alpha0=1",
                  "points": 4096, "converged": conv, "seconds": dt, "neutral_points_per_s": conv / dt,
                  "newton_steps_mean": float(r["steps"].mean()),
                  "alpha_neutral_at_Re_6000_12000_24000": [float(r["alpha"][0]), float(r["alpha"][1024]), float(r["alpha"][3072])]}), flush=True)
ch.free()
eng.close()
