#!/usr/bin/env python
"""Which shooting kernel for 9472..65536 single solves: one thread per point, the lane-pair forms (OSB_PAIR_MAX lifts
their size limit), the queue kernel (OSB_QUEUE_MIN=512).  One JSON line per size; every form must return the same bits.
Measured on B200: one thread per point wins from 16384 points on (1.11 M eig/s against 0.86 M / 0.81 M for the pair forms
at 16384; 1.25 M against 1.07 M / 0.98 M at 32768), the pair forms at 9472 (0.74 M against 0.66 M): the thresholds in
launch_shoot stand."""
import json, math, os, sys, time
import numpy as np
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os_stab_b200 as osb
eng = osb.Engine(0); S2 = math.sqrt(2.0)
bl = eng.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
for n in (9472, 16384, 24576, 32768, 49152, 65536):
    al = (0.179 + 0.000002 * np.arange(n)) * S2
    Re = np.full(n, 580.0 * S2); g = np.full(n, complex(0.36412287, 0.0079597206))
    row = {"n": n}
    ref = None
    for name, env in (("plain", {"OSB_NO_PAIR": "1"}), ("pair_v1", {"OSB_PAIR_MAX": "1000000", "OSB_PAIR_V1": "1"}),
                      ("pair_new16", {"OSB_PAIR_MAX": "1000000", "OSB_PAIR_V1": "0", "OSB_PAIR_LPW": "16"}),
                      ("queue", {"OSB_QUEUE_MIN": "512"})):
        for k in ("OSB_NO_PAIR", "OSB_PAIR_MAX", "OSB_PAIR_V1", "OSB_PAIR_LPW", "OSB_QUEUE_MIN"): os.environ.pop(k, None)
        os.environ.update(env)
        eng.shoot_temporal(opts, bl, al, Re, g)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); r = eng.shoot_temporal(opts, bl, al, Re, g); best = min(best, time.perf_counter() - t0)
        if ref is None: ref = r["c"].copy()
        row[name] = round(n / best); row[name + "_same"] = bool(np.array_equal(ref, r["c"]))
    print(json.dumps(row), flush=True)
