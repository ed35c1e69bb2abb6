import json, math, os, sys, time
import numpy as np
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os_stab_b200 as osb
eng = osb.Engine(0); S2 = math.sqrt(2.0)
bl = eng.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
for n in (1024, 2048, 4096, 6144, 9472):
    al = (0.179 + 0.00001 * np.arange(n)) * S2
    Re = np.full(n, 580.0 * S2); g = np.full(n, complex(0.36412287, 0.0079597206))
    row = {"n": n}
    for name, env in [("v1", {"OSB_PAIR_V1": "1"}), ("auto", {})] + [(f"lpw{l}", {"OSB_PAIR_V1": "0", "OSB_PAIR_LPW": str(l)}) for l in (2, 4, 6, 8, 10, 12, 16)]:
        for k in ("OSB_PAIR_V1", "OSB_PAIR_LPW"): os.environ.pop(k, None)
        os.environ.update(env)
        eng.shoot_temporal(opts, bl, al, Re, g)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); r = eng.shoot_temporal(opts, bl, al, Re, g); best = min(best, time.perf_counter() - t0)
        row[name] = round(n / best)
    print(json.dumps(row), flush=True)
