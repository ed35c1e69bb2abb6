import json, math, os, sys, time
import numpy as np
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os_stab_b200 as osb
eng = osb.Engine(0); S2 = math.sqrt(2.0)
fact = S2 / 1.7207876
prof = eng.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, 1000, 0.0, 20.0 * fact, 0.5, 0.0)
opts = osb.shoot_defaults(osb.FLOW_ORRSPACE)
opts.nstep, opts.ymin, opts.ymax, opts.testalpha = 1000, 0.0, 20.0 * fact, 0.9
for nl in (1, 4096):
    Re3 = (1000.0 + 0.01 * np.arange(nl)) * fact
    f = lambda: eng.shoot_spatial_lines(opts, prof, Re3, np.full(nl, complex(0.042, 0.0)) * fact, np.full(nl, 0.002) * fact, np.full(nl, complex(0.13809592, 0.0051958399)) * fact, 50)
    f(); t0 = time.perf_counter(); r = f(); dt = time.perf_counter() - t0
    print(json.dumps({"lines": nl, "s": dt, "points_per_s": nl * 50 / dt, "alpha_sum": repr(complex(r["alpha"].sum())), "passes": int(r["passes"].sum()), "qend": int(r["qend"].sum())}))
