#!/usr/bin/env python
"""Phase split of the banded FD kernel (run with OSB_CHAN_PROFILE=1): 41 alpha at N = 128 and 512,
scan + final solve, then a solve with the converged shifts given."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
import os_stab_b200 as osb
eng = osb.Engine(0)
for N in (128, 512):
    ch = eng.chan(osb.DISC_FD, N)
    alpha = np.linspace(0.76, 1.16, 41)
    r = ch.solve(alpha, 1.0e4)
    print("N", N, "---- given shifts", flush=True)
    sys.stderr.flush()
    r2 = ch.solve(alpha, 1.0e4, sigma0=r["omega"])
    print("iters", r2["iters"].mean(), flush=True)
    ch.free()
eng.close()
