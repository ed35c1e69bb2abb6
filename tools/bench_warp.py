#!/usr/bin/env python
"""Warp-per-instance matrix kernels (n <= 64): Chebyshev operator at N = 24..64, 4096 alpha, shifts given, and orrncbl's
Newton search (N = 64, 4096 Re).  One JSON line each, with a checksum of the results (A/B: OSB_LIB=<other build>)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os_stab_b200 as osb  # noqa: E402

eng = osb.Engine(0)
npts = 4096
for N in (24, 32, 48, 64):
    ch = eng.chan(osb.DISC_CHEB, N)
    coarse = np.linspace(0.76, 1.16, 9)
    rc = ch.solve(coarse, 1.0e4)
    good = rc["omega"].real > 0
    alpha = np.linspace(0.76, 1.16, npts)
    sig = np.interp(alpha, coarse[good], rc["omega"].real[good]) + 1j * np.interp(alpha, coarse[good], rc["omega"].imag[good])
    for adj in (None, osb.ADJ_PENCIL):
        ch.solve(alpha, 1.0e4, sigma0=sig, want_adj=adj)
        dt = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            r = ch.solve(alpha, 1.0e4, sigma0=sig, want_adj=adj)
            dt = min(dt, time.perf_counter() - t0)
        print(json.dumps({"N": N, "npts": npts, "adjoint": adj is not None, "eig_per_s": round(npts / dt, 1),
                          "iters": float(r["iters"].mean()), "conv": int((r["status"] == 0).sum()),
                          "omega_sum": repr(complex(r["omega"].sum())),
                          "avec_sum": repr(complex(r["avec"].sum())) if adj is not None else None}), flush=True)
    ch.free()
ch = eng.chan(osb.DISC_CHEB_NUM, 64)
Re5 = 6000.0 * (1.0 + np.arange(4096) / 1024.0)
ch.neutral(Re5, 1.0, sfac=0.01)
t0 = time.perf_counter()
r = ch.neutral(Re5, 1.0, sfac=0.01)
dt = time.perf_counter() - t0
print(json.dumps({"case": "orrncbl Newton search N=64, 4096 Re", "neutral_points_per_s": round(int((r["status"] == 0).sum()) / dt, 1),
                  "steps": float(r["steps"].mean()), "alpha_sum": repr(float(r["alpha"].sum()))}), flush=True)
ch.free()
eng.close()
