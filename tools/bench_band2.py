#!/usr/bin/env python
"""Nine-lane banded FD kernel at deck-sized and GPU-filling sweeps, shifts given; one JSON line per (N, sweep size).
OSB_LIB=<another build of libosb.so> python tools/bench_band2.py  gives the A/B partner."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os_stab_b200 as osb  # noqa: E402

eng = osb.Engine(0)
for N in (64, 128, 256, 512, 1024):
    ch = eng.chan(osb.DISC_FD, N)
    coarse = np.linspace(0.76, 1.16, 9)
    rc = ch.solve(coarse, 1.0e4)
    good = rc["omega"].real > 0
    for npts in (41, 1024, 2048, 4096, 8192, 32768):
        alpha = np.linspace(0.76, 1.16, npts)
        sig = np.interp(alpha, coarse[good], rc["omega"].real[good]) + 1j * np.interp(alpha, coarse[good], rc["omega"].imag[good])
        ch.solve(alpha, 1.0e4, sigma0=sig)
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            r = ch.solve(alpha, 1.0e4, sigma0=sig)
            best = min(best, time.perf_counter() - t0)
        n = N - 1
        gbs = npts * (13.0 * n * 16 * (1.0 + float(r["iters"].mean()))) / best / 1e9
        print(json.dumps({"N": N, "npts": npts, "eig_per_s": round(npts / best, 1), "iters": round(float(r["iters"].mean()), 3),
                          "conv": int((r["status"] == 0).sum()), "factor_GBps": round(gbs, 1),
                          "omega_checksum": repr(complex(r["omega"].sum()))}), flush=True)
    ch.free()
eng.close()
