#!/usr/bin/env python
"""Banded FD path: warp-per-instance vs thread-per-instance kernel, shifts given."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os_stab_b200 as osb  # noqa: E402

eng = osb.Engine(0)
for N in (64, 128, 256, 512):
    ch = eng.chan(osb.DISC_FD, N)
    coarse = np.linspace(0.76, 1.16, 9)
    rc = ch.solve(coarse, 1.0e4)
    good = rc["omega"].real > 0
    for npts in (41, 4096, 32768):
        alpha = np.linspace(0.76, 1.16, npts)
        sig = np.interp(alpha, coarse[good], rc["omega"].real[good]) + 1j * np.interp(alpha, coarse[good], rc["omega"].imag[good])
        out = {"N": N, "npts": npts}
        for mode in ("warp", "thread"):
            if mode == "thread":
                os.environ["OSB_FD_THREAD"] = "1"
            else:
                os.environ.pop("OSB_FD_THREAD", None)
            ch.solve(alpha, 1.0e4, sigma0=sig)
            best = 1e30
            for _ in range(3):
                t0 = time.perf_counter()
                r = ch.solve(alpha, 1.0e4, sigma0=sig)
                best = min(best, time.perf_counter() - t0)
            out[mode + "_eig_per_s"] = npts / best
            out[mode + "_iters"] = float(r["iters"].mean())
            out[mode + "_conv"] = int((r["status"] == 0).sum())
        os.environ.pop("OSB_FD_THREAD", None)
        out["speedup"] = out["warp_eig_per_s"] / out["thread_eig_per_s"]
        print(json.dumps(out), flush=True)
    ch.free()
eng.close()
