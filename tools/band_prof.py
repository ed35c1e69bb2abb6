#!/usr/bin/env python
"""Banded FD kernels under a profiler or with OSB_CHAN_PROFILE=1 (in-kernel phase split):
N = 512, `npts` alpha (default 41, a deck-sized sweep): scan + final solve, then two solves with the
converged shifts given (the last launches are the ones to capture)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os_stab_b200 as osb  # noqa: E402

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 41
N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
eng = osb.Engine(0)
ch = eng.chan(osb.DISC_FD, N)
coarse = np.linspace(0.76, 1.16, 9)
rc = ch.solve(coarse, 1.0e4)
good = rc["omega"].real > 0
alpha = np.linspace(0.76, 1.16, npts)
sig = np.interp(alpha, coarse[good], rc["omega"].real[good]) + 1j * np.interp(alpha, coarse[good], rc["omega"].imag[good])
for _ in range(2):
    r = ch.solve(alpha, 1.0e4, sigma0=sig)
print("N", N, "npts", npts, "iters", r["iters"].mean(), "converged", int((r["status"] == 0).sum()), flush=True)
ch.free()
eng.close()
