#!/usr/bin/env python
"""Times the channel matrix path (K3/K4) -- BASELINE config 4's throughput variant:
`npts` alpha values uniformly in [0.76, 1.16] at Re = 1e4 for N = 64..512, shifts
given (continuation values), so one LU + a few inverse iterations per point.
Prints one JSON object per N.  Not the headline metric (bench.py is)."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os_stab_b200 as osb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npts", type=int, default=1024)
    ap.add_argument("--sizes", default="64,128,256,512")
    ap.add_argument("--disc", type=int, default=osb.DISC_FD)
    args = ap.parse_args()
    eng = osb.Engine(0)
    for N in [int(x) for x in args.sizes.split(",")]:
        t0 = time.perf_counter()
        ch = eng.chan(args.disc, N)
        t_tables = time.perf_counter() - t0
        coarse = np.linspace(0.76, 1.16, 9)
        t0 = time.perf_counter()
        rc = ch.solve(coarse, 1.0e4)  # scan: no guesses
        t_scan = time.perf_counter() - t0
        alpha = np.linspace(0.76, 1.16, args.npts)
        # shifts: interpolate the physical-branch scan results in alpha
        good = rc["omega"].real > 0
        sig = np.interp(alpha, coarse[good], rc["omega"].real[good]) + 1j * np.interp(alpha, coarse[good], rc["omega"].imag[good])
        ch.solve(alpha, 1.0e4, sigma0=sig)  # warm-up at full size (the stream-ordered pool grows on first use)
        dt = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            r = ch.solve(alpha, 1.0e4, sigma0=sig)
            dt = min(dt, time.perf_counter() - t0)
        n = N - 1
        # at least one factorisation per point (a re-shift adds one; the kernel does not report them)
        nfac = float(args.npts)
        lu_flops = (8.0 / 3.0) * n ** 3 * nfac
        print(json.dumps({
            "N": N, "npts": args.npts, "seconds": dt, "eig_per_s": args.npts / dt,
            "converged": int((r["status"] == 0).sum()), "mean_inverse_iterations": float(r["iters"].mean()),
            "dense_lu_equiv_tflops_at_one_lu_per_point": lu_flops / dt / 1e12,
            "path": "banded LU (FD, half bandwidth 4)" if (args.disc == osb.DISC_FD and not __import__("os").environ.get("OSB_FD_DENSE")) else "dense blocked LU + DMMA",
            "tables_s": t_tables, "scan_9pts_s": t_scan}), flush=True)
        ch.free()
    eng.close()


if __name__ == "__main__":
    main()
