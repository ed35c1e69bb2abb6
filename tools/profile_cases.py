#!/usr/bin/env python
"""One fixed-size workload per kernel family, for the ncu captures behind profiles/r02_* (tools/capture_profiles.sh
runs each case plain, then under `ncu --set full`; tools/ncu_traffic.py digests the reports).

  python tools/profile_cases.py <case>
cases: shoot (queue kernel, 524288 config-5 points), conte (channel flow), spatial (orrspace lines, lane-pair kernel),
       dense512 / dense256 / dense128 / dense64 (Chebyshev operator, shifts given), band512 / band512_small / band64
       (FD banded, 32768 / 4096 alpha), neutral (orrncbl Newton search), blasius (16384 profiles)
Prints one line: case, units per launch, seconds."""
import math
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import os_stab_b200 as osb  # noqa: E402

S2 = math.sqrt(2.0)
UNITS = {"shoot": 524288, "conte": 262144, "spatial": 4096 * 50, "dense512": 296, "dense256": 1184, "dense128": 4096,
         "dense64": 4096, "band512": 32768, "band512_small": 4096, "band64": 32768, "neutral": 4096, "blasius": 16384}


def chan_case(eng, disc, N, npts, reps=2):
    # launches of the eigenvalue kernel in this case: [0] shift scan of the 9 coarse points, [1] their final solve,
    # [2], [3] the full batch (the one to capture: ncu -s 2); the cross-check against the spectrum would add launches
    import os

    os.environ["OSB_CHAN_NO_VERIFY"] = "1"
    ch = eng.chan(disc, N)
    coarse = np.linspace(0.76, 1.16, 9)
    rc = ch.solve(coarse, 1.0e4)
    good = rc["omega"].real > 0
    alpha = np.linspace(0.76, 1.16, npts)
    sig = np.interp(alpha, coarse[good], rc["omega"].real[good]) + 1j * np.interp(alpha, coarse[good], rc["omega"].imag[good])
    t0 = time.perf_counter()
    for _ in range(reps):
        r = ch.solve(alpha, 1.0e4, sigma0=sig)
    dt = (time.perf_counter() - t0) / reps
    ch.free()
    return dt, int((r["status"] == 0).sum())


def main():
    case = sys.argv[1]
    eng = osb.Engine(0)
    if case == "shoot":
        sys.argv = ["bench.py"]
        import bench

        idx, nre = bench.workload_indices(1, 0, UNITS[case])
        alpha, Re, guess = bench.build_inputs(idx, nre)
        prof = eng.solve_bl()
        opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
        t0 = time.perf_counter()
        for _ in range(2):
            r = eng.shoot_temporal(opts, prof, alpha, Re, guess)
        dt, conv = (time.perf_counter() - t0) / 2, int((r["status"] == 0).sum())
    elif case == "conte":
        prof = eng.channel_profile()
        opts = osb.shoot_defaults(osb.FLOW_CONTE)
        na = nr = 512
        A, R = np.meshgrid(0.95 + 0.1 * np.arange(na) / (na - 1), 9000.0 + 2000.0 * np.arange(nr) / (nr - 1))
        g = np.full(A.size, complex(0.23753, 0.00374)) + 0.2 * (A.ravel() - 1.0) * (-0.3 + 0.05j)
        t0 = time.perf_counter()
        for _ in range(2):
            r = eng.shoot_temporal(opts, prof, A.ravel(), R.ravel(), g)
        dt, conv = (time.perf_counter() - t0) / 2, int((r["status"] == 0).sum())
    elif case == "spatial":
        fact = math.sqrt(2.0) / 1.7207876
        ymax = 20.0 * fact
        prof = eng.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, 1000, 0.0, ymax, 0.5, 0.0)
        opts = osb.shoot_defaults(osb.FLOW_ORRSPACE)
        opts.nstep, opts.ymin, opts.ymax, opts.testalpha = 1000, 0.0, ymax, 0.9
        nl = 4096
        Re3 = (1000.0 + 0.01 * np.arange(nl)) * fact
        t0 = time.perf_counter()
        r = eng.shoot_spatial_lines(opts, prof, Re3, np.full(nl, complex(0.042, 0.0)) * fact, np.full(nl, 0.002) * fact,
                                    np.full(nl, complex(0.13809592, 0.0051958399)) * fact, 50)
        dt, conv = time.perf_counter() - t0, int((r["status"] == 0).sum())
    elif case.startswith("dense"):
        dt, conv = chan_case(eng, osb.DISC_CHEB, int(case[5:]), UNITS[case])
    elif case.startswith("band"):
        dt, conv = chan_case(eng, osb.DISC_FD, 64 if case == "band64" else 512, UNITS[case])
    elif case == "neutral":
        ch = eng.chan(osb.DISC_CHEB_NUM, 64)
        t0 = time.perf_counter()
        r = ch.neutral(6000.0 * (1.0 + np.arange(4096) / 1024.0), 1.0, sfac=0.01)
        dt, conv = time.perf_counter() - t0, int((r["status"] == 0).sum())
    elif case == "blasius":
        beta = np.linspace(-0.05, 0.3, UNITS[case])
        t0 = time.perf_counter()
        o = eng.solve_bl_batch(0.47 + 0.9 * beta, beta, osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 12.0)
        dt, conv = time.perf_counter() - t0, int(np.isfinite(o["U"]).all(axis=1).sum())
    else:
        raise SystemExit(f"unknown case {case}")
    print(f"{case}: units {UNITS[case]} converged {conv} seconds {dt:.4f}")
    eng.close()


if __name__ == "__main__":
    main()
