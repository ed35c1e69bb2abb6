#!/usr/bin/env python
"""Blasius neutral curve by shooting: solves a (alpha, Re) contebl grid
with the batched shooter and extracts the contour Im c = 0 (both branches) by linear
interpolation in alpha along every Re row.  The reference has no result for this
(orrncbl's boundary-layer branch is dead code), so this is "parity unpinned";
the physical check is the critical Reynolds number Re_delta* = 519.4 (R = 301.8 in
contebl's units).

  python tools/neutral_from_grid.py [--na 512] [--nre 256] [--out neutral.dat] [--refine]

--refine polishes every contour point with osb_shoot_neutral (secant on alpha to |Im c| <= 1e-10).
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (grid definition + seeds)
import os_stab_b200 as osb  # noqa: E402


def neutral_curve(eng, prof, na=512, nre=256):
    alpha_in = 0.05 + 0.25 * np.arange(na) / (na - 1)
    Re_in = 300.0 * 6.0 ** (np.arange(nre) / (nre - 1))
    seeds = np.load(ROOT / "bench_data" / "cfg5_seeds_65x65.npy")
    jj, ii = np.meshgrid(np.arange(na), np.arange(nre))
    x = jj * (64.0 / (na - 1)); y = ii * (64.0 / (nre - 1))
    x0 = np.minimum(x.astype(int), 63); y0 = np.minimum(y.astype(int), 63)
    fx, fy = x - x0, y - y0
    guess = ((1 - fy) * ((1 - fx) * seeds[y0, x0] + fx * seeds[y0, x0 + 1]) + fy * ((1 - fx) * seeds[y0 + 1, x0] + fx * seeds[y0 + 1, x0 + 1]))
    A, R = np.meshgrid(alpha_in, Re_in)
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    r = eng.shoot_temporal(opts, prof, (A * bench.S2).ravel(), (R * bench.S2).ravel(), guess.ravel())
    ci = r["c"].imag.reshape(nre, na)
    ok = (r["status"] == 0).reshape(nre, na)
    rows = []
    for i in range(nre):
        s = np.sign(ci[i])
        for j in np.nonzero((s[:-1] * s[1:] < 0) & ok[i, :-1] & ok[i, 1:])[0]:
            t = ci[i, j] / (ci[i, j] - ci[i, j + 1])
            a0 = alpha_in[j] + t * (alpha_in[j + 1] - alpha_in[j])
            cr = r["c"].real.reshape(nre, na)
            rows.append((Re_in[i], a0, cr[i, j] + t * (cr[i, j + 1] - cr[i, j]), 1 if ci[i, j] < 0 else -1))
    return np.array(rows), ci, Re_in, alpha_in


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--na", type=int, default=512)
    ap.add_argument("--nre", type=int, default=256)
    ap.add_argument("--out", default="neutral.dat")
    ap.add_argument("--refine", action="store_true")
    args = ap.parse_args()
    eng = osb.Engine(0)
    prof = eng.solve_bl()
    rows, ci, Re_in, alpha_in = neutral_curve(eng, prof, args.na, args.nre)
    if args.refine:
        import time

        opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
        t0 = time.perf_counter()
        r = eng.shoot_neutral(opts, prof, rows[:, 0] * bench.S2, rows[:, 1] * bench.S2, rows[:, 2] + 0j, dalpha_rel=2e-3)
        dt = time.perf_counter() - t0
        ok = r["status"] == 0
        print(f"refined {int(ok.sum())} of {len(rows)} points to |Im c| <= 1e-10 in {dt * 1e3:.1f} ms "
              f"({len(rows) / dt:.0f} neutral points/s, {r['steps'].mean():.2f} secant steps each)")
        rows[ok, 1] = r["alpha"][ok] / bench.S2
        rows[ok, 2] = r["c"].real[ok]
    np.savetxt(args.out, rows, header="Re alpha c_r branch(+1 lower, -1 upper)")
    unstable = (ci > 0).any(axis=1)
    print(f"{len(rows)} neutral points; first unstable row Re = {Re_in[unstable][0]:.2f} (Re_delta* = {Re_in[unstable][0] * 1.7207876:.1f})")
    prof.free()
    eng.close()


if __name__ == "__main__":
    main()
