"""Generates bench_data/cfg5_seeds_65x65.npy: converged contebl eigenvalues on the
65x65 coarse (alpha, Re) grid of BASELINE config 5, by continuation
from the config-1 anchor.  bench.py interpolates this table bilinearly to seed
the 4096x4096 sweep; it is workload definition (synthetic input), not product.

Uses the CPU oracle as the fixture generator (test infrastructure); run once:
    python tests/golden/make_cfg5_seeds.py
alpha_j = 0.05 + 0.25 j/64,  Re_i = 300 * 6**(i/64)   (contebl INPUT units)
"""
import math
import sys
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
import oracle_binding as ob  # noqa: E402

NA = NR = 65
S2 = math.sqrt(2.0)
ALPHA = 0.05 + 0.25 * np.arange(NA) / (NA - 1)
RE = 300.0 * 6.0 ** (np.arange(NR) / (NR - 1))

_o = None
_prof = None


def _init():
    global _o, _prof
    if _o is None:
        _o = ob.load(rebuild=False)
        _prof, _, _ = _o.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)


def solve(alpha, Re, guess):
    _init()
    return _o.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, _prof,
                             alpha * S2, Re * S2, guess, hist_cap=20)


def walk(a0, r0, c0, a1, r1, cprev=None, depth=0):
    """Continuation from (a0,r0,c0) to (a1,r1): try directly, else bisect the path
    (log in Re).  cprev = (a,r,c) of an earlier point for linear extrapolation."""
    guess = c0
    if cprev is not None:
        ap, rp, cp = cprev
        # linear extrapolation along the line parameter
        d0 = math.hypot((a0 - ap) / 0.25, math.log(r0 / rp) / math.log(6.0))
        d1 = math.hypot((a1 - a0) / 0.25, math.log(r1 / r0) / math.log(6.0))
        if d0 > 0:
            guess = c0 + (c0 - cp) * d1 / d0
    r = solve(a1, r1, guess)
    ok = r["status"] == 0 and r["passes"] <= 12 and abs(r["c"] - guess) < 1.0e-2
    if ok:
        return r["c"]
    if depth > 6:
        raise RuntimeError(f"continuation failed near alpha={a1} Re={r1}")
    am, rm = 0.5 * (a0 + a1), math.sqrt(r0 * r1)
    cm = walk(a0, r0, c0, am, rm, cprev, depth + 1)
    return walk(am, rm, cm, a1, r1, (a0, r0, c0), depth + 1)


def do_column(args):
    """All Re for one alpha index j, starting from the solved node (j, i0)."""
    j, i0, cstart = args
    col = np.zeros(NR, dtype=np.complex128)
    col[i0] = cstart
    for rng in (range(i0 + 1, NR), range(i0 - 1, -1, -1)):
        prev = None
        cur = (ALPHA[j], RE[i0], cstart)
        for i in rng:
            c = walk(cur[0], cur[1], cur[2], ALPHA[j], RE[i], prev)
            col[i] = c
            prev, cur = cur, (ALPHA[j], RE[i], c)
    return j, col


def main():
    ob.build()
    _init()
    anchor = solve(0.179, 580.0, complex(0.36413124, 0.0079527387))
    assert anchor["status"] == 0
    j0 = int(np.argmin(abs(ALPHA - 0.179)))
    i0 = int(np.argmin(abs(np.log(RE / 580.0))))
    c00 = walk(0.179, 580.0, anchor["c"], ALPHA[j0], RE[i0])
    # row i0: all alpha
    row = np.zeros(NA, dtype=np.complex128)
    row[j0] = c00
    for rng in (range(j0 + 1, NA), range(j0 - 1, -1, -1)):
        prev = None
        cur = (ALPHA[j0], RE[i0], c00)
        for j in rng:
            c = walk(cur[0], cur[1], cur[2], ALPHA[j], RE[i0], prev)
            row[j] = c
            prev, cur = cur, (ALPHA[j], RE[i0], c)
    table = np.zeros((NR, NA), dtype=np.complex128)  # [i (Re), j (alpha)]
    with ProcessPoolExecutor(8) as ex:
        for j, col in ex.map(do_column, [(j, i0, row[j]) for j in range(NA)]):
            table[:, j] = col
    out = ROOT / "bench_data" / "cfg5_seeds_65x65.npy"
    np.save(out, table)
    print("saved", out)
    for (i, j) in ((0, 0), (0, 64), (64, 0), (64, 64)):
        print(f"alpha={ALPHA[j]:.3f} Re={RE[i]:.1f} c={table[i, j]:.10f}")


if __name__ == "__main__":
    main()
