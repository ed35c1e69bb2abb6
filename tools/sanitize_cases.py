#!/usr/bin/env python
"""Every kernel family once, at the smallest sizes that still take every code path (ragged warps, partial panels,
chunk edges), for a run under compute-sanitizer:

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_cases.py shared
    compute-sanitizer --tool synccheck --error-exitcode 1 python tools/sanitize_cases.py shared

`shared` restricts the run to the kernels that communicate through shared memory (the matrix path and the queue
kernel).  Prints one line per case; results are checked for finiteness / convergence only (parity is the tests' job).
(On the B200 pool of this build compute-sanitizer is closed by the operators; the plain run -- 2.5 s -- is what was
executed there: every kernel family and every A/B form at ragged sizes, all finite and converged.)"""
import math
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os_stab_b200 as osb  # noqa: E402

S2 = math.sqrt(2.0)
only_shared = len(sys.argv) > 1 and sys.argv[1] == "shared"
eng = osb.Engine(0)


def say(name, ok):
    print(f"{name:58s} {'ok' if ok else 'FAILED'}", flush=True)
    if not ok:
        sys.exit(2)


def setenv(**kw):
    for k in ("OSB_NO_PAIR", "OSB_PAIR_V1", "OSB_PAIR_LPW", "OSB_QUEUE_MIN", "OSB_STRICT", "OSB_FD_THREAD", "OSB_FD_DENSE",
              "OSB_CHAN_BIG", "OSB_CHAN_NO_WARP", "OSB_CHAN_NO_VERIFY"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in kw.items()})


# ---------------- shooting path
nstep = 200
bl = eng.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, nstep, 0.0, 17.0, 0.5, 0.0)
opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
opts.nstep = nstep
n = 37
al = (0.179 + 0.0005 * (np.arange(n) - 18)) * S2
Re = np.full(n, 580.0 * S2)
c0 = np.full(n, complex(0.3641, 0.0080))
if not only_shared:
    say("K1 blasius (single) + stage table", bool(np.isfinite(bl.tables()["U"]).all()))
    o = eng.solve_bl_batch(np.linspace(0.45, 0.6, 19), np.linspace(-0.02, 0.1, 19), osb.FLOW_CONTEBL, osb.INT_CK45, nstep, 0.0, 12.0)
    say("K1 blasius batch (19 profiles) + transpose", bool(np.isfinite(o["U"]).all()))
    for name, env in (("pair kernel, single-loop form, auto", {}), ("pair kernel, 3 lines per warp", {"OSB_PAIR_V1": 0, "OSB_PAIR_LPW": 3}),
                      ("pair kernel, loop-nest form", {"OSB_PAIR_V1": 1}), ("one-thread kernel", {"OSB_NO_PAIR": 1}),
                      ("strict-parity kernel", {"OSB_STRICT": 1})):
        setenv(**env)
        r = eng.shoot_temporal(opts, bl, al, Re, c0, want_resid=True, hist_cap=4)
        say("K2 " + name, bool((r["status"] == 0).mean() > 0.9))
setenv(OSB_QUEUE_MIN=512)
nq = 700
r = eng.shoot_temporal(opts, bl, (0.179 + 0.00002 * np.arange(nq)) * S2, np.full(nq, 580.0 * S2), np.full(nq, complex(0.3641, 0.0080)))
say("K2 queue kernel (700 points, ragged last CTA)", bool((r["status"] == 0).mean() > 0.9))
setenv()
if not only_shared:
    e = eng.eigfun(opts, bl, Re[:3], r["c"][:3], alpha=al[:3])
    say("K2b eigenfunction pass", bool(np.isfinite(e["y"]).all()))
    fact = S2 / 1.7207876
    sp = eng.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, nstep, 0.0, 20.0 * fact, 0.5, 0.0)
    so = osb.shoot_defaults(osb.FLOW_ORRSPACE)
    so.nstep, so.ymin, so.ymax, so.testalpha = nstep, 0.0, 20.0 * fact, 0.9
    for name, env in (("pair", {}), ("one-thread", {"OSB_NO_PAIR": 1})):
        setenv(**env)
        r3 = eng.shoot_spatial_lines(so, sp, np.array([1000.0, 990.0, 1010.0, 950.0, 1000.0]) * fact, np.full(5, complex(0.042, 0.0)) * fact,
                                     np.full(5, 0.002) * fact, np.full(5, complex(0.13809592, 0.0051958399)) * fact, 4)
        say(f"K2 continuation lines ({name} kernel)", bool(np.isfinite(r3["alpha"]).all()))
    setenv()
    sp.free()
bl.free()

# ---------------- matrix path
setenv(OSB_CHAN_NO_VERIFY=1)
for disc, N, env, name in ((osb.DISC_CHEB, 24, {}, "warp kernel"), (osb.DISC_CHEB, 64, {}, "warp kernel"),
                           (osb.DISC_CHEB, 64, {"OSB_CHAN_NO_WARP": 1}, "CTA kernel, 128 threads"),
                           (osb.DISC_CHEB, 101, {}, "CTA kernel, 128 threads"), (osb.DISC_CHEB, 190, {}, "two-level LU, two CTAs/SM"),
                           (osb.DISC_CHEB, 190, {"OSB_CHAN_BIG": 1}, "two-level LU, 512 threads"),
                           (osb.DISC_CHEB, 190, {"OSB_CHAN_BIG": 0}, "single-level LU"),
                           (osb.DISC_FD, 75, {}, "banded, nine lanes"), (osb.DISC_FD, 75, {"OSB_FD_THREAD": 1}, "banded, one thread"),
                           (osb.DISC_FD, 75, {"OSB_FD_DENSE": 1}, "FD operator on the dense kernel")):
    setenv(OSB_CHAN_NO_VERIFY=1, **env)
    ch = eng.chan(disc, N)
    alpha = np.linspace(0.95, 1.05, 7)
    sig = np.full(7, complex(0.2375, 0.0037))
    adj = None if name == "banded, nine lanes" else (osb.ADJ_PENCIL if disc == osb.DISC_CHEB else osb.ADJ_INVBA)  # nine lanes: no adjoint form
    r = ch.solve(alpha, 1.0e4, sigma0=sig, want_evec=True, want_adj=adj)
    say(f"K4 N={N} {name} (+ eigenvector{', adjoint' if adj is not None else ''})", bool(np.isfinite(r["omega"]).all() and np.isfinite(r["evec"]).all()))
    ch.free()
setenv(OSB_CHAN_NO_VERIFY=1)
ch = eng.chan(osb.DISC_CHEB, 32)
w = ch.spectrum(np.array([1.0]), 1.0e4)
say("K4c full spectrum N=32", bool(np.isfinite(w["spectrum"]).all()))
ch.free()
for N, env, name in ((64, {}, "warp kernel"), (64, {"OSB_CHAN_NO_WARP": 1}, "CTA kernel")):
    setenv(OSB_CHAN_NO_VERIFY=1, **env)
    ch = eng.chan(osb.DISC_CHEB_NUM, N)
    rn = ch.neutral([10000.0, 9000.0, 12000.0], 1.0, sfac=0.01, maxit=40)
    say(f"orrncbl Newton search N={N} ({name})", bool(np.isfinite(rn["alpha"]).all()))
    ch.free()
setenv()
eng.close()
print("all cases ran")
