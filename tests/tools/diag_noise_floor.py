"""Final |err| (the quantity tested against errtol = 1e-14) on the GPU vs the CPU oracle, config-5 sample."""
import sys, math, pathlib
import numpy as np
R = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
import os_stab_b200 as osb, oracle_binding as ob, bench

o = ob.load(); eng = osb.Engine(0); prof = eng.solve_bl(); oprof, _, _ = o.solve_bl()
opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
idx, nre = bench.workload_indices(1, 0)
rng = np.random.default_rng(3)
sub = np.sort(rng.choice(idx.size, 1500, replace=False))
al, Re, c0 = bench.build_inputs(idx[sub], nre)
r = eng.shoot_temporal(opts, prof, al, Re, c0, want_resid=True, hist_cap=20)
ge, oe, gp, op = [], [], [], []
h_g, h_o = [], []
for k in range(sub.size):
    rr = o.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, oprof, al[k], Re[k], c0[k], hist_cap=20)
    oe.append(rr["abserr"]); op.append(rr["passes"]); ge.append(r["resid"][k, 0]); gp.append(r["passes"][k])
    # |err| of every pass after the 4th: the noise floor samples
    for p in range(4, rr["passes"]): h_o.append(math.hypot(rr["history"][p, 2], rr["history"][p, 3]))
    for p in range(4, r["passes"][k]): h_g.append(math.hypot(r["history"][k, p, 2], r["history"][k, p, 3]))
ge, oe = np.array(ge), np.array(oe)
print("final |err|  gpu: median %.2e  90%% %.2e  max %.2e | oracle: median %.2e  90%% %.2e  max %.2e" % (
    np.median(ge), np.quantile(ge, .9), ge.max(), np.median(oe), np.quantile(oe, .9), oe.max()))
print("passes gpu mean %.3f oracle mean %.3f ; hist gpu" % (np.mean(gp), np.mean(op)), np.bincount(gp, minlength=12)[:12], "oracle", np.bincount(op, minlength=12)[:12])
h_g, h_o = np.array(h_g), np.array(h_o)
print("|err| of passes >= 5: gpu n=%d median %.2e frac<1e-14 %.3f | oracle n=%d median %.2e frac<1e-14 %.3f" % (
    h_g.size, np.median(h_g), (h_g < 1e-14).mean(), h_o.size, np.median(h_o), (h_o < 1e-14).mean()))
