"""Where config 5 does not converge (maxcount / non-finite): does the CPU oracle fail at the same points?"""
import sys, math, pathlib
import numpy as np
R = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
import os_stab_b200 as osb, oracle_binding as ob, bench

o = ob.load(); eng = osb.Engine(0); prof = eng.solve_bl(); oprof, _, _ = o.solve_bl()
opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
idx, nre = bench.workload_indices(1, 0)
al, Re, c0 = bench.build_inputs(idx, nre)
r = eng.shoot_temporal(opts, prof, al, Re, c0)
bad = np.nonzero(r["status"] != 0)[0]
print("full grid", idx.size, "points; GPU status counts", np.bincount(r["status"], minlength=3))
ref = o.shoot_temporal_batch(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, oprof, al[bad], Re[bad], c0[bad])
print("oracle status at the GPU's failing points:", np.bincount(ref["status"], minlength=3), "(0 conv, 1 maxcount, 2 nonfinite)")
print("identical status fraction", (ref["status"] == r["status"][bad]).mean())
i, j = idx[bad] // 4096, idx[bad] % 4096
print("failing rows i (Re index) range", i.min(), i.max(), "cols j (alpha index) range", j.min(), j.max())
for k in range(0, bad.size, max(1, bad.size // 12)):
    b = bad[k]
    print(" i %4d j %4d alpha %.5f Re %.2f seed %.6f%+.6fi | gpu st %d passes %d c %.8f%+.8fi | oracle st %d passes %d c %.8f%+.8fi" % (
        i[k], j[k], al[b].real / math.sqrt(2), Re[b] / math.sqrt(2), c0[b].real, c0[b].imag, r["status"][b], r["passes"][b], r["c"][b].real, r["c"][b].imag,
        ref["status"][k], ref["passes"][k], ref["c"][k].real, ref["c"][k].imag))
