"""Writes tests/golden/oracle_golden.json: outputs of the CPU oracle for the reference's own
decks / defaults (contebl d, conte d, test.inp, contebl.inp, sweep.inp, chan.inp, fd-64/128.inp
ends, nc.inp).  The reference cannot be run here (no Fortran compiler), so these are oracle
outputs, pinned against README.md:24-26 etc. by tests/test_oracle_kat.py; the fixture guards the
oracle against accidental change and gives the GPU tests a file-based expectation.

    python tests/golden/make_oracle_golden.py
"""
import json
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
import oracle_binding as ob  # noqa: E402

S2 = math.sqrt(2.0)


def cplx(z):
    return [float(np.real(z)), float(np.imag(z))]


def main():
    o = ob.load()
    g = {}
    prof, npass, f2p = o.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
    g["blasius_ck45_1000_17"] = dict(npass=npass, f2p=f2p, U_100=float(prof[1][100]), Uspl_100=float(prof[2][100]),
                                     U_sum=float(prof[1][:1001].sum()))
    r = o.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof, 0.179 * S2, 580.0 * S2,
                         complex(0.36413124, 0.79527387e-2), eigfun=True)
    y = r["y"] / r["vmax"]
    g["contebl_default"] = dict(c=cplx(r["c"]), passes=r["passes"], qend=r["qend"], history=r["history"].tolist(),
                                phi_every_100=[cplx(v) for v in y[::100, 0]], dphi_every_100=[cplx(v) for v in y[::100, 1]])
    r = o.shoot_temporal(ob.FLOW_CONTE, ob.INT_CK45, 2000, 10.0, -1.0, 1.0, None, 1.0, 10000.0, complex(0.23753, 0.00374))
    g["conte_default"] = dict(c=cplx(r["c"]), passes=r["passes"], qend=r["qend"])
    p2, _, _ = o.solve_bl(ob.FLOW_ORRSOM, ob.INT_RK4, 400, 0.0, 30.0, 0.5, 0.0)
    r = o.shoot_temporal(ob.FLOW_ORRSOM, ob.INT_RK4, 4000, 10.0, 0.0, 30.0, p2, 0.179 * S2, 580.0 * S2, complex(0.36413124, 0.0079527387))
    g["orrsom_test_inp"] = dict(c=cplx(r["c"]), passes=r["passes"], qend=r["qend"])
    fact = math.sqrt(2.0) / 1.7207876
    p3, _, _ = o.solve_bl(ob.FLOW_ORRSPACE, ob.INT_RK4, 1000, 0.0, 20.0 * fact, 0.5, 0.0)
    r = o.shoot_spatial_line(ob.INT_RK4, 1000, 0.9, 0.0, 20.0 * fact, p3, 1000.0 * fact, complex(0.042, 0) * fact, 0.002 * fact,
                             complex(0.13809592, 0.0051958399) * fact, 50)
    g["orrspace_sweep_inp"] = dict(fort17=[[float((w / fact).real), float((a / fact).real), float((a / fact).imag)]
                                           for w, a in zip(r["omega"], r["alpha"])], passes=r["passes"].tolist())
    c = ob.chan(ob.DISC_FD, 64)
    g["orrfdchan_fd64"] = {f"{a:.2f}": cplx(c.solve_fd(a, 1.0e4)["omega"]) for a in (0.76, 0.86, 0.96, 1.0, 1.06, 1.14)}
    c.close()
    c = ob.chan(ob.DISC_FD, 128)
    g["orrfdchan_fd128"] = {f"{a:.2f}": cplx(c.solve_fd(a, 1.0e4)["omega"]) for a in (0.76, 0.96, 1.0, 1.15)}
    c.close()
    c = ob.chan(ob.DISC_CHEB, 64)
    s = c.solve_col(1.0, 1.0e4)["spectrum"]
    g["orrcolchan_chan_inp"] = dict(idx62=cplx(s[62]), idx63=cplx(s[63]), idx64=cplx(s[64]), quad_exact=cplx(c.arbiter(1.0, 1.0e4, s[62])))
    c.close()
    c = ob.chan(ob.DISC_CHEB_NUM, 64)
    tr = c.neutral(10000.0, 0.01, 1.0)
    g["orrncbl_nc_inp"] = dict(steps=len(tr), first=tr[0].tolist(), last=tr[-1].tolist())
    c.close()
    out = ROOT / "tests" / "golden" / "oracle_golden.json"
    out.write_text(json.dumps(g, indent=1))
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
