"""The committed fixture tests/golden/oracle_golden.json (made by tests/golden/make_oracle_golden.py)
against (i) the oracle as built now -- bit for bit on the same machine class -- and (ii), on the GPU,
the CUDA path through the C ABI."""
import json
import math
from pathlib import Path

import numpy as np
import pytest

import oracle_binding as ob

G = json.loads((Path(__file__).resolve().parent / "golden" / "oracle_golden.json").read_text())
S2 = math.sqrt(2.0)


def z(p):
    return complex(p[0], p[1])


def test_oracle_reproduces_fixture(oracle):
    prof, npass, f2p = oracle.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
    b = G["blasius_ck45_1000_17"]
    assert (npass, f2p, float(prof[1][100])) == (b["npass"], b["f2p"], b["U_100"])
    r = oracle.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof, 0.179 * S2, 580.0 * S2,
                              complex(0.36413124, 0.79527387e-2))
    d = G["contebl_default"]
    assert abs(r["c"] - z(d["c"])) < 1e-15 and (r["passes"], r["qend"]) == (d["passes"], d["qend"])
    assert np.allclose(r["history"], np.array(d["history"]), rtol=0, atol=1e-15)
    c = ob.chan(ob.DISC_FD, 64)
    for a, w in G["orrfdchan_fd64"].items():
        assert abs(c.solve_fd(float(a), 1.0e4)["omega"] - z(w)) < 1e-12
    c.close()


@pytest.mark.gpu
def test_gpu_against_fixture(engine, osb):
    p = engine.solve_bl()
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    r = engine.shoot_temporal(opts, p, [0.179 * S2], [580.0 * S2], [complex(0.36413124, 0.79527387e-2)])
    d = G["contebl_default"]
    assert abs(r["c"][0] - z(d["c"])) / abs(z(d["c"])) < 1e-10 and r["qend"][0] == d["qend"]
    e = engine.eigfun(opts, p, [580.0 * S2], r["c"], alpha=[0.179 * S2])
    y = e["y"][0] / e["vmax"][0]
    assert np.max(np.abs(y[::100, 0] - np.array([z(v) for v in d["phi_every_100"]]))) < 1e-8
    assert np.max(np.abs(y[::100, 1] - np.array([z(v) for v in d["dphi_every_100"]]))) < 1e-8
    p.free()
    ch = engine.channel_profile()
    o2 = osb.shoot_defaults(osb.FLOW_CONTE)
    r = engine.shoot_temporal(o2, ch, [1.0], [10000.0], [complex(0.23753, 0.00374)])
    assert abs(r["c"][0] - z(G["conte_default"]["c"])) / abs(r["c"][0]) < 1e-10 and r["qend"][0] == G["conte_default"]["qend"]
    ch.free()
    fd = engine.chan(osb.DISC_FD, 64)
    al = sorted(G["orrfdchan_fd64"])
    w = fd.solve([float(a) for a in al], 1.0e4)["omega"]
    fd.free()
    ref = np.array([z(G["orrfdchan_fd64"][a]) for a in al])
    assert np.max(np.abs(w - ref) / np.abs(ref)) < 1e-9
    cb = engine.chan(osb.DISC_CHEB, 64)
    wc = cb.solve([1.0], 1.0e4)["omega"][0]
    cb.free()
    assert abs(wc - z(G["orrcolchan_chan_inp"]["quad_exact"])) / abs(wc) < 1e-10
    assert abs(wc - z(G["orrcolchan_chan_inp"]["idx62"])) / abs(wc) < 2e-9
