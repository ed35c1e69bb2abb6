"""GPU parity: the CUDA shooting path (through the C ABI) against the CPU oracle
on the same inputs.  Tolerances (BASELINE.json north_star): eigenvalues 1e-10
relative, eigenfunctions 1e-8 after normalisation, fp64.  Integer outputs
(passes, qend) are compared exactly where the decision margins allow it and as
a mismatch fraction on grids (a normalisation decision sitting within rounding
of its threshold may legitimately move by one step under FMA contraction)."""
import math

import numpy as np
import pytest

import oracle_binding as ob

pytestmark = pytest.mark.gpu
S2 = math.sqrt(2.0)
EIG_RTOL = 1e-10
EIGFUN_TOL = 1e-8


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)) / np.abs(np.asarray(b))


@pytest.fixture(scope="module")
def bl(engine, osb):
    p = engine.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
    yield p
    p.free()


@pytest.fixture(scope="module")
def bl_oracle(oracle):
    return oracle.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)


# ---------------------------------------------------------------- K1 profile
@pytest.mark.parametrize("variant,integ,nbl,ymax", [
    (ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 17.0),
    (ob.FLOW_CONTEBL, ob.INT_RK4, 1000, 17.0),
    (ob.FLOW_ORRSOM, ob.INT_RK4, 400, 30.0),
    (ob.FLOW_ORRSPACE, ob.INT_RK4, 1000, 20.0 * math.sqrt(2.0) / 1.7207876),
])
def test_profile_tables_bit_exact(engine, oracle, variant, integ, nbl, ymax):
    """K1 uses non-contracting arithmetic in the reference's order: bit-exact."""
    p = engine.solve_bl(variant, integ, nbl, 0.0, ymax, 0.5, 0.0)
    t = p.tables()
    p.free()
    prof, npass, f2p = oracle.solve_bl(variant, integ, nbl, 0.0, ymax, 0.5, 0.0)
    assert t["npass"] == npass
    assert t["f2p"] == f2p
    for k, name in enumerate(("ydat", "U", "Uspl", "d2U", "d2Uspl")):
        assert np.array_equal(t[name], prof[k][: nbl + 1]), name


def test_profile_batch_falkner_skan(engine, oracle, osb):
    beta = np.array([0.0, 0.1, -0.05, 0.3, 0.0, 0.5, 1.0])
    f2p = np.array([0.5, 0.6, 0.4, 0.8, 0.45, 0.9, 1.2])
    out = engine.solve_bl_batch(f2p, beta, osb.FLOW_CONTEBL, osb.INT_CK45, 600, 0.0, 12.0)
    nfinite = 0
    for i in range(beta.size):
        prof, npass, f2 = oracle.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, 600, 0.0, 12.0, f2p[i], beta[i])
        assert out["npass"][i] == npass
        # a secant that runs away gives NaN in the reference too: same NaNs, same places
        assert np.array_equal(out["f2p"][i], f2, equal_nan=True)
        assert np.array_equal(out["U"][i], prof[1][:601], equal_nan=True)
        assert np.array_equal(out["Uspl"][i], prof[2][:601], equal_nan=True)
        assert np.array_equal(out["d2Uspl"][i], prof[4][:601], equal_nan=True)
        nfinite += int(np.isfinite(prof[1][:601]).all())
    assert nfinite >= 4


# ---------------------------------------------------------------- K2 shooting
def test_contebl_default_point(engine, osb, oracle, bl, bl_oracle):
    """Config 1: README.md:24-26."""
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    r = engine.shoot_temporal(opts, bl, [0.179 * S2], [580.0 * S2], [complex(0.36413124, 0.79527387e-2)],
                              want_resid=True, hist_cap=20)
    ref = oracle.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, bl_oracle[0],
                                0.179 * S2, 580.0 * S2, complex(0.36413124, 0.79527387e-2))
    assert r["status"][0] == 0
    assert rel(r["c"][0], ref["c"]) < EIG_RTOL
    assert abs(r["c"][0] - complex(0.36412287, 0.0079597206)) < 1e-8
    assert r["passes"][0] == ref["passes"] == 5
    assert r["qend"][0] == ref["qend"] == 21
    # the stdout pass table (shoot.f:745-747).  Passes 1-2 integrate the same c as the
    # reference; |err| must agree, but its PHASE is fixed by rounding residues (DESIGN.md
    # "err(c) is rounding-phased") and differs between any two arithmetics, so the secant /
    # Muller iterates of passes 3-4 agree only to ~1e-10 while the converged value agrees to 1e-15.
    h = r["history"][0][:5]
    assert np.array_equal(h[:2, :2], ref["history"][:2, :2])
    assert np.max(np.abs(h[2:4, :2] - ref["history"][2:4, :2])) < 1e-9
    assert np.max(np.abs(h[4, :2] - ref["history"][4, :2])) < 1e-13
    mag = np.hypot(h[:, 2], h[:, 3])
    mag_ref = np.hypot(ref["history"][:, 2], ref["history"][:, 3])
    assert np.max(np.abs(mag[:2] / mag_ref[:2] - 1)) < 1e-9
    assert r["resid"][0, 0] < 1e-14 and r["resid"][0, 1] < 1e-14


def test_contebl_rk4_point(engine, osb, oracle):
    p = engine.solve_bl(osb.FLOW_CONTEBL, osb.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    opts.integrator = osb.INT_RK4
    r = engine.shoot_temporal(opts, p, [0.179 * S2], [580.0 * S2], [complex(0.36413124, 0.79527387e-2)])
    p.free()
    assert rel(r["c"][0], complex(0.36412271287, 0.0079597135036)) < 1e-10
    assert (r["passes"][0], r["qend"][0], r["status"][0]) == (5, 21, 0)


def test_conte_channel_point(engine, osb, oracle):
    """Config 2: conte defaults (conte.f:61-68)."""
    prof = engine.channel_profile()
    for integ, expect in ((osb.INT_CK45, complex(0.23752648882047672, 0.00373967062299434)),
                          (osb.INT_RK4, complex(0.23752648711564373, 0.0037396708967126395))):
        opts = osb.shoot_defaults(osb.FLOW_CONTE)
        opts.integrator = integ
        r = engine.shoot_temporal(opts, prof, [1.0], [10000.0], [complex(0.23753, 0.00374)], hist_cap=20)
        ref = oracle.shoot_temporal(ob.FLOW_CONTE, integ, 2000, 10.0, -1.0, 1.0, None, 1.0, 10000.0,
                                    complex(0.23753, 0.00374))
        assert rel(r["c"][0], ref["c"]) < EIG_RTOL
        assert rel(r["c"][0], expect) < EIG_RTOL
        assert r["qend"][0] == ref["qend"] == 115
        assert r["status"][0] == 0
        assert abs(r["passes"][0] - ref["passes"]) <= 1  # errtol 1e-15 sits at the rounding floor
    prof.free()


def test_contebl_grid_vs_oracle(engine, osb, oracle, bl, bl_oracle):
    """24 x 24 patch of the config-5 grid around the anchor, seeded by continuation."""
    na = nr = 24
    alpha = 0.179 + 0.0008 * (np.arange(na) - na // 2)
    Re = 580.0 * 1.004 ** (np.arange(nr) - nr // 2)
    A, R = np.meshgrid(alpha, Re)
    A, R = A.ravel() * S2, R.ravel() * S2
    # seeds: oracle solution at the anchor, everything else starts from it (|dc| ~ 1e-3)
    guess = np.full(A.size, complex(0.36412287, 0.0079597206))
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    r = engine.shoot_temporal(opts, bl, A, R, guess)
    ref = oracle.shoot_temporal_batch(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, bl_oracle[0], A, R, guess)
    ok = (ref["status"] == 0)
    assert ok.mean() > 0.95
    assert np.array_equal(r["status"][ok], ref["status"][ok])
    err = rel(r["c"][ok], ref["c"][ok])
    assert err.max() < EIG_RTOL, err.max()
    # integer outputs: identical except where a decision sits on its threshold
    # qend steps by one across the grid; points sitting on a step edge may flip
    assert (r["qend"][ok] != ref["qend"][ok]).mean() < 0.05
    assert np.abs(r["qend"][ok] - ref["qend"][ok]).max() <= 1
    assert (np.abs(r["passes"][ok] - ref["passes"][ok]) > 1).mean() < 0.01
    assert (r["passes"][ok] != ref["passes"][ok]).mean() < 0.15


def test_complex_alpha_and_ragged_batches(engine, osb, oracle, bl, bl_oracle):
    """alpha with an imaginary part (contebl accepts alphai, contebl.f:73-74,175) and
    batch sizes that are not a multiple of the CTA size, including empty."""
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    empty = engine.shoot_temporal(opts, bl, np.zeros(0, complex), np.zeros(0), np.zeros(0, complex))
    assert empty["c"].size == 0
    for n in (1, 31, 33, 130):
        al = (0.179 + 0.00001 * np.arange(n) + 1j * 0.0005 * (np.arange(n) % 3)) * S2
        Re = np.full(n, 580.0 * S2)
        guess = np.full(n, complex(0.36412287, 0.0079597206))
        r = engine.shoot_temporal(opts, bl, al, Re, guess)
        ref = oracle.shoot_temporal_batch(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, bl_oracle[0], al, Re, guess)
        ok = ref["status"] == 0
        assert ok.all()
        assert rel(r["c"], ref["c"]).max() < EIG_RTOL


def test_maxcount_and_nonfinite_are_reported(engine, osb, bl):
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    # far outside the basin of the iteration: must not claim convergence
    r = engine.shoot_temporal(opts, bl, [0.179 * S2, 0.179 * S2], [580.0 * S2, 580.0 * S2],
                              [complex(0.9, 0.3), complex(0.36413124, 0.0079527387)])
    assert r["status"][1] == osb.ST_CONVERGED
    assert r["status"][0] in (osb.ST_MAXCOUNT, osb.ST_NONFINITE) or abs(r["c"][0] - r["c"][1]) > 1e-3
    # underflow domain: first Gram-Schmidt divides by zero (contebl.f:80) -> flagged, not silent
    r = engine.shoot_temporal(opts, bl, [0.4 * S2], [3000.0 * S2], [complex(0.3, 0.0)])
    assert r["status"][0] in (osb.ST_NONFINITE, osb.ST_MAXCOUNT)


def test_orrsom_decks(engine, osb, oracle):
    for nbl, nstep, ymax, al, Re, guess, expect, pq in (
            (400, 4000, 30.0, 0.179, 580.0, complex(0.36413124, 0.0079527387),
             complex(0.3640704766590239, 0.008003326110564948), (5, 39)),
            (1000, 1000, 17.0, 0.1453, 581.1, complex(0.35, 0.01),
             complex(0.3497873683492642, 0.01208610121615455), (7, 18))):
        p = engine.solve_bl(osb.FLOW_ORRSOM, osb.INT_RK4, nbl, 0.0, ymax, 0.5, 0.0)
        opts = osb.shoot_defaults(osb.FLOW_ORRSOM)
        opts.nstep, opts.ymax = nstep, ymax
        r = engine.shoot_temporal(opts, p, [al * S2], [Re * S2], [guess])
        p.free()
        assert rel(r["c"][0], expect) < EIG_RTOL
        assert (r["passes"][0], r["qend"][0]) == pq
        assert r["status"][0] == 0


def test_orrspace_sweep_line(engine, osb, oracle):
    """Config 3: sweep.inp, 50 omega steps with continuation inside the kernel.
    Points whose Muller iteration wanders (17-32 passes) amplify rounding:
    1e-10 is asserted where the oracle needed <= 10 passes, 5e-9 elsewhere."""
    fact = math.sqrt(2.0) / 1.7207876
    ymin, ymax = 0.0, 20.0 * fact
    p = engine.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
    opts = osb.shoot_defaults(osb.FLOW_ORRSPACE)
    opts.nstep, opts.ymin, opts.ymax, opts.testalpha = 1000, ymin, ymax, 0.9
    nl = 3
    Re = np.array([1000.0, 1000.0, 990.0]) * fact
    om = np.full(nl, complex(0.042, 0.0)) * fact
    dom = np.full(nl, 0.002) * fact
    a0 = np.full(nl, complex(0.13809592, 0.0051958399)) * fact
    r = engine.shoot_spatial_lines(opts, p, Re, om, dom, a0, 50)
    p.free()
    prof, _, _ = oracle.solve_bl(ob.FLOW_ORRSPACE, ob.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
    for line in (0, 2):
        ref = oracle.shoot_spatial_line(ob.INT_RK4, 1000, 0.9, ymin, ymax, prof, Re[line], om[line], dom[line], a0[line], 50)
        assert np.all(r["status"][line] == 0)
        e = rel(r["alpha"][line], ref["alpha"])
        easy = ref["passes"] <= 10
        assert e[easy].max() < EIG_RTOL, e[easy].max()
        assert e.max() < 5e-9, e.max()
    assert np.array_equal(r["alpha"][0], r["alpha"][1])  # identical lines -> identical bits


# ---------------------------------------------------------------- K2b eigenfunction
def test_eigenfunction_contebl(engine, osb, oracle, bl, bl_oracle):
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    r = engine.shoot_temporal(opts, bl, [0.179 * S2], [580.0 * S2], [complex(0.36413124, 0.79527387e-2)])
    e = engine.eigfun(opts, bl, [580.0 * S2], r["c"], alpha=[0.179 * S2])
    ref = oracle.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, bl_oracle[0],
                                0.179 * S2, 580.0 * S2, complex(0.36413124, 0.79527387e-2), eigfun=True)
    y = e["y"][0] / e["vmax"][0]
    yr = ref["y"] / ref["vmax"]
    assert e["qend"][0] == ref["qend"]
    for i in range(4):
        scale = np.max(np.abs(yr[:, i]))
        assert np.max(np.abs(y[:, i] - yr[:, i])) / scale < EIGFUN_TOL


def test_eigenfunction_batches_are_chunked(engine, osb, bl, monkeypatch):
    """The second pass keeps 272 B per step per point of scratch; osb_eigfun walks large batches in
    chunks (8 GB by default).  Chunked and unchunked results must be identical."""
    n = 10
    alpha = (0.179 + 0.001 * np.arange(n)) * S2
    Re = np.full(n, 580.0 * S2)
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    r = engine.shoot_temporal(opts, bl, alpha, Re, np.full(n, complex(0.36412287, 0.0079597206)))
    assert (r["status"] == 0).all()
    whole = engine.eigfun(opts, bl, Re, r["c"], alpha=alpha)
    monkeypatch.setenv("OSB_EIGFUN_CHUNK", "3")
    parts = engine.eigfun(opts, bl, Re, r["c"], alpha=alpha)
    monkeypatch.delenv("OSB_EIGFUN_CHUNK")
    for k in ("y", "vmax", "qend"):
        assert np.array_equal(whole[k], parts[k]), k
    assert np.abs(whole["y"]).max() > 0


def test_eigenfunction_conte(engine, osb, oracle):
    prof = engine.channel_profile()
    opts = osb.shoot_defaults(osb.FLOW_CONTE)
    r = engine.shoot_temporal(opts, prof, [1.0], [10000.0], [complex(0.23753, 0.00374)])
    e = engine.eigfun(opts, prof, [10000.0], r["c"], alpha=[1.0])
    prof.free()
    ref = oracle.shoot_temporal(ob.FLOW_CONTE, ob.INT_CK45, 2000, 10.0, -1.0, 1.0, None, 1.0, 10000.0,
                                complex(0.23753, 0.00374), eigfun=True)
    y = e["y"][0] / e["vmax"][0]
    yr = ref["y"] / ref["vmax"]
    for i in range(4):
        scale = np.max(np.abs(yr[:, i]))
        assert np.max(np.abs(y[:, i] - yr[:, i])) / scale < EIGFUN_TOL


# ---------------------------------------------------------------- full-size property
def test_full_size_properties(engine, osb, bl):
    """At bench sizes the oracle is too slow; use size-independent properties:
    (i) results do not depend on batch composition/order (points are independent),
    (ii) a converged c fed back as the guess stays put (idempotence) to 1e-12."""
    n = 200_000
    rng = np.random.default_rng(7)
    al = (0.179 + 0.004 * (rng.random(n) - 0.5)) * S2
    Re = 580.0 * (1 + 0.02 * (rng.random(n) - 0.5)) * S2
    guess = np.full(n, complex(0.36412287, 0.0079597206))
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    r = engine.shoot_temporal(opts, bl, al, Re, guess)
    assert (r["status"] == 0).mean() > 0.999
    perm = rng.permutation(n)
    r2 = engine.shoot_temporal(opts, bl, al[perm], Re[perm], guess[perm])
    assert np.array_equal(r2["c"], r["c"][perm])
    assert np.array_equal(r2["passes"], r["passes"][perm])
    sub = slice(0, 20_000)
    r3 = engine.shoot_temporal(opts, bl, al[sub], Re[sub], r["c"][sub])
    ok = (r["status"][sub] == 0) & (r3["status"] == 0)
    assert rel(r3["c"][ok], r["c"][sub][ok]).max() < 1e-11


def test_multi_device_context_matches_single(osb, engine, bl):
    """osb_create_multi (one host thread driving several GPUs, the deployment of the
    single-process Fortran hosts): results must be bit-identical to one device."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = 50_001  # 13 cyclic tiles of 4096 points, the last one ragged
    al = (0.179 + 0.0000001 * np.arange(n)) * S2
    Re = 580.0 * (1.0 + 0.0000004 * np.arange(n)) * S2
    guess = np.full(n, complex(0.36412287, 0.0079597206))
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    one = engine.shoot_temporal(opts, bl, al, Re, guess, want_resid=True, hist_cap=4)
    eng2 = osb.Engine(devices=[0, 1])
    p2 = eng2.solve_bl()
    two = eng2.shoot_temporal(opts, p2, al, Re, guess, want_resid=True, hist_cap=4)
    p2.free()
    eng2.close()
    for k in ("c", "passes", "qend", "status", "resid", "history"):
        assert np.array_equal(one[k], two[k]), k


def test_multi_device_context_shards_every_entry_point(osb, engine, bl):
    """every batched entry point shards by point / line / profile / Reynolds number over the devices of
    a multi-device context (cyclic tiles for the shooting sweeps and continuation lines, contiguous shares with one
    worker thread per device elsewhere; every device holds its own copy of the tables) -- and returns what one
    device returns, bit for bit."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    eng2 = osb.Engine(devices=[0, 1])
    same = lambda a, b, keys: [np.testing.assert_array_equal(a[k], b[k], err_msg=k) for k in keys]
    # spatial continuation lines (orrspace), cyclic tiles of lines
    fact = math.sqrt(2.0) / 1.7207876
    ymax = 20.0 * fact
    so = osb.shoot_defaults(osb.FLOW_ORRSPACE)
    so.nstep, so.ymin, so.ymax, so.testalpha = 1000, 0.0, ymax, 0.9
    nl = 301
    Re = (900.0 + 0.5 * np.arange(nl)) * fact
    args = (Re, np.full(nl, complex(0.042, 0.0)) * fact, np.full(nl, 0.002) * fact, np.full(nl, complex(0.13809592, 0.0051958399)) * fact, 3)
    p1 = engine.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, 1000, 0.0, ymax, 0.5, 0.0)
    p2 = eng2.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, 1000, 0.0, ymax, 0.5, 0.0)
    same(engine.shoot_spatial_lines(so, p1, *args, hist_cap=3), eng2.shoot_spatial_lines(so, p2, *args, hist_cap=3),
         ("alpha", "passes", "qend", "status", "history"))
    p1.free(); p2.free()
    # eigenfunction pass, chunked over the devices
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    n = 11
    al = (0.179 + 0.001 * np.arange(n)) * S2
    Rt = np.full(n, 580.0 * S2)
    c = engine.shoot_temporal(opts, bl, al, Rt, np.full(n, complex(0.36412287, 0.0079597206)))["c"]
    q2 = eng2.solve_bl()
    same(engine.eigfun(opts, bl, Rt, c, alpha=al), eng2.eigfun(opts, q2, Rt, c, alpha=al), ("y", "vmax", "qend"))
    q2.free()
    # batched profiles
    beta = np.linspace(-0.05, 0.3, 37)
    f2p = 0.47 + 0.9 * beta
    same(engine.solve_bl_batch(f2p, beta, nbl=400, ymax=12.0), eng2.solve_bl_batch(f2p, beta, nbl=400, ymax=12.0),
         ("U", "Uspl", "d2U", "d2Uspl", "f2p", "npass"))
    # matrix path: dense (Chebyshev) and banded (FD) sweeps with and without shifts, adjoints, spectrum, Newton search
    alpha = np.linspace(0.8, 1.1, 23)
    for disc, N in ((osb.DISC_CHEB, 64), (osb.DISC_CHEB, 96), (osb.DISC_FD, 128)):
        c1, c2 = engine.chan(disc, N), eng2.chan(disc, N)
        r1 = c1.solve(alpha, 1.0e4, want_evec=True, want_adj=osb.ADJ_PENCIL)
        r2 = c2.solve(alpha, 1.0e4, want_evec=True, want_adj=osb.ADJ_PENCIL)
        same(r1, r2, ("omega", "iters", "status", "evec", "avec"))
        g1, g2 = c1.solve(alpha, 1.0e4, sigma0=r1["omega"] * (1 + 1e-4)), c2.solve(alpha, 1.0e4, sigma0=r1["omega"] * (1 + 1e-4))
        same(g1, g2, ("omega", "iters", "status"))
        if N <= 96:
            same(c1.spectrum(alpha[:5], 1.0e4), c2.spectrum(alpha[:5], 1.0e4), ("spectrum", "status"))
        c1.free(); c2.free()
    n1, n2 = engine.chan(osb.DISC_CHEB_NUM, 64), eng2.chan(osb.DISC_CHEB_NUM, 64)
    Rn = 6000.0 * (1.0 + np.arange(9) / 8.0)
    same(n1.neutral(Rn, 1.0, sfac=0.05, want_trace=True), n2.neutral(Rn, 1.0, sfac=0.05, want_trace=True),
         ("alpha", "omega", "steps", "status", "trace"))
    n1.free(); n2.free()
    eng2.close()


def test_blasius_neutral_curve_critical_reynolds(engine, bl):
    """SURVEY 8f item 4 (parity unpinned: the reference has no Blasius neutral curve).  Physical
    check: the critical Reynolds number of the Blasius layer, Re_delta* = 519.4 (R = 301.8)."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import neutral_from_grid as nf

    rows, ci, Re_in, alpha_in = nf.neutral_curve(engine, bl, na=256, nre=96)
    assert len(rows) > 100
    unstable = (ci > 0).any(axis=1)
    Rcrit = Re_in[unstable][0]
    assert 300.0 <= Rcrit < 310.0, Rcrit          # grid rows are 1.9 % apart in Re
    # the nose of the curve: alpha ~ 0.176 (alpha delta* = 0.303), c_r ~ 0.397
    nose = rows[np.argmin(rows[:, 0])]
    assert abs(nose[1] - 0.176) < 0.01 and abs(nose[2] - 0.397) < 0.01
    # two branches at higher Re
    hi = rows[np.abs(rows[:, 0] - 1000.0) < 40.0]
    assert (hi[:, 3] > 0).any() and (hi[:, 3] < 0).any()


@pytest.mark.parametrize("n", [500, 1000, 2000, 4000, 8000, 16000])
def test_resolution_ladder_decks(engine, osb, oracle, n):
    """tiny / small / normal / big / bigger / biggest.inp of the reference: nbl = nstep = n,
    Ymax = 30 -- the largest sizes the reference ships (stage table 16000 x 6 rows)."""
    p = engine.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, n, 0.0, 30.0, 0.5, 0.0)
    t = p.tables()
    prof, npass, _ = oracle.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, n, 0.0, 30.0, 0.5, 0.0)
    assert t["npass"] == npass and np.array_equal(t["U"], prof[1][: n + 1]) and np.array_equal(t["Uspl"], prof[2][: n + 1])
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    opts.nstep, opts.ymax = n, 30.0
    r = engine.shoot_temporal(opts, p, [0.179 * S2], [580.0 * S2], [complex(0.36413124, 0.0079527387)])
    p.free()
    ref = oracle.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, n, 10.0, 0.0, 30.0, prof, 0.179 * S2, 580.0 * S2,
                                complex(0.36413124, 0.0079527387))
    assert r["status"][0] == 0 and ref["status"] == 0
    assert rel(r["c"][0], ref["c"]) < EIG_RTOL
    assert abs(r["qend"][0] - ref["qend"]) <= 1
    # errtol = 1e-14 is at the rounding floor of err; after 16000 steps the floor itself is ~1e-14,
    # so the number of passes spent bouncing on it depends on the arithmetic (oracle 6, FMA 9)
    assert abs(r["passes"][0] - ref["passes"]) <= (1 if n <= 4000 else 6)


def test_contebl_inp_deck(engine, osb, oracle, bl, bl_oracle):
    """contebl.inp: alpha = 0.1453, Re = 581.1, guess 0.35 + 0.01i (a far guess: 7 passes)."""
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    r = engine.shoot_temporal(opts, bl, [0.1453 * S2], [581.1 * S2], [complex(0.35, 0.01)], hist_cap=20)
    ref = oracle.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, bl_oracle[0], 0.1453 * S2, 581.1 * S2,
                                complex(0.35, 0.01))
    assert r["status"][0] == 0 and ref["status"] == 0
    assert rel(r["c"][0], ref["c"]) < EIG_RTOL
    assert (r["passes"][0], r["qend"][0]) == (ref["passes"], ref["qend"])
    assert abs(r["c"][0] - complex(0.34978737, 0.01208610)) < 1e-5  # the mode orrsom (RK4, linear-search splines) finds for this deck


def test_orrspace_space_deck_eigenfunction(engine, osb, oracle):
    """space.inp: one spatial point (omega = 0.08, testalpha = 0.2) with the eigenfunction pass."""
    fact = math.sqrt(2.0) / 1.7207876
    ymin, ymax = 0.0, 20.0 * fact
    p = engine.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
    opts = osb.shoot_defaults(osb.FLOW_ORRSPACE)
    opts.nstep, opts.ymin, opts.ymax, opts.testalpha = 1000, ymin, ymax, 0.2
    Re, om, a0 = 1000.0 * fact, complex(0.08, 0.0) * fact, complex(0.228047, -0.00651) * fact
    r = engine.shoot_spatial_lines(opts, p, [Re], [om], [0.0], [a0], 1)
    e = engine.eigfun(opts, p, [Re], r["alpha"][0], omega=[om])
    p.free()
    prof, _, _ = oracle.solve_bl(ob.FLOW_ORRSPACE, ob.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
    ref = oracle.shoot_spatial_point(ob.INT_RK4, 1000, 0.2, ymin, ymax, prof, Re, om, a0)
    assert r["status"][0, 0] == 0 and ref["status"] == 0
    assert rel(r["alpha"][0, 0], ref["alpha"]) < 1e-9      # errtol of orrspace is 1e-8 (orrspace.f:277)
    assert abs(r["alpha"][0, 0] / fact - complex(0.2318124, -0.0064175)) < 1e-6
    assert e["qend"][0] == ref["qend"]
    y, yr = e["y"][0] / e["vmax"][0], ref["y"] / ref["vmax"]
    for i in range(4):
        assert np.max(np.abs(y[:, i] - yr[:, i])) / np.max(np.abs(yr[:, i])) < 1e-7


@pytest.mark.parametrize("n", [512, 513, 1300, 37888 + 7])
def test_queue_kernel_equals_plain_kernel_bitwise(engine, osb, bl, n, monkeypatch):
    """Large batches go through the compacting persistent-CTA kernel, smaller ones through the lane-pair or
    the one-thread-per-point kernel: same per-point arithmetic, so chunked and unchunked results must be
    identical bit for bit (c, passes, qend, status, |err|, pass history), whatever the schedule --
    including complex alpha mixed in (per-warp real-alpha fast path) and ragged tails."""
    rng = np.random.default_rng(n)
    al = (0.179 + 0.004 * (rng.random(n) - 0.5)) * S2 + 0j
    al[::7] += 1j * 0.0003 * S2
    Re = 580.0 * (1 + 0.02 * (rng.random(n) - 0.5)) * S2
    guess = np.full(n, complex(0.36412287, 0.0079597206))
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    monkeypatch.setenv("OSB_QUEUE_MIN", "512")  # the queue kernel normally starts at 65536 points
    whole = engine.shoot_temporal(opts, bl, al, Re, guess, want_resid=True, hist_cap=6)
    monkeypatch.delenv("OSB_QUEUE_MIN")
    m = min(n, 1300)  # the chunked (plain-kernel) pass over the first m points
    for lo in range(0, m, 100):
        hi = min(lo + 100, m)
        part = engine.shoot_temporal(opts, bl, al[lo:hi], Re[lo:hi], guess[lo:hi], want_resid=True, hist_cap=6)
        for k in ("c", "passes", "qend", "status", "resid", "history"):
            assert np.array_equal(part[k], whole[k][lo:hi]), (k, lo)
    assert (whole["status"] == 0).mean() > 0.99


def test_config5_full_grid(engine, osb, oracle, bl, bl_oracle):
    """BASELINE config 5 at full size (4096 x 4096 points, the bench workload): every point solved,
    a spread sample checked against the oracle at 1e-10, and -- for ALL points -- the size-independent
    property that the converged c is a smooth function of (alpha, Re): neighbouring points of the
    grid differ by O(grid spacing), so a wrong mode or a garbage value anywhere would stand out."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import bench

    idx, nre = bench.workload_indices(1, 0)
    assert idx.size == 4096 * 4096
    al, Re, c0 = bench.build_inputs(idx, nre)
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    r = engine.shoot_temporal(opts, bl, al, Re, c0)
    ok = r["status"] == 0
    assert ok.mean() > 0.99999
    assert 5.0 < r["passes"][ok].mean() < 6.5
    c = r["c"].reshape(4096, 4096)
    okg = ok.reshape(4096, 4096)
    da = np.abs(np.diff(c, axis=1))[okg[:, 1:] & okg[:, :-1]]
    dr = np.abs(np.diff(c, axis=0))[okg[1:, :] & okg[:-1, :]]
    assert da.max() < 2e-3 and dr.max() < 2e-3, (da.max(), dr.max())  # grid spacing: 6e-5 in alpha, 4e-4 in ln Re
    pos = np.unique(np.floor(np.arange(400) * ((idx.size - 1) / 399.0)).astype(np.int64))
    ref = oracle.shoot_temporal_batch(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, bl_oracle[0], al[pos], Re[pos], c0[pos])
    both = (ref["status"] == 0) & ok[pos]
    assert both.mean() > 0.99
    assert rel(r["c"][pos][both], ref["c"][both]).max() < EIG_RTOL


def test_conte_and_orrsom_small_grids(engine, osb, oracle):
    """The other two temporal flows on small (alpha, Re) grids against the oracle."""
    # conte: channel, CK45, n = 2000
    prof = engine.channel_profile()
    A, R = np.meshgrid(np.linspace(0.98, 1.02, 5), np.linspace(9800.0, 10200.0, 5))
    g = np.full(A.size, complex(0.23752649, 0.00373967))
    opts = osb.shoot_defaults(osb.FLOW_CONTE)
    r = engine.shoot_temporal(opts, prof, A.ravel(), R.ravel(), g)
    prof.free()
    ref = oracle.shoot_temporal_batch(ob.FLOW_CONTE, ob.INT_CK45, 2000, 10.0, -1.0, 1.0, None, A.ravel() + 0j, R.ravel(), g)
    ok = (ref["status"] == 0) & (r["status"] == 0)
    assert ok.mean() > 0.9
    assert rel(r["c"][ok], ref["c"][ok]).max() < EIG_RTOL
    assert np.abs(r["qend"][ok] - ref["qend"][ok]).max() <= 1
    # orrsom: Blasius, RK4, spline-of-spline U'', loose tolerances (1e-8 / 1e-12)
    p = engine.solve_bl(osb.FLOW_ORRSOM, osb.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    oprof, _, _ = oracle.solve_bl(ob.FLOW_ORRSOM, ob.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    A, R = np.meshgrid(np.linspace(0.176, 0.182, 4), np.linspace(570.0, 590.0, 4))
    g = np.full(A.size, complex(0.3641, 0.0080))
    opts = osb.shoot_defaults(osb.FLOW_ORRSOM)
    opts.nstep, opts.ymax = 1000, 17.0
    r = engine.shoot_temporal(opts, p, A.ravel() * S2, R.ravel() * S2, g)
    p.free()
    ref = oracle.shoot_temporal_batch(ob.FLOW_ORRSOM, ob.INT_RK4, 1000, 10.0, 0.0, 17.0, oprof, A.ravel() * S2 + 0j, R.ravel() * S2, g)
    ok = (ref["status"] == 0) & (r["status"] == 0)
    assert ok.all()
    # orrsom stops at |err| < 1e-8: the converged values agree to that tolerance's consequence, ~1e-9
    assert rel(r["c"], ref["c"]).max() < 5e-9
    assert np.array_equal(r["passes"], ref["passes"])


def test_pair_kernel_equals_plain_kernel_bitwise(engine, osb, bl, monkeypatch):
    """Batches too small to fill the GPU run with two lanes per point (one per solution vector,
    exchanged by shuffles; step k+1 started while step k is still being tested, continuation points
    of a line taken up pass by pass).  The arithmetic per vector is that of the one-thread kernel, so every
    output must be identical bit for bit, for every flow variant, both integrators, odd batch
    sizes (the last pair of a warp without a neighbour pair) and lines that converge after
    different numbers of passes inside one warp."""
    fact = S2 / 1.7207876

    def both(fn):
        # every form of the pair kernel: the default choice, the loop-nest form (OSB_PAIR_V1=1), and the single-loop
        # form with the speculative step at 1, 3 (ragged) and 16 lines per warp -- against one thread per point
        keys = ("OSB_NO_PAIR", "OSB_PAIR_V1", "OSB_PAIR_LPW")
        outs = []
        for env in ({}, {"OSB_PAIR_V1": "1"}, {"OSB_PAIR_V1": "0", "OSB_PAIR_LPW": "1"},
                    {"OSB_PAIR_V1": "0", "OSB_PAIR_LPW": "3"}, {"OSB_PAIR_V1": "0", "OSB_PAIR_LPW": "16"},
                    {"OSB_NO_PAIR": "1"}):
            for k in keys:
                monkeypatch.delenv(k, raising=False)
            for k, val in env.items():
                monkeypatch.setenv(k, val)
            outs.append(fn())
        for k in keys:
            monkeypatch.delenv(k, raising=False)
        a = outs[0]
        for b in outs[1:]:
            for k in a:
                assert np.array_equal(a[k], b[k], equal_nan=True), k
        return a

    rng = np.random.default_rng(5)
    for n in (1, 17, 333):
        al = (0.179 + 0.01 * (rng.random(n) - 0.5)) * S2 + 0j
        al[::5] += 1j * 0.0003 * S2
        Re = 580.0 * (1 + 0.05 * (rng.random(n) - 0.5)) * S2
        guess = np.full(n, complex(0.36412287, 0.0079597206))
        for integ in (osb.INT_CK45, osb.INT_RK4):
            opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
            opts.integrator = integ
            r = both(lambda: engine.shoot_temporal(opts, bl, al, Re, guess, want_resid=True, hist_cap=8))
            assert (r["status"] == 0).mean() > 0.8, (n, integ, (r["status"] == 0).mean())
    assert len(set(r["passes"].tolist())) > 1  # divergent pass counts were exercised
    # conte (channel, e3/e4 start, Hermitian) and orrsom (analytic inner product)
    ch = engine.channel_profile()
    opts = osb.shoot_defaults(osb.FLOW_CONTE)
    al = np.linspace(0.95, 1.05, 9)
    both(lambda: engine.shoot_temporal(opts, ch, al, np.full(9, 1.0e4), np.full(9, complex(0.23753, 0.00374))))
    ch.free()
    p = engine.solve_bl(osb.FLOW_ORRSOM, osb.INT_RK4, 400, 0.0, 30.0, 0.5, 0.0)
    opts = osb.shoot_defaults(osb.FLOW_ORRSOM)
    opts.nstep, opts.ymax = 4000, 30.0
    both(lambda: engine.shoot_temporal(opts, p, np.array([0.179, 0.18, 0.175]) * S2, np.full(3, 580.0 * S2),
                                       np.full(3, complex(0.36413124, 0.0079527387))))
    p.free()
    # orrspace continuation lines
    p = engine.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, 1000, 0.0, 20.0 * fact, 0.5, 0.0)
    opts = osb.shoot_defaults(osb.FLOW_ORRSPACE)
    opts.nstep, opts.ymin, opts.ymax, opts.testalpha = 1000, 0.0, 20.0 * fact, 0.9
    Re = np.array([1000.0, 990.0, 1010.0, 950.0, 1000.0]) * fact
    r = both(lambda: engine.shoot_spatial_lines(opts, p, Re, np.full(5, complex(0.042, 0.0)) * fact,
                                                np.full(5, 0.002) * fact,
                                                np.full(5, complex(0.13809592, 0.0051958399)) * fact, 12))
    assert np.all(r["status"] == 0)
    # points that do NOT converge inside a line: every form must report the same status per (line, omega) and seed the
    # next omega from the same value -- a pass budget of 3 (every point ends by maxcount) and a line started far from
    # any mode (non-finite or maxcount along the way)
    opts.maxcount = 3
    r = both(lambda: engine.shoot_spatial_lines(opts, p, Re, np.full(5, complex(0.042, 0.0)) * fact,
                                                np.full(5, 0.002) * fact,
                                                np.full(5, complex(0.13809592, 0.0051958399)) * fact, 6))
    assert np.all(r["status"] == osb.ST_MAXCOUNT) and np.all(r["passes"] == 3)
    opts.maxcount = 12
    a0 = np.full(5, complex(0.13809592, 0.0051958399)) * fact
    a0[1] = complex(2.5, -1.5)   # nowhere near a mode
    a0[3] = complex(1.0e-9, 0.0)  # c = omega / alpha overflows the asymptotes
    r = both(lambda: engine.shoot_spatial_lines(opts, p, Re, np.full(5, complex(0.042, 0.0)) * fact,
                                                np.full(5, 0.002) * fact, a0, 6))
    assert (r["status"][[0, 4]] == 0).all()
    assert (r["status"][[1, 3]] != 0).any()
    p.free()


def test_device_allocation_failure_is_reported_and_recoverable(engine, osb, bl):
    """A request whose scratch cannot be allocated must come back as an error code with the CUDA
    message (no crash, no CPU path), release what it had allocated, and leave the context usable."""
    import ctypes as C

    n = 4096
    al = np.full(2 * n, 0.0); al[0::2] = 0.179 * S2
    Re = np.full(n, 580.0 * S2)
    cc = np.tile([0.36412287, 0.0079597206], n)
    ip = np.zeros(3 * n, dtype=np.int32)
    bogus_history = np.zeros(4)  # never written: the device buffer for it cannot be allocated
    dp, ipp = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    free0 = _free_bytes()
    for _ in range(3):
        rc = engine.lib.osb_shoot_temporal(
            engine.ctx, C.byref(opts), bl.handle, n, al.ctypes.data_as(dp), Re.ctypes.data_as(dp), cc.ctypes.data_as(dp),
            ip[:n].ctypes.data_as(ipp), ip[n:2 * n].ctypes.data_as(ipp), ip[2 * n:].ctypes.data_as(ipp), None,
            bogus_history.ctypes.data_as(dp), 2 ** 31 - 1)
        assert rc < 0
        assert b"memory" in engine.lib.osb_last_error(engine.ctx).lower()
    r = engine.shoot_temporal(opts, bl, [0.179 * S2], [580.0 * S2], [complex(0.36413124, 0.79527387e-2)])
    assert r["status"][0] == 0 and abs(r["c"][0] - complex(0.36412287, 0.0079597206)) < 1e-8
    # nothing of the failed calls stays allocated beyond the pool's cache of the small buffers
    assert free0 - _free_bytes() < 64 << 20


def _free_bytes():
    import torch

    return torch.cuda.mem_get_info(0)[0]


def test_repeated_calls_do_not_leak_device_memory(osb):
    """Every entry point that builds a stage table (temporal, spatial, eigenfunction, neutral search) releases
    it with the call's other scratch: device memory stays flat over 1000 calls, and a destroyed context gives
    its private pool back to the driver."""
    eng = osb.Engine(0)
    p = eng.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    opts.nstep = 4000  # 384 KB of stage table per call
    opts.maxcount = 2
    a, R, c = [0.179 * S2], [580.0 * S2], [complex(0.36413124, 0.79527387e-2)]
    for _ in range(20):
        eng.shoot_temporal(opts, p, a, R, c)
    free0 = _free_bytes()
    for k in range(1000):
        eng.shoot_temporal(opts, p, a, R, c)
        if k % 100 == 0:
            eng.eigfun(opts, p, R, c, alpha=a)
            eng.shoot_neutral(opts, p, R, [0.179 * S2], c, maxit=2)
    assert free0 - _free_bytes() < 8 << 20, free0 - _free_bytes()
    before = _free_bytes()
    p.free()
    eng.close()
    assert _free_bytes() >= before


def test_neutral_curve_by_secant_on_alpha(engine, osb, oracle, bl, bl_oracle):
    """osb_shoot_neutral (no reference counterpart): both branches of the Blasius neutral
    curve for 48 Reynolds numbers, started from the sign changes of a coarse grid.  Checks: Im c = 0
    to 1e-10 at the returned alpha, agreement with the grid contour, the CPU oracle sees the same
    neutral points, and the nose of the curve is where the literature puts it."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import neutral_from_grid as nf

    rows, ci, Re_in, alpha_in = nf.neutral_curve(engine, bl, na=128, nre=48)
    rows = rows[rows[:, 0] > 330.0]                       # away from the nose, where the branches merge
    assert len(rows) >= 60
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    Re = rows[:, 0] * S2
    a0 = rows[:, 1] * S2
    c0 = rows[:, 2] + 0j
    r = engine.shoot_neutral(opts, bl, Re, a0, c0, dalpha_rel=2e-3, tol=1e-10, maxit=20)
    assert (r["status"] == 0).all(), np.nonzero(r["status"])[0]
    assert np.abs(r["c"].imag).max() <= 1e-10
    assert r["steps"].max() <= 8
    assert np.abs(r["alpha"] / S2 - rows[:, 1]).max() < 2e-3   # the contour was linear interpolation on a coarse grid
    assert np.abs(r["c"].real - rows[:, 2]).max() < 5e-3
    # the oracle at the returned alpha: neutral to its own rounding
    prof = bl_oracle[0]
    for i in (0, len(rows) // 2, len(rows) - 1):
        ref = oracle.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof, r["alpha"][i], Re[i], r["c"][i])
        assert abs(ref["c"].imag) < 1e-9 and abs(ref["c"] - r["c"][i]) < 1e-9
    # a second call from perturbed starts lands on the same points
    r2 = engine.shoot_neutral(opts, bl, Re, r["alpha"] * 1.004, r["c"], dalpha_rel=-1e-3, tol=1e-10, maxit=20)
    assert (r2["status"] == 0).all()
    # |d Im c / d alpha| ~ 0.05..0.3 on the curve, so Im c to 1e-10 pins alpha to ~1e-8
    assert np.abs(r2["alpha"] - r["alpha"]).max() < 1e-7


def test_private_conte_copies_with_cash_karp(engine, osb, oracle):
    """orrsom / orrspace call the shared steppers of shoot.f, so a USE_RKCK45 build (gcc.mak:46-48) runs
    their private CONTE copies with CRKCK45 as well: the non-default kernel instantiations
    (analytic inner product + Cash-Karp, orrsom's loop rule + Cash-Karp) against the oracle."""
    fact = S2 / 1.7207876
    ymin, ymax = 0.0, 20.0 * fact
    p = engine.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
    prof, _, _ = oracle.solve_bl(ob.FLOW_ORRSPACE, ob.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
    opts = osb.shoot_defaults(osb.FLOW_ORRSPACE)
    opts.integrator = osb.INT_CK45
    opts.nstep, opts.ymin, opts.ymax, opts.testalpha = 1000, ymin, ymax, 0.9
    Re = np.array([1000.0, 900.0]) * fact
    om = np.full(2, complex(0.06, 0.0)) * fact
    dom = np.full(2, 0.004) * fact
    a0 = np.array([complex(0.18304, -0.00168), complex(0.1845, 0.0005)]) * fact
    r = engine.shoot_spatial_lines(opts, p, Re, om, dom, a0, 6)
    p.free()
    for line in range(2):
        ref = oracle.shoot_spatial_line(ob.INT_CK45, 1000, 0.9, ymin, ymax, prof, Re[line], om[line], dom[line], a0[line], 6)
        assert np.all(r["status"][line] == 0) and np.all(ref["status"] == 0)
        e = rel(r["alpha"][line], ref["alpha"])
        assert e.max() < 5e-9, e
        assert e[ref["passes"] <= 10].max() < EIG_RTOL
    p = engine.solve_bl(osb.FLOW_ORRSOM, osb.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    prof, _, _ = oracle.solve_bl(ob.FLOW_ORRSOM, ob.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    opts = osb.shoot_defaults(osb.FLOW_ORRSOM)
    opts.integrator = osb.INT_CK45
    opts.nstep, opts.ymax = 1000, 17.0
    al = np.array([0.1453, 0.16, 0.179]) * S2
    ReT = np.array([581.1, 581.1, 580.0]) * S2
    g = np.array([complex(0.35, 0.01), complex(0.355, 0.012), complex(0.3641, 0.008)])
    r = engine.shoot_temporal(opts, p, al, ReT, g)
    p.free()
    for i in range(3):
        ref = oracle.shoot_temporal(ob.FLOW_ORRSOM, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof, al[i], ReT[i], g[i])
        assert r["status"][i] == 0 and ref["status"] == 0
        assert rel(r["c"][i], ref["c"]) < 5e-9          # orrsom stops at |err| < 1e-8
        assert r["passes"][i] == ref["passes"] and r["qend"][i] == ref["qend"]


def test_eigenfunction_orrsom(engine, osb, oracle):
    """orrsom's second pass (orrsom.f:455-503) for test.inp-like settings, against the oracle."""
    p = engine.solve_bl(osb.FLOW_ORRSOM, osb.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    prof, _, _ = oracle.solve_bl(ob.FLOW_ORRSOM, ob.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    opts = osb.shoot_defaults(osb.FLOW_ORRSOM)
    opts.nstep, opts.ymax = 1000, 17.0
    al, Re, g = 0.179 * S2, 580.0 * S2, complex(0.36413124, 0.0079527387)
    r = engine.shoot_temporal(opts, p, [al], [Re], [g])
    e = engine.eigfun(opts, p, [Re], r["c"], alpha=[al])
    p.free()
    ref = oracle.shoot_temporal(ob.FLOW_ORRSOM, ob.INT_RK4, 1000, 10.0, 0.0, 17.0, prof, al, Re, g, eigfun=True)
    assert r["status"][0] == 0 and rel(r["c"][0], ref["c"]) < 5e-9
    assert e["qend"][0] == ref["qend"]
    y = e["y"][0] / e["vmax"][0]
    yr = ref["y"] / ref["vmax"]
    for i in range(4):
        scale = np.max(np.abs(yr[:, i]))
        assert np.max(np.abs(y[:, i] - yr[:, i])) / scale < 1e-7   # the eigenvalues themselves differ by ~1e-9 (errtol 1e-8)


@pytest.mark.parametrize("n,ymax", [(8, 4.0), (16, 6.0), (37, 8.0)])
def test_coarsest_resolutions(engine, osb, oracle, n, ymax):
    """nbl = nstep down to the smallest profile the C ABI accepts: physically meaningless, but the GPU must
    walk the same iteration as the oracle (same status; same eigenvalue when both converge)."""
    p = engine.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, n, 0.0, ymax, 0.5, 0.0)
    t = p.tables()
    prof, npass, _ = oracle.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, n, 0.0, ymax, 0.5, 0.0)
    assert t["npass"] == npass and np.array_equal(t["U"], prof[1][: n + 1])
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    opts.nstep, opts.ymax = n, ymax
    al = np.array([0.179, 0.25]) * S2
    Re = np.array([580.0, 400.0]) * S2
    g = np.array([complex(0.36, 0.008), complex(0.4, 0.0)])
    r = engine.shoot_temporal(opts, p, al, Re, g)
    p.free()
    for i in range(2):
        ref = oracle.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, n, 10.0, 0.0, ymax, prof, al[i], Re[i], g[i])
        if ref["status"] == 0 and r["status"][i] == 0:
            assert rel(r["c"][i], ref["c"]) < 1e-9, (n, i)
        else:
            # a non-converging iteration wanders chaotically (rounding-seeded): only the outcome class is compared
            assert (ref["status"] != 0) == (r["status"][i] != 0) or abs(r["passes"][i] - ref["passes"]) <= 20


@pytest.mark.parametrize("beta,f2p", [(0.1, 0.6), (-0.05, 0.4)])
def test_falkner_skan_profiles_through_the_shooter(engine, osb, oracle, beta, f2p):
    """beta != 0 (contebl.f:66, the Falkner-Skan family): profile and eigenvalues against the oracle."""
    p = engine.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 17.0, f2p, beta)
    t = p.tables()
    prof, npass, f2 = oracle.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 0.0, 17.0, f2p, beta)
    assert t["npass"] == npass and t["f2p"] == f2 and np.array_equal(t["U"], prof[1][:1001])
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    al = np.array([0.17, 0.2]) * S2
    Re = np.array([700.0, 900.0]) * S2
    # continuation from the Blasius eigenvalue in a few oracle steps would be the driver's job; here both
    # sides start from the same rough guess and must end on the same root
    g = np.full(2, complex(0.36, 0.005))
    r = engine.shoot_temporal(opts, p, al, Re, g)
    p.free()
    for i in range(2):
        ref = oracle.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof, al[i], Re[i], g[i])
        assert (ref["status"] == 0) == (r["status"][i] == 0)
        if ref["status"] == 0:
            assert rel(r["c"][i], ref["c"]) < EIG_RTOL
            assert r["qend"][i] == ref["qend"]


def test_self_seeding_grid_sweep(engine, osb, bl):
    """osb_shoot_grid_temporal: the config-5 (alpha, Re) plane from ONE anchor guess (the contebl defaults), no
    oracle-made seed table.  257 x 257 points of the grid; the nodes it shares with the committed 65 x 65 table
    (converged by the CPU oracle's own continuation, tests/golden/make_cfg5_seeds.py) must carry the same
    eigenvalues to 1e-10, the device-made seeds must sit inside the iteration's basin and
    the pass counts must be those of a well-seeded sweep."""
    from pathlib import Path

    seeds = np.load(Path(__file__).resolve().parent.parent / "bench_data" / "cfg5_seeds_65x65.npy")  # [i_Re, j_alpha]
    n = 257
    alpha = (0.05 + 0.25 * np.arange(n) / (n - 1)) * S2
    Re = 300.0 * 6.0 ** (np.arange(n) / (n - 1)) * S2
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    r = engine.shoot_grid_temporal(opts, bl, alpha, Re, 0.179 * S2, 580.0 * S2, complex(0.36413124, 0.79527387e-2), want_seed=True)
    ok = r["status"] == 0
    print("coarse phase:", r["info"], " converged", ok.mean(), " passes/point", r["passes"][ok].mean())
    assert r["info"]["given_up"] == 0
    assert ok.mean() > 0.9995
    sub = r["c"][::4, ::4]
    sok = ok[::4, ::4]
    assert rel(sub[sok], seeds[sok]).max() < 1e-10
    d = np.abs(r["seed"] - r["c"])[ok]
    hist = {f"<{t:g}": float(np.mean(d < t)) for t in (1e-6, 1e-5, 1e-4, 1e-3, 3e-3)}
    print("seed error histogram:", hist, " max", d.max())
    assert d.max() < 3e-3 and np.mean(d < 1e-4) > 0.9
    assert r["passes"][ok].mean() < 6.5
    # degenerate grids: a single point, a single row / column
    one = engine.shoot_grid_temporal(opts, bl, [0.2 * S2], [600.0 * S2], 0.179 * S2, 580.0 * S2, complex(0.36413124, 0.79527387e-2))
    assert one["status"][0, 0] == 0
    ref = engine.shoot_temporal(opts, bl, [0.2 * S2], [600.0 * S2], one["c"][0])
    assert rel(one["c"][0, 0], ref["c"][0]) < 1e-12
    row = engine.shoot_grid_temporal(opts, bl, alpha[100:140], [580.0 * S2], 0.179 * S2, 580.0 * S2, complex(0.36413124, 0.79527387e-2), coarse=5)
    assert (row["status"] == 0).all() and row["c"].shape == (1, 40)
    col = engine.shoot_grid_temporal(opts, bl, [0.179 * S2], Re[60:90], 0.179 * S2, 580.0 * S2, complex(0.36413124, 0.79527387e-2), coarse=3)
    assert (col["status"] == 0).all() and col["c"].shape == (30, 1)
