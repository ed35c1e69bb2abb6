"""Strict parity of the shooting path: integers exact, eigenvalues bit for bit.

The fast kernels regroup the RHS, fold the step into the tableau, contract to FMA and replace acos by a
squared comparison; their eigenvalues agree with the oracle to ~1e-14 but pass counts / qend can differ
where a decision sits on its threshold.  The STRICT build (OSB_STRICT=1,
os-stab_b200/csrc/shoot_strict.cu, compiled with -fmad=false) performs every operation in the reference's
order.  The chain of evidence:

  (1) oracle with libm's complex functions  ==  oracle with glibc's complex formulas restated over libm's
      REAL exp / sincos / hypot / acos                               bit for bit   (CPU, below)
  (2) the portable real functions of osb_pmath.h are within 1-2 ulp of libm's     (CPU, below)
  (3) oracle over the portable functions vs oracle over libm: eigenvalues ~1e-13   (CPU, below; what the
      choice of libm is worth -- any two builds of the reference differ by as much)
  (4) GPU strict build  ==  oracle over the same portable functions  bit for bit:
      c, passes, qend, status, the per-pass history                                (GPU)
  (5) GPU fast build vs GPU strict build: <= 1e-10 relative                        (GPU)
  (6) per-point sensitivity classes: tolerance max(1e-10, 10 x the movement of the
      oracle's own answer under a 1-ulp perturbation of its input)                (GPU)
"""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_binding as ob

S2 = math.sqrt(2.0)
GUESS = complex(0.36413124, 0.79527387e-2)
FACT = math.sqrt(2.0) / 1.7207876


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)) / np.abs(np.asarray(b))


def grid(na=24, nr=24):
    alpha = 0.179 + 0.0008 * (np.arange(na) - na // 2)
    Re = 580.0 * 1.004 ** (np.arange(nr) - nr // 2)
    A, R = np.meshgrid(alpha, Re)
    return A.ravel() * S2, R.ravel() * S2, np.full(A.size, complex(0.36412287, 0.0079597206))


@pytest.fixture()
def oracle_math(oracle):
    yield oracle
    oracle.set_math(None)


# ------------------------------------------------------------------ CPU: (1) (2) (3)
def test_complex_formulas_over_libm_reproduce_libm(oracle_math):
    """(1) cabs = hypot, glibc's csqrt and cexp formulas: mode "libm" must give the same bits as the
    libm complex functions on every flow (so the only thing mode "portable" changes is four real functions)."""
    o = oracle_math
    prof, _, _ = o.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
    A, R, G = grid(5, 5)
    A = A + 1j * np.where(np.arange(A.size) % 3 == 0, 0.0004, 0.0)
    sprof, _, _ = o.solve_bl(ob.FLOW_ORRSPACE, ob.INT_RK4, 400, 0.0, 20.0 * FACT, 0.5, 0.0)

    def run():
        out = []
        r = o.shoot_temporal_batch(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof, A, R, G)
        out += [r["c"].copy(), r["passes"].copy(), r["qend"].copy()]
        r = o.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof, 0.179 * S2, 580.0 * S2, GUESS, eigfun=True)
        out += [r["history"].copy(), r["y"].copy(), np.array([r["vmax"]])]
        r = o.shoot_temporal(ob.FLOW_CONTE, ob.INT_CK45, 400, 10.0, -1.0, 1.0, None, 1.0, 10000.0, complex(0.23753, 0.00374))
        out += [np.array([r["c"]]), r["history"].copy()]
        r = o.shoot_temporal(ob.FLOW_ORRSOM, ob.INT_RK4, 400, 10.0, 0.0, 17.0, prof, 0.179 * S2, 580.0 * S2, GUESS)
        out += [np.array([r["c"]]), r["history"].copy()]
        r = o.shoot_spatial_line(ob.INT_RK4, 400, 0.9, 0.0, 20.0 * FACT, sprof, 1000.0 * FACT, 0.042 * FACT, 0.002 * FACT,
                                 complex(0.13809592, 0.0051958399) * FACT, 4)
        out += [r["alpha"].copy(), r["passes"].copy()]
        return out

    o.set_math(None)
    a = run()
    o.set_math("libm")
    b = run()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_portable_functions_are_within_an_ulp_of_libm():
    """(2) exp, sin, cos, acos <= 1 ulp, hypot <= 2 ulp from libm over the argument ranges the path uses."""
    pm = ob.build_pmath()
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(7)
    n = 400000

    def ulps(a, b):
        return np.abs(a - b) / np.spacing(np.abs(b))

    x = -400.0 * rng.random(n)
    y = np.empty(n)
    pm.pmh_exp_v(C.c_long(n), x.ctypes.data_as(dp), y.ctypes.data_as(dp))
    assert ulps(y, np.exp(x)).max() <= 1.0
    x = 4000.0 * (rng.random(n) - 0.5)
    s, c = np.empty(n), np.empty(n)
    pm.pmh_sincos_v(C.c_long(n), x.ctypes.data_as(dp), s.ctypes.data_as(dp), c.ctypes.data_as(dp))
    assert ulps(s, np.sin(x)).max() <= 1.0 and ulps(c, np.cos(x)).max() <= 1.0
    a = rng.random(n) * 10.0 ** (-20 * rng.random(n))
    b = rng.random(n) * 10.0 ** (-20 * rng.random(n))
    h = np.empty(n)
    pm.pmh_hypot_v(C.c_long(n), a.ctypes.data_as(dp), b.ctypes.data_as(dp), h.ctypes.data_as(dp))
    assert ulps(h, np.hypot(a, b)).max() <= 2.0
    z = np.zeros(n)
    pm.pmh_hypot_v(C.c_long(n), a.ctypes.data_as(dp), z.ctypes.data_as(dp), h.ctypes.data_as(dp))
    assert np.array_equal(h, a)  # |x + 0i| is exact, as in libm
    x = np.concatenate([2.0 * rng.random(n) - 1.0, 1.0 - 10.0 ** (-16 * rng.random(1000)), [1.0, -1.0, 0.0]])
    y = np.empty(x.size)
    pm.pmh_acos_v(C.c_long(x.size), x.ctypes.data_as(dp), y.ctypes.data_as(dp))
    assert ulps(y[:n], np.arccos(x[:n])).max() <= 1.0
    assert np.abs(y[n:] - np.arccos(x[n:])).max() <= 4e-16


def test_choice_of_libm_moves_the_eigenvalue_by_rounding_only(oracle_math):
    """(3) the oracle over the portable functions against the oracle over libm: what swapping the four
    elementary functions is worth on the 24 x 24 grid.  Eigenvalues agree far inside 1e-10; pass counts may
    differ at a few points (errtol = 1e-14 sits on the rounding floor of err) --
    the same spread any two builds of the reference (x87 / SSE, another glibc) show."""
    o = oracle_math
    prof, _, _ = o.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
    A, R, G = grid(12, 12)
    o.set_math(None)
    a = o.shoot_temporal_batch(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof, A, R, G)
    o.set_math("portable")
    b = o.shoot_temporal_batch(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof, A, R, G)
    ok = (a["status"] == 0) & (b["status"] == 0)
    assert ok.mean() > 0.95
    d = rel(b["c"][ok], a["c"][ok])
    print(f"libm vs portable: max rel {d.max():.2e}, passes differ at {np.mean(a['passes'][ok] != b['passes'][ok]):.1%}, "
          f"qend differ at {np.mean(a['qend'][ok] != b['qend'][ok]):.1%}")
    assert d.max() < 1e-11
    assert np.abs(a["qend"][ok] - b["qend"][ok]).max() <= 1


# ------------------------------------------------------------------ GPU: (4) (5) (6)
@pytest.fixture()
def strict(monkeypatch):
    monkeypatch.setenv("OSB_STRICT", "1")
    yield monkeypatch


def _bits_equal(r, ref, what):
    assert np.array_equal(r["status"], ref["status"]), what
    assert np.array_equal(r["passes"], ref["passes"]), (what, np.nonzero(r["passes"] != ref["passes"])[0][:8])
    assert np.array_equal(r["qend"], ref["qend"]), (what, np.nonzero(r["qend"] != ref["qend"])[0][:8])
    same = r["c"].view(np.float64) == ref["c"].view(np.float64)
    nan_both = np.isnan(r["c"].view(np.float64)) & np.isnan(ref["c"].view(np.float64))
    assert (same | nan_both).all(), (what, rel(r["c"], ref["c"]).max())


@pytest.mark.gpu
def test_strict_build_is_bit_identical_on_the_contebl_grid(engine, osb, oracle_math, strict):
    """(4) 24 x 24 patch of the config-5 grid + complex alpha + the contebl.inp deck with its far guess."""
    o = oracle_math
    o.set_math("portable")
    prof, _, _ = o.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
    bl = engine.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    A, R, G = grid()
    A = np.concatenate([A, [0.1453 * S2, (0.179 + 0.0005j) * S2, (0.18 - 0.001j) * S2]])
    R = np.concatenate([R, [581.1 * S2, 580.0 * S2, 575.0 * S2]])
    G = np.concatenate([G, [complex(0.35, 0.01), GUESS, GUESS]])
    r = engine.shoot_temporal(opts, bl, A, R, G, want_resid=True, hist_cap=20)
    ref = o.shoot_temporal_batch(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof, A, R, G)
    _bits_equal(r, ref, "contebl grid")
    assert (ref["status"] == 0).mean() > 0.95
    # the whole stdout pass table of the default point (shoot.f:745-747), err included, bit for bit
    one = engine.shoot_temporal(opts, bl, [0.179 * S2], [580.0 * S2], [GUESS], want_resid=True, hist_cap=20)
    oref = o.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof, 0.179 * S2, 580.0 * S2, GUESS)
    assert one["passes"][0] == oref["passes"] and one["qend"][0] == oref["qend"]
    assert np.array_equal(one["history"][0][: oref["passes"]], oref["history"])
    assert one["resid"][0, 0] == oref["abserr"] and one["resid"][0, 1] == oref["absdc"]
    # RK4 build of contebl (gcc.mak without USE_RKCK45)
    opts.integrator = osb.INT_RK4
    blr = engine.solve_bl(osb.FLOW_CONTEBL, osb.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    profr, _, _ = o.solve_bl(ob.FLOW_CONTEBL, ob.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    r = engine.shoot_temporal(opts, blr, A[:64], R[:64], G[:64])
    ref = o.shoot_temporal_batch(ob.FLOW_CONTEBL, ob.INT_RK4, 1000, 10.0, 0.0, 17.0, profr, A[:64], R[:64], G[:64])
    _bits_equal(r, ref, "contebl RK4")
    bl.free(); blr.free()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [500, 1000, 2000, 4000, 8000, 16000])
def test_strict_build_is_bit_identical_on_the_ladder_decks(engine, osb, oracle_math, strict, n):
    """(4) the resolution ladder (small.inp ... bigger.inp: nbl = nstep = n); at n = 16000 the fast build
    differs from the oracle by up to 6 passes -- the strict build does not."""
    o = oracle_math
    o.set_math("portable")
    prof, _, _ = o.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, n, 0.0, 17.0, 0.5, 0.0)
    bl = engine.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, n, 0.0, 17.0, 0.5, 0.0)
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    opts.nstep = n
    r = engine.shoot_temporal(opts, bl, [0.179 * S2], [580.0 * S2], [GUESS], hist_cap=20)
    ref = o.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, n, 10.0, 0.0, 17.0, prof, 0.179 * S2, 580.0 * S2, GUESS)
    bl.free()
    assert (r["passes"][0], r["qend"][0], r["status"][0]) == (ref["passes"], ref["qend"], ref["status"])
    assert r["c"][0] == ref["c"]
    assert np.array_equal(r["history"][0][: ref["passes"]], ref["history"])


@pytest.mark.gpu
def test_strict_build_is_bit_identical_on_conte_and_orrsom(engine, osb, oracle_math, strict):
    """(4) conte (channel, unit-vector start, <= test, errtol 1e-15) and orrsom (its own scaled RHS
    orrsom.f:637-639, linear-search splines, 180 acos / pi <= testalpha), both integrators."""
    o = oracle_math
    o.set_math("portable")
    ch = engine.channel_profile()
    al = 1.0 + 0.01 * np.arange(-4, 5)
    Re = np.full(al.size, 10000.0)
    g = np.full(al.size, complex(0.23753, 0.00374))
    for integ in (osb.INT_CK45, osb.INT_RK4):
        opts = osb.shoot_defaults(osb.FLOW_CONTE)
        opts.integrator = integ
        r = engine.shoot_temporal(opts, ch, al, Re, g)
        ref = o.shoot_temporal_batch(ob.FLOW_CONTE, integ, 2000, 10.0, -1.0, 1.0, None, al, Re, g)
        _bits_equal(r, ref, f"conte integ={integ}")
    ch.free()
    p = engine.solve_bl(osb.FLOW_ORRSOM, osb.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    prof, _, _ = o.solve_bl(ob.FLOW_ORRSOM, ob.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    al = (0.179 + 0.002 * np.arange(-6, 7)) * S2
    Re = np.full(al.size, 580.0 * S2)
    g = np.full(al.size, GUESS)
    for integ in (osb.INT_RK4, osb.INT_CK45):
        opts = osb.shoot_defaults(osb.FLOW_ORRSOM)
        opts.integrator, opts.nstep = integ, 1000
        r = engine.shoot_temporal(opts, p, al, Re, g)
        ref = o.shoot_temporal_batch(ob.FLOW_ORRSOM, integ, 1000, 10.0, 0.0, 17.0, prof, al, Re, g)
        _bits_equal(r, ref, f"orrsom integ={integ}")
    p.free()


@pytest.mark.gpu
def test_strict_build_is_bit_identical_on_the_orrspace_sweep(engine, osb, oracle_math, strict):
    """(4) sweep.inp: 50 continuation steps in omega (orrspace.f:186-195), analytic inner product, cosine
    test, c = omega/alpha in the RHS -- every fort.17 row bit for bit, wandering points included."""
    o = oracle_math
    o.set_math("portable")
    ymin, ymax = 0.0, 20.0 * FACT
    p = engine.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
    prof, _, _ = o.solve_bl(ob.FLOW_ORRSPACE, ob.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
    opts = osb.shoot_defaults(osb.FLOW_ORRSPACE)
    opts.nstep, opts.ymin, opts.ymax, opts.testalpha = 1000, ymin, ymax, 0.9
    Re = np.array([1000.0, 990.0]) * FACT
    om = np.full(2, complex(0.042, 0.0)) * FACT
    dom = np.full(2, 0.002) * FACT
    a0 = np.full(2, complex(0.13809592, 0.0051958399)) * FACT
    r = engine.shoot_spatial_lines(opts, p, Re, om, dom, a0, 50)
    p.free()
    for line in range(2):
        ref = o.shoot_spatial_line(ob.INT_RK4, 1000, 0.9, ymin, ymax, prof, Re[line], om[line], dom[line], a0[line], 50)
        assert np.array_equal(r["alpha"][line], ref["alpha"])
        assert np.array_equal(r["passes"][line], ref["passes"])
        assert np.array_equal(r["qend"][line], ref["qend"])
        assert np.array_equal(r["status"][line], ref["status"])


@pytest.mark.gpu
def test_fast_build_agrees_with_strict_build(engine, osb, monkeypatch):
    """(5) the default (fast) kernels against the strict build on the same device: eigenvalues <= 1e-10
    relative wherever both converged; integer outputs may differ only where a decision sits on its threshold."""
    bl = engine.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    A, R, G = grid(48, 48)
    fast = engine.shoot_temporal(opts, bl, A, R, G)
    monkeypatch.setenv("OSB_STRICT", "1")
    slow = engine.shoot_temporal(opts, bl, A, R, G)
    monkeypatch.delenv("OSB_STRICT")
    bl.free()
    ok = (fast["status"] == 0) & (slow["status"] == 0)
    assert ok.mean() > 0.9
    d = rel(fast["c"][ok], slow["c"][ok])
    print(f"fast vs strict: max rel {d.max():.2e}; passes differ at {np.mean(fast['passes'][ok] != slow['passes'][ok]):.1%}, "
          f"qend at {np.mean(fast['qend'][ok] != slow['qend'][ok]):.1%}")
    assert d.max() < 1e-10
    assert np.abs(fast["qend"][ok] - slow["qend"][ok]).max() <= 1


@pytest.mark.gpu
def test_sensitivity_classed_tolerances(engine, osb, oracle):
    """(6) SURVEY 7.3 P2(i).  orrsom and orrspace stop as soon as |err| < 1e-8 (orrsom.f:203, orrspace.f:277):
    the value they REPORT (ctemp, the last value integrated) is still a distance delta = |c_next - ctemp| away
    from the root their own last Muller step points at -- 1e-13 ... 1e-9 depending on how the iteration arrived
    -- and which iterate the loop stops at depends on the phase of err, which is fixed by rounding residues
    (DESIGN.md section 4).  Two arithmetics therefore agree on such a point to delta, not to 1e-10.  Per point:
        sens  = largest movement of the oracle's own answer under +-1 ulp perturbations of the guess
        delta = the oracle's and the GPU's unapplied last update (resid / last two history rows)
        tol   = max(1e-10, 10 sens, 2 (delta_oracle + delta_gpu))
    and the classes are reported.  (The strict build reproduces the oracle bit for bit on the same flows.)"""
    rows = []
    # orrsom decks around the default point, far guesses on the low side (up to 13 passes)
    p = engine.solve_bl(osb.FLOW_ORRSOM, osb.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    prof, _, _ = oracle.solve_bl(ob.FLOW_ORRSOM, ob.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    al = (0.179 + 0.002 * np.arange(-16, 9)) * S2
    Re = np.full(al.size, 580.0 * S2)
    g = np.full(al.size, GUESS)
    opts = osb.shoot_defaults(osb.FLOW_ORRSOM)
    opts.nstep = 1000
    r = engine.shoot_temporal(opts, p, al, Re, g, want_resid=True)
    p.free()
    base = oracle.shoot_temporal_batch(ob.FLOW_ORRSOM, ob.INT_RK4, 1000, 10.0, 0.0, 17.0, prof, al, Re, g)
    assert (base["status"] == 0).all() and (r["status"] == 0).all()
    sens = np.zeros(al.size)
    for dg in (np.nextafter(g.real, 1.0) + 1j * g.imag, np.nextafter(g.real, 0.0) + 1j * g.imag,
               g.real + 1j * np.nextafter(g.imag, 1.0)):
        pert = oracle.shoot_temporal_batch(ob.FLOW_ORRSOM, ob.INT_RK4, 1000, 10.0, 0.0, 17.0, prof, al, Re, dg)
        sens = np.maximum(sens, rel(pert["c"], base["c"]))
    delta = (base["absdc"] + r["resid"][:, 1]) / np.abs(base["c"])
    rows.append(("orrsom", rel(r["c"], base["c"]), sens, delta))
    # orrspace sweep.inp line: the perturbation is applied to the first alpha guess and propagates by continuation
    ymin, ymax = 0.0, 20.0 * FACT
    p = engine.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
    prof, _, _ = oracle.solve_bl(ob.FLOW_ORRSPACE, ob.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
    opts = osb.shoot_defaults(osb.FLOW_ORRSPACE)
    opts.nstep, opts.ymin, opts.ymax, opts.testalpha = 1000, ymin, ymax, 0.9
    a0 = complex(0.13809592, 0.0051958399) * FACT
    rr = engine.shoot_spatial_lines(opts, p, [1000.0 * FACT], [0.042 * FACT], [0.002 * FACT], [a0], 50, hist_cap=50)
    p.free()
    assert (rr["status"] == 0).all()
    line = lambda a: oracle.shoot_spatial_line(ob.INT_RK4, 1000, 0.9, ymin, ymax, prof, 1000.0 * FACT, 0.042 * FACT,
                                               0.002 * FACT, a, 50)
    base = line(a0)
    sens = np.zeros(50)
    for a in (complex(np.nextafter(a0.real, 1.0), a0.imag), complex(np.nextafter(a0.real, 0.0), a0.imag),
              complex(a0.real, np.nextafter(a0.imag, 1.0))):
        sens = np.maximum(sens, rel(line(a)["alpha"], base["alpha"]))
    # the GPU's unapplied update from the last two rows of its per-pass table: |err| / |d err / d alpha|
    dg = np.zeros(50)
    for i in range(50):
        n = rr["passes"][0, i]
        h = rr["history"][0, i]
        a_n, a_m = complex(h[n - 1, 0], h[n - 1, 1]), complex(h[n - 2, 0], h[n - 2, 1])
        e_n, e_m = complex(h[n - 1, 2], h[n - 1, 3]), complex(h[n - 2, 2], h[n - 2, 3])
        dg[i] = abs(e_n) / abs((e_n - e_m) / (a_n - a_m))
    delta = (base["absdc"] + dg) / np.abs(base["alpha"])
    rows.append(("orrspace sweep.inp", rel(rr["alpha"][0], base["alpha"]), sens, delta))
    for name, err, sens, delta in rows:
        tol = np.maximum(1e-10, np.maximum(10.0 * sens, 2.0 * delta))
        cls = {"tol = 1e-10": float(np.mean(tol <= 1e-10)),
               "tol = 10 x 1-ulp sensitivity": float(np.mean((tol > 1e-10) & (10.0 * sens >= 2.0 * delta))),
               "tol = 2 x unapplied last update (errtol 1e-8)": float(np.mean((tol > 1e-10) & (10.0 * sens < 2.0 * delta)))}
        print(f"{name}: max err {err.max():.2e}, max sens {sens.max():.2e}, max delta {delta.max():.2e}, classes {cls}, "
              f"err/tol max {np.max(err / tol):.2f}; points beyond 1e-10: {int(np.sum(err > 1e-10))} of {err.size}")
        assert (err <= tol).all(), (name, np.max(err / tol))
