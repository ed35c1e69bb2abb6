"""Pins the CPU oracle (oracle/os_shoot_oracle.c) against every known answer the
reference holds for the shooting path:

  README.md:24-26     contebl default eigenvalue, 8 printed digits
  contebl.f:75-76     old converged value left as the default guess (1e-5 window)
  conte.f:65-66       ditto for the channel
  sweep.inp:15-16     converged spatial alpha embedded in the deck (8 digits)

plus closed-form unit checks of the restated numerics and the surveyor's
independently restated values."""
import math

import numpy as np
import pytest

import oracle_binding as ob

S2 = math.sqrt(2.0)


def fortran_e(x, digits=8):
    """Fortran Ew.d 'E17.8' mantissa/exponent as printed (0.dddddddE+ee)."""
    if x == 0:
        return "0." + "0" * digits + "E+00"
    e = int(math.floor(math.log10(abs(x)))) + 1
    m = x / 10.0 ** e
    s = f"{m:.{digits}f}"
    if s.startswith("1.") or s.startswith("-1."):  # rounding carried
        m /= 10.0
        e += 1
        s = f"{m:.{digits}f}"
    return f"{s}E{e:+03d}"


@pytest.fixture(scope="module")
def bl_ck45(oracle):
    return oracle.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)


def test_blasius_wall_shear(bl_ck45):
    prof, npass, f2p = bl_ck45
    # Blasius f''(0) = 0.4696 (Mack's scaling); secant stops when |1 - f'(ymax)| <= 1e-10
    assert abs(f2p - 0.46959999) < 2e-8
    assert abs(prof[1][1000] - 1.0) <= 1e-10
    assert npass == 6
    assert prof[1][0] == 0.0


def test_spline_reproduces_cubic_with_matching_end_condition(oracle):
    # SPLINE's "cantilever" end condition (lambda = 1) is exact for a quadratic
    nbl = 50
    prof, _, _ = oracle.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, nbl, 0.0, 5.0, 0.5, 0.0)
    y = prof[0][: nbl + 1]
    assert np.allclose(np.diff(y), 0.1)
    # Uspl approximates U'' : compare with a finite difference in the interior
    U = prof[1][: nbl + 1]
    fd = (U[2:] - 2 * U[1:-1] + U[:-2]) / 0.01
    assert np.max(np.abs(prof[2][1:nbl] - fd)) < 5e-3


def test_readme_eigenvalue_contebl_default(oracle, bl_ck45):
    """README.md:24-26:  c = 0.36412287E+00 + i 0.79597206E-02."""
    prof = bl_ck45[0]
    r = oracle.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof,
                              0.179 * S2, 580.0 * S2, complex(0.36413124, 0.79527387e-2))
    assert r["status"] == 0
    assert fortran_e(r["c"].real) == "0.36412287E+00"
    assert fortran_e(r["c"].imag) == "0.79597206E-02"
    assert r["passes"] == 5 and r["qend"] == 21
    # surveyor's independent restatement, 11 digits
    assert abs(r["c"] - complex(0.36412286722, 0.0079597205884)) < 5e-12
    # the default guess in the source is an older converged value (contebl.f:75-76)
    assert abs(r["c"] - complex(0.36413124, 0.0079527387)) < 2e-5
    # pass table of shoot.f:745-747, first row is the guess itself
    assert r["history"][0, 0] == 0.36413124
    assert abs(r["history"][2, 0] - 0.364122854) < 1e-9


def test_contebl_rk4_build_differs_in_7th_digit(oracle):
    prof, _, _ = oracle.solve_bl(ob.FLOW_CONTEBL, ob.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    r = oracle.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_RK4, 1000, 10.0, 0.0, 17.0, prof,
                              0.179 * S2, 580.0 * S2, complex(0.36413124, 0.79527387e-2))
    assert abs(r["c"] - complex(0.36412271287, 0.0079597135036)) < 5e-12


def test_conte_channel_default(oracle):
    """conte.f:61-68 defaults; guess 0.23753+0.00374i is an old converged value."""
    r = oracle.shoot_temporal(ob.FLOW_CONTE, ob.INT_CK45, 2000, 10.0, -1.0, 1.0, None, 1.0, 10000.0,
                              complex(0.23753, 0.00374))
    assert r["status"] == 0
    assert abs(r["c"] - complex(0.23753, 0.00374)) < 1e-5
    assert abs(r["c"] - complex(0.23752648882, 0.0037396706230)) < 5e-12
    assert r["passes"] == 5 and r["qend"] == 115
    # Orszag (1971) TS mode for Re=1e4, alpha=1: 0.23752649 + 0.00373967i
    assert abs(r["c"] - complex(0.23752649, 0.00373967)) < 1e-8


def test_orrsom_decks(oracle):
    prof, _, _ = oracle.solve_bl(ob.FLOW_ORRSOM, ob.INT_RK4, 400, 0.0, 30.0, 0.5, 0.0)
    r = oracle.shoot_temporal(ob.FLOW_ORRSOM, ob.INT_RK4, 4000, 10.0, 0.0, 30.0, prof, 0.179 * S2, 580.0 * S2,
                              complex(0.36413124, 0.0079527387))  # test.inp
    assert abs(r["c"] - complex(0.36407047666, 0.0080033261106)) < 5e-12
    assert (r["passes"], r["qend"]) == (5, 39)
    prof, _, _ = oracle.solve_bl(ob.FLOW_ORRSOM, ob.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    r = oracle.shoot_temporal(ob.FLOW_ORRSOM, ob.INT_RK4, 1000, 10.0, 0.0, 17.0, prof, 0.1453 * S2, 581.1 * S2,
                              complex(0.35, 0.01))  # contebl.inp
    assert abs(r["c"] - complex(0.34978736835, 0.012086101216)) < 5e-12
    assert (r["passes"], r["qend"]) == (7, 18)


def test_orrspace_sweep_deck(oracle):
    """sweep.inp: 50 omega steps with continuation; row 1 must reproduce the deck's
    own converged guess (sweep.inp:15-16) to its 8 digits."""
    fact = math.sqrt(2.0) / 1.7207876  # orrspace.f:160
    ymin, ymax = 0.0 * fact, 20.0 * fact
    prof, _, _ = oracle.solve_bl(ob.FLOW_ORRSPACE, ob.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
    r = oracle.shoot_spatial_line(ob.INT_RK4, 1000, 0.9, ymin, ymax, prof, 1000.0 * fact,
                                  complex(0.042, 0.0) * fact, 0.002 * fact,
                                  complex(0.13809592, 0.0051958399) * fact, 50)
    al = r["alpha"] / fact
    assert fortran_e(al[0].real) == "0.13809592E+00"
    assert abs(al[0].imag - 0.0051958399) < 1e-8  # deck value came from an older build
    assert np.all(r["status"] == 0)
    expect = {0: (0.138095915, 0.00519583038, 4), 1: (0.143181773, 0.00434772463, 5),
              9: (0.183042716, -0.00167594002, 6), 19: (0.231812435, -0.00641750284, 18),
              25: (0.260715718, -0.00743718092, 25), 28: (0.275061388, -0.00738817368, 32),
              49: (0.372260299, 0.00445007249, 6)}
    for i, (ar, ai, npass) in expect.items():
        assert abs(al[i].real - ar) < 2e-9 and abs(al[i].imag - ai) < 2e-10
        assert r["passes"][i] == npass
    assert r["passes"].sum() == 402


def test_eigenfunction_pass_is_normalised_and_satisfies_wall_bc(oracle, bl_ck45):
    prof = bl_ck45[0]
    r = oracle.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof,
                              0.179 * S2, 580.0 * S2, complex(0.36413124, 0.79527387e-2), eigfun=True)
    y = r["y"] / r["vmax"]
    # phi(0) = 0 enforced exactly by B (shoot.f:699-700); phi'(0) = err ~ 1e-14
    assert abs(y[1000, 0]) < 1e-12 and abs(y[1000, 1]) < 1e-10
    # vmax is the phi of largest modulus over k = 1..nstep (oracle D-2)
    assert abs(np.max(np.abs(y[1:, 0])) - 1.0) < 1e-12
    # free stream: decays like exp(-alpha y)
    assert abs(y[0, 0]) < 0.1
