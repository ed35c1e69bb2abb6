"""Pins the matrix-path oracle (oracle/os_chan_oracle.c).  The reference holds no
golden output for this path ("parity unpinned" by reference tests); the anchors are
chan.inp:4 (the TS mode is sorted index 62 at N=64), Orszag's (1971) eigenvalue for
plane Poiseuille flow at Re=1e4, alpha=1, and the surveyor's independent numpy
restatement."""
import numpy as np
import pytest

import oracle_binding as ob


def test_fd_decks_alpha1(oracle):
    for N, expect in ((64, 0.23625074532 + 0.00057798924j), (128, 0.23717152516 + 0.0028833849j)):
        c = ob.chan(ob.DISC_FD, N)
        r = c.solve_fd(1.0, 1.0e4)
        c.close()
        assert abs(r["omega"] - expect) < 5e-10


def test_fd128_sweep_ends(oracle):
    c = ob.chan(ob.DISC_FD, 128)
    for a, expect in ((0.76, 0.157291368 - 0.00234695529j), (0.96, 0.223441777 + 0.00339899319j),
                      (1.16, 0.291367005 - 0.00643886905j)):
        assert abs(c.solve_fd(a, 1.0e4)["omega"] - expect) < 2e-9
    c.close()


def test_collocation_chan_inp_index_62(oracle):
    """chan.inp line 4 asks for eigenvector 62: two spurious modes (Im ~ 3) sort above the TS mode."""
    c = ob.chan(ob.DISC_CHEB, 64)
    r = c.solve_col(1.0, 1.0e4, which=62)
    c.close()
    s = r["spectrum"]
    assert s.size == 65
    assert s[63].imag > 3.0 and s[64].imag > 3.0 and abs(s[63].real) > 20
    assert abs(s[62] - (0.23752623674 + 0.0037399251j)) < 5e-10
    assert abs(s[62] - (0.23752649 + 0.00373967j)) < 5e-7  # Orszag 1971, converged value
    ev = r["evec"]
    assert abs(np.max(np.abs(ev)) - 1.0) < 1e-12 and abs(ev[0]) < 1e-8 and abs(ev[64]) < 1e-8
    # symmetric (even) mode about the centreline
    assert np.max(np.abs(ev - ev[::-1])) < 1e-6


def test_neutral_point_nc_inp(oracle):
    """nc.inp: N=64, Re=1e4, sfac=.01, alpha=1 -> upper-branch neutral point."""
    c = ob.chan(ob.DISC_CHEB_NUM, 64)
    tr = c.neutral(10000.0, 0.01, 1.0)
    c.close()
    assert 15 <= len(tr) <= 19
    assert abs(tr[-1, 0] - 1.0947279) < 2e-6
    assert abs(tr[-1, 1] - 0.2698801) < 2e-6 and abs(tr[-1, 2]) < 1e-8
    # the coded dw/da (left vector of inv(B)A, oracle D-7) is ~250x smaller than the true one
    assert abs(complex(tr[0, 3], tr[0, 4])) < 0.01


def test_mapped_bl_operator_reproduces_the_shooting_eigenvalue():
    """orrcolbl.f restated (mapped semi-infinite domain, metrics m1..m4) with the Blasius profile of the
    shooting oracle as the mean flow: the most unstable mode of the Chebyshev operator is the
    Tollmien-Schlichting mode of README.md:24-26, c = 0.36412287 + 0.0079597206i."""
    import math

    S2 = math.sqrt(2.0)
    o = ob.load()
    prof, _, _ = o.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, 2000, 0.0, 20.0, 0.5, 0.0)
    yd, U, Us = prof[0][:2001], prof[1][:2001], prof[2][:2001]
    N, lmap = 64, 8.0
    eta = np.cos(np.arange(N + 1) * math.pi / N)
    with np.errstate(divide="ignore"):
        y = lmap * (1 + eta) / (1 - eta)
    inside = np.isfinite(y) & (y <= yd[-1])
    merge = int(np.nonzero(~inside)[0].max())
    yy = np.where(inside, y, yd[-1])
    h = yd[1] - yd[0]
    i = np.clip((yy / h).astype(int), 0, yd.size - 2)
    A, B = (yd[i + 1] - yy) / h, (yy - yd[i]) / h
    Ug = np.where(inside, A * U[i] + B * U[i + 1] + ((A ** 3 - A) * Us[i] + (B ** 3 - B) * Us[i + 1]) * h * h / 6, 1.0)
    c = ob.chan(ob.DISC_CHEB, N, bl=(lmap, Ug, merge))
    al, Re = 0.179 * S2, 580.0 * S2
    r = c.solve_fd(al, Re, want_spectrum=True)
    c.close()
    cph = r["omega"] / al
    assert abs(cph - complex(0.36412287, 0.0079597206)) < 5e-8
    # the rest of the spectrum is stable; the continuous branch sits at c -> 1 - 0.000309i
    spec = r["spectrum"] / al
    assert (spec.imag < cph.imag + 1e-12).all()
    branch = spec[np.abs(spec.real - 1.0) < 1e-3]
    assert branch.size > 5 and np.abs(branch.imag + 0.000309).min() < 2e-6
