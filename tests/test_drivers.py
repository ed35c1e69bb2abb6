"""C++ twin drivers (os-stab_b200/host/osb_drivers): Fortran edit-descriptor and deck
emulation on CPU; on the GPU the drivers must reproduce the reference's stdout records
and fort.N files (README.md:24-26 line included)."""
import math
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_binding as ob

ROOT = Path(__file__).resolve().parent.parent
DRV = ROOT / "os-stab_b200" / "host" / "osb_drivers"
S2 = math.sqrt(2.0)

TEMPORAL_DECK = """#
# temporal deck (layout of the reference's contebl.inp / test.inp: positional lines)
#
nbl = {nbl}              # number of in profile
nstep = {nstep}            # number of integration points
testalpha = 10.0        # Ortho. constraint (0 - 90 deg)
Re = {Re}              # Reynolds number
beta = 0.0              # Falker-Skan parameter
fp2 = 0.50              # Guess for profile slope
alphar = {ar}          # Real part of alpha
alphai = 0.0            # Imag part of alpha
cr = {cr}   # Guess for Re(c)
ci = {ci}   # Guess for Im(c)
Ymin = 0.0              # Minimum Y
Ymax = {ymax}             # Maximum Y
"""

SPATIAL_DECK = """#
# spatial deck (values of the reference's sweep.inp)
#
nbl = 1000              # number of in profile
nstep = 1000            # number of integration points
testalpha = 0.9         # Ortho. constraint (0 - 1)
Re = 1000.0             # Reynolds number
beta = 0.0              # Falker-Skan parameter
fp2 = 0.50              # Guess for profile slope
alphar = 0.13809592E+00 # Real part of alpha
alphai = 0.51958399E-02 # Imag part of alpha
omegar = 0.042          # Guess for Re(omega)
omegai = 0.00           # Guess for Im(omega)
Ymin = 0.0              # Minimum Y
Ymax = 20.0             # Maximum Y
niter = {niter}              # number of iterations
domega = 0.002          # step size
eigfun = {eigfun}              # compute eigenfunctions
"""


def run(prog, stdin, cwd):
    return subprocess.run([str(DRV), prog], input=stdin, capture_output=True, text=True, cwd=cwd, timeout=600)


def test_fortran_edit_descriptors():
    out = subprocess.run([str(DRV), "--selftest-format"], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0] == "E17.8|   0.36412287E+00|  -0.94619873E-04|   0.00000000E+00|"
    assert out[1] == "E20.10|    0.5800000000E+03|    0.1790000000E+00|"
    assert out[2] == "E17.8E3|  0.23752649E+000|"
    assert out[3] == "ES16.8E3| 1.70000000E+001|-1.50000000E-007| 0.00000000E+000|"
    assert out[4] == "1PE20.12E3| 5.000000000000E-001|"
    assert out[5] == "ES17.8|   2.88657986E-15|   1.00000000-101|"  # 3-digit exponent drops the E
    assert out[6] == "E15.8| 0.10000000E+01| 0.10000000+101|"       # rounding carries into the exponent
    assert out[7] == "I5|   21|*****|"


def test_deck_reader(tmp_path):
    deck = tmp_path / "t.inp"
    deck.write_text(TEMPORAL_DECK.format(nbl=400, nstep=4000, Re=580.0, ar=0.179, cr="0.3641312400E+00", ci="0.7952738700E-02", ymax=30.0))
    out = subprocess.run([str(DRV), "--selftest-deck", str(deck)], capture_output=True, text=True, check=True).stdout.split()
    assert out[:2] == ["400", "4000"]
    vals = [float(x) for x in out[2:]]
    assert vals == [10.0, 580.0, 0.0, 0.5, 0.179, 0.0, 0.36413124, 0.0079527387, 0.0, 30.0]
    # (G20.10): no decimal point => scaled by 1e-10; only 20 characters are read (oracle D-8)
    r = subprocess.run([str(DRV), "--selftest-readr", "Re = 580   # no point"], capture_output=True, text=True, check=True)
    assert abs(float(r.stdout) - 580e-10) < 1e-20
    r = subprocess.run([str(DRV), "--selftest-readr", "x = 1.5D+01"], capture_output=True, text=True, check=True)
    assert float(r.stdout) == 15.0


def test_driver_refuses_to_run_without_a_gpu(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run("contebl", "d\n", tmp_path)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_contebl_default_transcript(tmp_path, oracle):
    r = run("contebl", "d\n", tmp_path)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert "          Solve Orr-Sommerfeld for Boundary Layer using shooting" in out
    assert " Read from keyboard, file, default (k,f,d) ==> " in out
    assert " Re =     0.5800000000E+03" in out and " nbl =  1000" in out
    assert " c = (    0.3641312400E+00,     0.7952738700E-02)" in out
    assert " Blasius velocity profile completed..." in out
    # README.md:24-26
    assert "Eigenvalue =    0.36412287E+00    0.79597206E-02     21" in out
    assert "  Eigenvalue iterations =     5" in out
    lines = [ln for ln in out.splitlines() if ln.startswith("    1   0.36413124E+00   0.79527387E-02")]
    assert len(lines) == 1
    # fort.10: y, spline(U), U'' (spline), U, d2U ; fort.11-14: eigenfunction
    prof, _, _ = oracle.solve_bl()
    f10 = np.loadtxt(tmp_path / "fort.10")
    assert f10.shape == (1001, 5)
    assert np.max(np.abs(f10[:, 3] - prof[1][:1001])) < 1e-8 and abs(f10[1000, 0] - 17.0) < 1e-12
    ref = oracle.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof, 0.179 * S2, 580.0 * S2,
                                complex(0.36413124, 0.79527387e-2), eigfun=True)
    yr = ref["y"] / ref["vmax"]
    for i in range(4):
        f = np.loadtxt(tmp_path / f"fort.{11 + i}")
        assert f.shape == (1001, 3) and f[0, 0] == 17.0
        assert np.max(np.abs(f[:, 1] + 1j * f[:, 2] - yr[:, i])) < 2e-8 * max(1.0, np.max(np.abs(yr[:, i])))


@pytest.mark.gpu
def test_contebl_deck_file(tmp_path, oracle):
    (tmp_path / "t.inp").write_text(TEMPORAL_DECK.format(nbl=1000, nstep=1000, Re=581.1, ar=0.1453, cr=0.35, ci=0.01, ymax=17.0))
    r = run("contebl", "f\nt.inp\n", tmp_path)
    assert r.returncode == 0, r.stderr
    prof, _, _ = oracle.solve_bl()
    ref = oracle.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, prof, 0.1453 * S2, 581.1 * S2, complex(0.35, 0.01))
    ln = [x for x in r.stdout.splitlines() if x.startswith("Eigenvalue =")][0].split()
    assert abs(float(ln[2]) - ref["c"].real) < 1e-8 and abs(float(ln[3]) - ref["c"].imag) < 1e-9
    assert int(ln[4]) == ref["qend"]


@pytest.mark.gpu
def test_conte_default_transcript(tmp_path):
    r = run("conte", "d\n", tmp_path)
    assert r.returncode == 0, r.stderr
    # format 40 of shoot.f:330-331: 'Eigenvalue = ',e17.8e3,1x,e17.8e3,2x,i5
    assert "Eigenvalue =   0.23752649E+000   0.37396706E-002    115" in r.stdout
    assert " Channel velocity profile completed..." in r.stdout
    f10 = np.loadtxt(tmp_path / "fort.10")
    assert f10.shape == (4001, 3) and f10[2000, 1] == 1.0


@pytest.mark.gpu
def test_orrsom_deck_on_stdin(tmp_path):
    deck = TEMPORAL_DECK.format(nbl=400, nstep=4000, Re=580.0, ar=0.179, cr="0.3641312400E+00", ci="0.7952738700E-02", ymax=30.0)
    r = run("orrsom", deck, tmp_path)
    assert r.returncode == 0, r.stderr
    assert "Eigenvalue =    0.36407048E+00    0.80033261E-02     39" in r.stdout
    assert "Second pass: computing and writing eigenfunction" in r.stdout
    assert (tmp_path / "fort.15").exists()


@pytest.mark.gpu
def test_orrspace_sweep_fort17(tmp_path, oracle):
    (tmp_path / "sweep.inp").write_text(SPATIAL_DECK.format(niter=50, eigfun=0))
    r = run("orrspace", "sweep.inp\n", tmp_path)
    assert r.returncode == 0, r.stderr
    f17 = np.loadtxt(tmp_path / "fort.17")
    assert f17.shape == (50, 3)
    expect = {0: (0.042, 0.138095915, 0.00519583038), 9: (0.060, 0.183042716, -0.00167594002),
              28: (0.098, 0.275061388, -0.00738817368), 49: (0.140, 0.372260299, 0.00445007249)}
    for i, (w, ar, ai) in expect.items():
        assert abs(f17[i, 0] - w) < 1e-9 and abs(f17[i, 1] - ar) < 2e-8 and abs(f17[i, 2] - ai) < 2e-9
    assert r.stdout.count("Omega = (") == 50
    # per-pass table (orrspace.f:485-488): 402 rows in total for this deck, first row = the deck's guess
    rows = [ln for ln in r.stdout.splitlines() if len(ln) > 80 and ln[:3].strip().isdigit() and "Omega" not in ln]
    assert abs(len(rows) - 402) <= 3
    assert rows[0].split()[0] == "1" and abs(float(rows[0].split()[1]) - 0.13809592) < 1e-7
    assert "displacement thickness to match Mack" in r.stdout


@pytest.mark.gpu
def test_orrspace_space_deck_writes_u_and_v(tmp_path, oracle):
    """space.inp (niter = 1, eigfun = 1): fort.11-14 and the u, v disturbance velocities of orrspace.f:547-578 --
    fort.15 / fort.16 at nine phases of a half period (format 21), fort.18 normalised by the largest |u| --
    against the same formulas applied to the oracle's eigenfunction.  orrspace stops at |err| < 1e-8
    (orrspace.f:277), so alpha itself is only converged to ~1e-9 and the normalised eigenfunctions of two
    arithmetics agree to ~1e-8: asserted at 1e-7 like test_orrspace_space_deck_eigenfunction."""
    deck = SPATIAL_DECK.format(niter=1, eigfun=1).replace("testalpha = 0.9 ", "testalpha = 0.2 ")
    deck = deck.replace("alphar = 0.13809592E+00", "alphar = 0.228047").replace("alphai = 0.51958399E-02", "alphai = -0.00651")
    deck = deck.replace("omegar = 0.042 ", "omegar = 0.08 ").replace("domega = 0.002", "domega = 0.0")
    (tmp_path / "space.inp").write_text(deck)
    r = run("orrspace", "space.inp\n", tmp_path)
    assert r.returncode == 0, r.stderr
    fact = math.sqrt(2.0) / 1.7207876
    ymin, ymax, nstep = 0.0, 20.0 * fact, 1000
    prof, _, _ = oracle.solve_bl(ob.FLOW_ORRSPACE, ob.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
    Re, om, a0 = 1000.0 * fact, complex(0.08, 0.0) * fact, complex(0.228047, -0.00651) * fact
    ref = oracle.shoot_spatial_point(ob.INT_RK4, nstep, 0.2, ymin, ymax, prof, Re, om, a0)
    y, alpha = ref["y"], ref["alpha"]
    t = (ymax - (ymax - ymin) / nstep * np.arange(nstep + 1)) * fact
    f11 = np.loadtxt(tmp_path / "fort.11")
    assert np.max(np.abs(f11[:, 0] - t)) < 1e-7
    phases = np.exp(1j * (-om * np.arange(9) * (3.14159265358979323846 / om.real / 8)))
    u15 = (y[:, 1] / ref["vmax"])[:, None] * phases[None, :]
    v16 = (-1j * alpha * y[:, 0] / ref["vmax"])[:, None] * phases[None, :]
    f15, f16, f18 = (np.loadtxt(tmp_path / f"fort.{k}") for k in (15, 16, 18))
    assert f15.shape == (nstep + 1, 10) and f16.shape == (nstep + 1, 10) and f18.shape == (nstep + 1, 5)
    assert np.max(np.abs(f15[:, 1:] - u15.real)) < 1e-7 * np.abs(u15).max()
    assert np.max(np.abs(f16[:, 1:] - v16.real)) < 1e-7 * np.abs(v16).max()
    uu, vv = y[:, 1], -1j * alpha * y[:, 0]
    umax = uu[np.argmax(np.abs(uu))]
    assert np.max(np.abs(f18[:, 1] + 1j * f18[:, 2] - uu / umax)) < 1e-7
    assert np.max(np.abs(f18[:, 3] + 1j * f18[:, 4] - vv / umax)) < 1e-7
    assert abs(np.abs(f18[:, 1] + 1j * f18[:, 2]).max() - 1.0) < 1e-8
    # every line of fort.15 is format 21: one ES16.8E3 field, 1x, then nine fields each followed by a blank
    line = (tmp_path / "fort.15").read_text().splitlines()[3]
    assert len(line) == 16 + 1 + 9 * 17


@pytest.mark.gpu
def test_orrfdchan_fd64_deck(tmp_path):
    r = run("orrfdchan", "64\n10000\n0.76 1.16 0.02\n0 0 1\neig-64\n0\n", tmp_path)
    assert r.returncode == 0, r.stderr
    # nar = aint((maxar-minar)/incar)+1 in fp64 (orrfdchan.f:89): (1.16-0.76)/0.02 = 19.999999999999996 -> 20
    nar = int((1.16 - 0.76) / 0.02) + 1
    assert nar == 20
    assert "  nar =    20  nai =     1" in r.stdout
    dat = np.loadtxt(tmp_path / "eig-64.dat")
    assert dat.shape == (nar, 4)
    c = ob.chan(ob.DISC_FD, 64)
    ref = np.array([c.solve_fd(a, 1.0e4)["omega"] for a in dat[:, 0]])
    c.close()
    assert np.max(np.abs(dat[:, 2] + 1j * dat[:, 3] - ref)) < 2e-8
    assert (tmp_path / "eig-64.g").exists() and (tmp_path / "eig-64.q").exists()
    g = (tmp_path / "eig-64.g").read_text().splitlines()
    assert g[0] == "   20     1 "


@pytest.mark.gpu
def test_orrncbl_nc_deck(tmp_path):
    r = run("orrncbl", "64\n10000\n.01\n1\n", tmp_path)
    assert r.returncode == 0, r.stderr
    rows = [ln.split() for ln in r.stdout.splitlines() if ln.startswith("  0.1") and len(ln.split()) == 3]
    assert 15 <= len(rows) <= 19
    assert abs(float(rows[-1][0]) - 1.0947279) < 5e-5 and abs(float(rows[-1][2])) < 1e-8


@pytest.mark.gpu
def test_orrcolchan_chan_deck(tmp_path):
    r = run("orrcolchan", "64\n10000\n1 0\n62\n-1\n", tmp_path)
    assert r.returncode == 0, r.stderr
    assert " Eigenvalues for N+1 =   65" in r.stdout
    # the listing: all N+1 = 65 rows (index, w_r, w_i, c_r, c_i), ascending in w_i (orrcolchan.f:505-513)
    rows = [ln.split() for ln in r.stdout.splitlines() if len(ln.split()) == 5 and ln.split()[0].isdigit()]
    assert [int(x[0]) for x in rows] == list(range(65))
    wi = np.array([float(x[2]) for x in rows])
    assert (np.diff(wi) >= 0).all()
    # chan.inp:4 asks for index 62: the Tollmien-Schlichting mode, right below the two spurious Im ~ 3 modes
    row = rows[62]
    assert abs(float(row[1]) - 0.2375262367) < 5e-9 and abs(float(row[2]) - 0.0037399251) < 5e-9
    assert float(rows[63][2]) > 1.0 and float(rows[64][2]) > 1.0
    c = ob.chan(ob.DISC_CHEB, 64)
    ref = c.solve_col(1.0, 1.0e4)["spectrum"]
    c.close()
    mine = np.array([float(x[1]) + 1j * float(x[2]) for x in rows])
    # least stable dozen modes row for row against the oracle's ZGEGV listing (the deep-damped ones are ill-conditioned)
    assert np.abs(mine[50:63] - ref[50:63]).max() < 5e-7
    f11 = np.loadtxt(tmp_path / "fort.11")
    assert f11.shape == (65, 5)
    v = f11[:, 1] + 1j * f11[:, 2]
    assert abs(np.abs(v).max() - 1.0) < 1e-8
    # columns 4-5: the adjoint eigenvector after the bi-orthonormalisation of orrcolchan.f:537-545, and the two
    # zdotc checks the reference prints before / after it (the second is 1)
    c = ob.chan(ob.DISC_CHEB, 64)
    refv = c.solve_col(1.0, 1.0e4, which=62)
    c.close()
    a = f11[:, 3] + 1j * f11[:, 4]
    assert a[0] == 0 and a[-1] == 0
    assert np.max(np.abs(a - refv["avec"])) < 2e-6 * np.abs(refv["avec"]).max()
    assert abs(np.vdot(a, v) - 1.0) < 1e-6   # 8 printed digits
    import re
    dots = re.findall(r"\(([-+0-9.eE]+),([-+0-9.eE]+)\)", r.stdout)   # the first follows the prompt on its line
    assert len(dots) == 2
    assert abs(complex(float(dots[1][0]), float(dots[1][1])) - 1.0) < 1e-12


UC_DECK = TEMPORAL_DECK.rstrip("\n") + """
Uc = {uc}                # reference frame velocity
delta = {delta}           # alpha stencil half width
deltauc = {duc}          # decrement of Uc per step
num = {num}                 # number of Uc steps
omega = 0               # guess is omega (1) or c (0)
saddle = {saddle}              # track the saddle point
"""


def _oracle_saddle(oracle, Re, alpha1, c, uc, delta, duc, num):
    """The saddle loop of conteucbl.f:232-282 on the CPU oracle (CONTE_IC rules + CRK4 + SHIFT_BL)."""
    prof0, _, _ = oracle.solve_bl(ob.FLOW_CONTEBL, ob.INT_RK4, 1000, 0.0, 17.0, 0.5, 0.0)
    ar = [alpha1.real, alpha1.real + delta, alpha1.real - delta]
    ai = [alpha1.imag, alpha1.imag - delta, alpha1.imag - delta]
    rows, alphas = [], complex(10.0, 10.0)
    for k in range(num):
        prof = oracle.profile_shift(prof0, uc)
        resid, it = 10.0, 0
        while abs(resid) > 1e-5 and it < 60:
            it += 1
            ev, alp = [], []
            for i in range(3):
                old = alphas
                a = complex(ar[i], ai[i]) * S2
                alp.append(a / S2)
                r = oracle.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_RK4, 1000, 10.0, 0.0, 17.0, prof, a, Re * S2, c)
                c = r["next"]  # COMMON c keeps the last update
                ev.append(c * alp[i])
            A = -1j * (ev[1] - ev[0]); B = -1j * (ev[2] - ev[0])
            alphas = (A * (alp[2] ** 2 - alp[0] ** 2) - B * (alp[1] ** 2 - alp[0] ** 2)) / (2 * B * (alp[0] - alp[1]) - 2 * A * (alp[0] - alp[2]))
            resid = alphas - old
            ar = [alphas.real, alphas.real + delta, alphas.real - delta]
            ai = [alphas.imag, alphas.imag - delta, alphas.imag - delta]
        rows.append((uc, ev[0], alphas, it))
        uc -= duc
    return rows


@pytest.mark.gpu
def test_conteucbl_default_is_contebl_rk4(tmp_path):
    """conteucbl with its defaults (uc = 0, saddle = 0) is contebl built without USE_RKCK45."""
    r = run("conteucbl", "d\n", tmp_path)
    assert r.returncode == 0, r.stderr
    assert "Eigenvalue =    0.36412271E+00    0.79597135E-02     21" in r.stdout
    assert " Uc =     0.0000000000E+00" in r.stdout and " Velocity Profile completed..." in r.stdout


@pytest.mark.gpu
def test_conteucbl_saddle_tracking(tmp_path, oracle):
    """Saddle point (zero group velocity in the frame moving with u_c) tracked for two frame speeds;
    parity unpinned by the reference (it ships no deck for this), checked against the same loop on the oracle."""
    uc, delta, duc, num = 0.40, 0.002, 0.01, 2
    c0 = complex(0.36412271 - uc, 0.0079597135)
    (tmp_path / "uc.inp").write_text(UC_DECK.format(nbl=1000, nstep=1000, Re=580.0, ar=0.179, cr=c0.real, ci=c0.imag, ymax=17.0,
                                                    uc=uc, delta=delta, duc=duc, num=num, saddle=1))
    r = run("conteucbl", "f\nuc.inp\n", tmp_path)
    assert r.returncode == 0, r.stderr + r.stdout[-2000:]
    rows = [ln.split() for ln in r.stdout.splitlines() if len(ln.split()) == 5 and ln.startswith(" 0.")]
    assert len(rows) == num
    ref = _oracle_saddle(oracle, 580.0, complex(0.179, 0.0), c0, uc, delta, duc, num)
    for row, (u, ev0, alphas, it) in zip(rows, ref):
        assert abs(float(row[0]) - u) < 1e-4
        assert abs(complex(float(row[1]), float(row[2])) - ev0) < 1e-7
        assert abs(complex(float(row[3]), float(row[4])) - alphas) < 1e-6


@pytest.mark.gpu
def test_orrfdchan_fdchan_deck_prints_eigenvector(tmp_path):
    """fdchan.inp (README.md:126): N=128, alpha=1, 'Print eigenvectors' = 1, then which = 1, 0.
    The reference writes the eigenvector of the mode it keeps (normalised by its largest
    component) to fort.11 whatever index is typed (orrfdchan.f:843-871)."""
    r = run("orrfdchan", "128\n10000\n1 1 1\n0 0 1\neig\n1\n1\n0\n", tmp_path)
    assert r.returncode == 0, r.stderr
    assert "  nar =     1  nai =     1" in r.stdout
    assert "w_r                w_i              count" in r.stdout
    listing = [ln.split() for ln in r.stdout.splitlines() if len(ln.split()) == 3 and ln.split()[2].isdigit() and "E" in ln]
    assert len(listing) == 127 and [int(x[2]) for x in listing] == list(range(1, 128))   # the whole spectrum, sorted
    assert r.stdout.count("Eigenvector for which eigenvalue (0 quits) ==>") == 2
    dat = np.loadtxt(tmp_path / "eig.dat").reshape(-1, 4)
    c = ob.chan(ob.DISC_FD, 128)
    ref = c.solve_fd(1.0, 1.0e4, want_evec=True)
    c.close()
    assert abs(dat[0, 2] + 1j * dat[0, 3] - ref["omega"]) < 2e-8   # 8 printed digits
    f11 = np.loadtxt(tmp_path / "fort.11")
    assert f11.shape == (129, 5)
    v = f11[:, 1] + 1j * f11[:, 2]
    assert abs(np.abs(v).max() - 1.0) < 1e-8 and v[0] == 0 and v[-1] == 0
    assert np.max(np.abs(v - ref["evec"])) < 1e-7
    # columns 4-5: ZGEEV's left eigenvector of inv(B)A as orrfdchan.f:865-871 lays it out (unshifted, two zero rows)
    c = ob.chan(ob.DISC_FD, 128)
    refa = c.solve_fd(1.0, 1.0e4, want_adj=True)["avec"]
    c.close()
    a = f11[:, 3] + 1j * f11[:, 4]
    assert a[-1] == 0 and a[-2] == 0 and abs(np.linalg.norm(a) - 1.0) < 1e-7
    assert np.max(np.abs(a - refa)) < 2e-8
    assert " escale = " in r.stdout
    f20 = np.loadtxt(tmp_path / "fort.20").reshape(-1, 2)
    assert f20.shape == (127, 2)
    c20 = f20[:, 0] + 1j * f20[:, 1]
    assert np.abs(c20 - ref["omega"] / 1.0).min() < 1e-8


@pytest.mark.gpu
def test_orrcolbl_and_orrfdbl_twins(tmp_path):
    """orrcolbl / orrfdbl: same prompts and files; the Blasius profile comes from this engine
    instead of the reference's missing profile file.  Collocation at N=64, Lmap=8 must reproduce the
    Tollmien-Schlichting eigenvalue of README.md:24-26 (alpha = 0.179 sqrt 2, Re = 580 sqrt 2)."""
    import math
    import struct

    S2 = math.sqrt(2.0)
    a0 = 0.179 * S2
    deck = f"64\n{580.0 * S2!r}\n8.0\n{a0 - 0.002!r} {a0 + 0.0021!r} 0.002\n0 0 1\nbl\n0\nblasius\n"
    r = run("orrcolbl", deck, tmp_path)
    assert r.returncode == 0, r.stderr
    assert "Solve Orr-Sommerfeld (Collocation)" in r.stdout and "Enter mapping metric ==>" in r.stdout
    assert "Read Mean Profile" in r.stdout and " Merging at i =   22" in r.stdout
    dat = np.loadtxt(tmp_path / "bl.dat")
    assert dat.shape == (3, 6)
    c = (dat[1, 2] + 1j * dat[1, 3]) / dat[1, 0]
    assert abs(c - complex(0.36412287, 0.0079597206)) < 5e-8
    # dw/dalpha (central difference of two extra solves) against the slope across the rows
    w = dat[:, 2] + 1j * dat[:, 3]
    slope = (w[2] - w[0]) / (dat[2, 0] - dat[0, 0])
    dwda = dat[1, 4] + 1j * dat[1, 5]
    assert abs(dwda - slope) < 2e-4 and 0.3 < dwda.real < 0.5      # group velocity of the TS wave
    # unformatted .g: record (nx, ny) as 8-byte integers, then 2 nx ny reals
    g = (tmp_path / "bl.g").read_bytes()
    assert struct.unpack("<i2qi", g[:24]) == (16, 3, 1, 16)
    assert struct.unpack("<i", g[24:28])[0] == 2 * 3 * 8
    q = (tmp_path / "bl.q").read_bytes()
    assert len(q) == (4 + 16 + 4) + (4 + 32 + 4) + (4 + 4 * 3 * 8 + 4)
    r = run("orrfdbl", f"128\n{580.0 * S2!r}\n2.0\n{a0!r} {a0!r} 1\n0 0 1\nfdbl\n0\nblasius\n", tmp_path)
    assert r.returncode == 0, r.stderr
    assert "Solve Orr-Sommerfeld (Finite-difference)" in r.stdout
    d2 = np.loadtxt(tmp_path / "fdbl.dat").reshape(-1, 4)
    c2 = (d2[0, 2] + 1j * d2[0, 3]) / d2[0, 0]
    assert abs(c2 - complex(0.36412287, 0.0079597206)) < 3e-3      # second-order differences on 128 points
