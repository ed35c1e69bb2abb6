"""GPU parity of the channel matrix path (K3/K4, through the C ABI) against the oracle
(which calls the reference's LAPACK routines).  FD: 1e-10 relative on the eigenvalue at
every N; Chebyshev collocation: 1e-10 holds at N=64 only -- the operator is so
ill-scaled that the reference's own LAPACK paths disagree beyond that, so
the bound is conditioning-aware."""
import numpy as np
import pytest

import oracle_binding as ob

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)) / np.abs(np.asarray(b))


@pytest.mark.parametrize("disc", [ob.DISC_FD, ob.DISC_CHEB, ob.DISC_CHEB_NUM])
def test_tables_bit_exact(engine, oracle, disc):
    """Product tables (D1hat on the host, the three O(N^3) products on the device, k ascending, no FMA) against
    the oracle's triple loops: same bits.  (Independence of the two is test_tables_against_multiprecision's job.)"""
    ch = engine.chan(disc, 64)
    t = ch.tables()
    ch.free()
    c = ob.chan(disc, 64)
    assert np.array_equal(t["eta"], c.table(0))
    assert np.array_equal(t["u"], c.table(1))
    assert np.array_equal(t["d2u"], c.table(3))
    assert np.array_equal(t["D2"], c.table(4))
    assert np.array_equal(t["D4"], c.table(5))
    c.close()


@pytest.mark.parametrize("disc,N", [(ob.DISC_FD, 48), (ob.DISC_CHEB, 48), (ob.DISC_FD, 96), (ob.DISC_CHEB, 64)])
def test_tables_against_multiprecision(engine, disc, N):
    """An independent check of the derivative tables (the product and the oracle restate the same short reference
    routines, so comparing those two proves little): FiniteD1 / CHEBYD and the products D2 = D1hat D1,
    D4 = D1hat^3 D1 of MAKE_DERIVATIVES (orrfdchan.f:305-342) evaluated with mpmath at 40 digits from the formulas
    in the reference; the fp64 tables (host-built D1hat, device-side products) must equal them to rounding."""
    import mpmath as mp

    mp.mp.dps = 40
    eta = [mp.cos(mp.mpf(i) * mp.pi / N) for i in range(N + 1)]
    D = mp.zeros(N + 1, N + 1)
    if disc == ob.DISC_FD:  # orrfdchan.f:1183-1215
        D[0, 0] = 1 / (eta[0] - eta[1]); D[0, 1] = -1 / (eta[0] - eta[1])
        D[N, N - 1] = 1 / (eta[N - 1] - eta[N]); D[N, N] = -1 / (eta[N - 1] - eta[N])
        for i in range(1, N):
            D[i, i - 1] = 1 / (eta[i - 1] - eta[i + 1]); D[i, i + 1] = -1 / (eta[i - 1] - eta[i + 1])
    else:  # orrcolchan.f:904-944
        c = [mp.mpf(1)] * (N + 1)
        c[0] = c[N] = mp.mpf(2)
        for j in range(N + 1):
            for k in range(N + 1):
                if j == 0 and k == 0: v = (2 * mp.mpf(N) ** 2 + 1) / 6
                elif j == N and k == N: v = -(2 * mp.mpf(N) ** 2 + 1) / 6
                elif j == k: v = -eta[j] / (2 * (1 - eta[j] ** 2))
                else: v = c[j] * (-1) ** (j + k) / (c[k] * (eta[j] - eta[k]))
                D[j, k] = v
    D1 = D.copy()
    for i in range(N + 1):
        D1[0, i] = 0; D1[N, i] = 0
    D2 = D * D1
    D4 = D * (D * D2)
    tofloat = lambda Mx: np.array([[float(Mx[i, j]) for j in range(N + 1)] for i in range(N + 1)])
    ch = engine.chan(disc, N)
    t = ch.tables()
    ch.free()
    assert np.max(np.abs(t["eta"] - np.array([float(e) for e in eta]))) < 5e-16
    D2f, D4f = tofloat(D2), tofloat(D4)
    assert np.abs(t["D2"] - D2f).max() / np.abs(D2f).max() < 1e-12
    assert np.abs(t["D4"] - D4f).max() / np.abs(D4f).max() < 1e-12
    if disc == ob.DISC_FD:  # the FD operator is banded: exact zeros outside half bandwidths 2 and 4
        i, j = np.indices(D2f.shape)
        assert np.all(t["D2"][np.abs(i - j) > 2] == 0) and np.all(t["D4"][np.abs(i - j) > 4] == 0)


@pytest.mark.parametrize("N,lo,hi,inc", [(64, 0.76, 1.16, 0.02), (128, 0.76, 1.16, 0.01)])
def test_orrfdchan_decks(engine, oracle, N, lo, hi, inc):
    """fd-64.inp / fd-128.inp: Re=1e4, alpha_r sweep, alpha_i = 0, no shift guesses given."""
    nar = int((hi - lo) / inc) + 1  # orrfdchan.f:89
    alpha = lo + np.arange(nar) * inc
    ch = engine.chan(ob.DISC_FD, N)
    r = ch.solve(alpha, 1.0e4, want_evec=True)
    ch.free()
    c = ob.chan(ob.DISC_FD, N)
    ref = np.array([c.solve_fd(a, 1.0e4)["omega"] for a in alpha])
    refv = c.solve_fd(alpha[nar // 2], 1.0e4, want_evec=True)["evec"]
    c.close()
    assert (r["status"] == 0).all()
    e = rel(r["omega"], ref)
    # the reference's rule (largest Im w, |Im w| < 10) sometimes keeps a spurious FD mode with
    # w_r < 0 (N=64: alpha = 0.76 and 1.16); the GPU path must return the same mode.  Those modes
    # are ill-conditioned: ZGEEV on inv(B)A and shift-invert on the pencil agree to ~1e-10 only.
    phys = ref.real > 0
    assert e[phys].max() < 1e-12, e[phys].max()
    assert e.max() < 1e-9, e.max()
    v = r["evec"][nar // 2]
    assert np.max(np.abs(v - refv)) < 1e-8


def test_orrfdchan_large_N(engine, oracle):
    for N in (256, 512):
        alpha = np.array([0.76, 1.0, 1.16])
        ch = engine.chan(ob.DISC_FD, N)
        r = ch.solve(alpha, 1.0e4)
        ch.free()
        c = ob.chan(ob.DISC_FD, N)
        ref = np.array([c.solve_fd(a, 1.0e4)["omega"] for a in alpha])
        c.close()
        assert (r["status"] == 0).all()
        assert rel(r["omega"], ref).max() < 1e-10, (N, rel(r["omega"], ref).max())


@pytest.mark.parametrize("N", [256, 512])
def test_fd_operator_on_the_dense_kernels(engine, oracle, N, monkeypatch):
    """OSB_FD_DENSE=1 sends the FD operator through the dense complex LU (DMMA trailing update) instead of
    the banded kernels: N=256 is the 256-thread two-per-SM kernel, N=512 the one-per-SM kernel (n > ~350,
    config 4's largest size).  Against the oracle's inv(B)A + ZGEEV (orrfdchan.f:792-797) and against the
    banded path, 1e-10 relative."""
    alpha = np.array([0.76, 0.9, 1.0, 1.16])
    ch = engine.chan(ob.DISC_FD, N)
    banded = ch.solve(alpha, 1.0e4)
    monkeypatch.setenv("OSB_FD_DENSE", "1")
    dense = ch.solve(alpha, 1.0e4, want_evec=True)
    dense_given = ch.solve(alpha, 1.0e4, sigma0=banded["omega"] * (1 + 1e-4))
    monkeypatch.delenv("OSB_FD_DENSE")
    bv = ch.solve(alpha, 1.0e4, want_evec=True)
    ch.free()
    c = ob.chan(ob.DISC_FD, N)
    ref = np.array([c.solve_fd(a, 1.0e4)["omega"] for a in alpha])
    c.close()
    assert (dense["status"] == 0).all() and (dense_given["status"] == 0).all()
    assert rel(dense["omega"], ref).max() < 1e-10, (N, rel(dense["omega"], ref).max())
    assert rel(dense_given["omega"], ref).max() < 1e-10
    assert rel(dense["omega"], banded["omega"]).max() < 1e-10
    assert np.max(np.abs(dense["evec"] - bv["evec"])) < 1e-8


def test_complex_alpha_and_given_shift(engine, oracle):
    """alpha_i != 0 (orrfdchan sweeps alpha_i too, orrfdchan.f:116-118) with a user shift."""
    ch = engine.chan(ob.DISC_FD, 64)
    c = ob.chan(ob.DISC_FD, 64)
    alpha = np.array([1.0 + 0.01j, 0.9 - 0.02j])
    ref = np.array([c.solve_fd(a, 1.0e4)["omega"] for a in alpha])
    r = ch.solve(alpha, 1.0e4, sigma0=ref + 1e-3)
    ch.free(); c.close()
    e = rel(r["omega"], ref)
    assert e[ref.real > 0].max() < 1e-12 and e.max() < 1e-9  # ref[0] is the spurious w_r < 0 mode


def test_orrcolchan_chan_inp(engine, oracle):
    """chan.inp: collocation N=64, alpha=1; the mode the deck asks for (sorted index 62)."""
    ch = engine.chan(ob.DISC_CHEB, 64)
    r = ch.solve([1.0], 1.0e4, want_evec=True)
    ch.free()
    c = ob.chan(ob.DISC_CHEB, 64)
    ref = c.solve_col(1.0, 1.0e4, which=62)
    c.close()
    # QZ on the ill-scaled collocation pencil vs shift-invert: 2.5e-10 apart at N=64
    assert rel(r["omega"][0], ref["spectrum"][62]) < 2e-9
    assert abs(r["omega"][0] - (0.23752649 + 0.00373967j)) < 5e-7
    v, vr = r["evec"][0], ref["evec"]
    assert np.max(np.abs(v - vr)) < 1e-6


def test_orrncbl_neutral_point(engine, oracle):
    """nc.inp through the batched Newton search, plus a small Re sweep."""
    ch = engine.chan(ob.DISC_CHEB_NUM, 64)
    r = ch.neutral([10000.0, 9000.0, 12000.0], 1.0, sfac=0.01, want_trace=True)
    ch.free()
    c = ob.chan(ob.DISC_CHEB_NUM, 64)
    for k, Re in enumerate((10000.0, 9000.0, 12000.0)):
        tr = c.neutral(Re, 0.01, 1.0)
        assert r["status"][k] == 0
        assert abs(r["steps"][k] - len(tr)) <= 1
        # the converged alpha is path dependent only within tol/|Im dw/da| (oracle D-7)
        assert abs(r["alpha"][k] - tr[-1, 0]) < 5e-5
        assert abs(r["omega"][k].real - tr[-1, 1]) < 2e-5
        assert abs(r["omega"][k].imag) <= 1e-8
        # first Newton step: same eigenvalue, same (bug-compatible) dw/da
        t0 = r["trace"][k, 0]
        assert abs(complex(t0[1], t0[2]) - complex(tr[0, 1], tr[0, 2])) < 1e-8
        assert abs(complex(t0[3], t0[4]) - complex(tr[0, 3], tr[0, 4])) < 1e-6 * abs(complex(tr[0, 3], tr[0, 4])) + 1e-9
    c.close()


@pytest.mark.parametrize("N", [64, 128, 256, 512])
def test_collocation_against_quad_arbiter(engine, oracle, N):
    """for Chebyshev collocation the reference's LAPACK routes are themselves
    inaccurate (ill-scaled pencil), so the fp64 oracle cannot arbitrate 1e-10.  The quad-precision
    arbiter solves the same discrete pencil to ~1e-25; the GPU path must be at least as close to
    it as the reference's route, and within 1e-10 where fp64 allows (N = 64).  N = 512 (config 4's
    largest size) runs the one-CTA-per-SM dense kernel (n > ~350)."""
    ch = engine.chan(ob.DISC_CHEB, N)
    r = ch.solve([1.0], 1.0e4)
    ch.free()
    c = ob.chan(ob.DISC_CHEB, N)
    spec = c.solve_col(1.0, 1.0e4)["spectrum"]
    w_ref = spec[np.argmin(np.abs(spec - (0.2375 + 0.0037j)))]
    w_quad = c.arbiter(1.0, 1.0e4, r["omega"][0])
    c.close()
    e_gpu = abs(r["omega"][0] - w_quad) / abs(w_quad)
    e_ref = abs(w_ref - w_quad) / abs(w_quad)
    print(f"N={N}: gpu err {e_gpu:.2e}, reference-route (ZGEGV) err {e_ref:.2e}")
    assert r["status"][0] == 0
    assert e_gpu <= max(e_ref, 1e-12)
    if N == 64:
        assert e_gpu < 1e-10
    # the discrete eigenvalue converges to the shooting value 0.23752648882+0.0037396706230i; at N = 512 the
    # fp64 rounding of the tables themselves (D4 ~ N^8 eps) moves the exact eigenvalue of the pencil by 6e-9
    if N >= 128:
        assert abs(w_quad - complex(0.23752648882, 0.0037396706230)) < (1e-9 if N <= 256 else 5e-8)


@pytest.mark.parametrize("N", [64, 100, 512])
def test_banded_warp_kernel_matches_thread_kernel(engine, N, monkeypatch):
    """The FD operator has two banded-LU kernels: a warp per instance (small sweeps) and a thread
    per instance (GPU-filling sweeps).  Same elimination (same pivots, L, U); the substitutions
    differ in rounding only, so eigenvalues agree to ~1e-13 and eigenvectors to 1e-9."""
    alpha = np.concatenate([np.linspace(0.76, 1.16, 37), [1.0 + 0.01j, 0.9 - 0.02j]])
    ch = engine.chan(ob.DISC_FD, N)
    monkeypatch.delenv("OSB_FD_THREAD", raising=False)
    first = ch.solve(alpha, 1.0e4)  # scan included
    sig = first["omega"]
    a = ch.solve(alpha, 1.0e4, sigma0=sig, want_evec=True)
    monkeypatch.setenv("OSB_FD_THREAD", "1")
    b = ch.solve(alpha, 1.0e4, sigma0=sig, want_evec=True)
    second = ch.solve(alpha, 1.0e4)
    monkeypatch.delenv("OSB_FD_THREAD")
    ch.free()
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    phys = a["omega"].real > 0
    # rounding of the two substitution orders is amplified by the conditioning of the FD operator (~N^4)
    assert rel(a["omega"], b["omega"])[phys].max() < (1e-12 if N <= 128 else 2e-11)
    # spurious FD modes (w_r < 0) are ill-conditioned: determined to ~1e-10 at N=64, ~1e-8 at N=512
    assert rel(a["omega"], b["omega"]).max() < (1e-10 if N <= 128 else 1e-6)
    assert np.max(np.abs(a["evec"] - b["evec"])[phys]) < 1e-8
    # the scan picks the same mode through either kernel (spurious FD modes: ill-conditioned, 1e-9)
    assert rel(first["omega"], second["omega"]).max() < (1e-9 if N <= 128 else 1e-6)


@pytest.mark.parametrize("disc,N,lmap", [(ob.DISC_CHEB, 64, 8.0), (ob.DISC_CHEB, 96, 4.0), (ob.DISC_FD, 128, 2.0)])
def test_mapped_boundary_layer_operator(engine, osb, oracle, disc, N, lmap):
    """orrcolbl / orrfdbl operator with this engine's Blasius profile as the mean flow,
    against the oracle's restatement (same grid, metrics, MAKE_MATRIX; ZGEEV on inv(B)A, most unstable mode)
    -- and, for collocation, against the shooting eigenvalue of the same flow (the two paths tie together)."""
    import math

    S2 = math.sqrt(2.0)
    prof = engine.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 2000, 0.0, 20.0, 0.5, 0.0)
    ch = engine.chan_bl(disc, N, lmap, prof=prof)
    alpha = np.array([0.16, 0.179, 0.2, 0.179 + 0.002j]) * S2
    Re = 580.0 * S2
    r = ch.solve(alpha, Re, want_evec=True)
    sp = ch.spectrum(alpha[1:2], Re)
    oc = ob.chan(disc, N, bl=(lmap, ch.U, ch.merge))
    ref = [oc.solve_fd(a, Re, want_evec=True) for a in alpha]
    # the host tables fold the metrics into D2', D4': compare them with the oracle's unfolded ones
    t = ch.tables()
    oD2 = oc.table(4)
    assert np.array_equal(t["eta"], oc.table(0)) and np.array_equal(t["u"], oc.table(1))
    oc.close()
    ch.free()
    assert (r["status"] == 0).all()
    w_ref = np.array([x["omega"] for x in ref])
    # orrcolbl.f:771-779 keeps the largest Im w of the FULL spectrum: the device spectrum agrees with the scan
    w_all = sp["spectrum"][0]
    assert sp["status"][0] == 0 and abs(w_all[np.argmax(w_all.imag)] - r["omega"][1]) < 1e-8 * abs(r["omega"][1])
    # parity with the LAPACK route on this ill-scaled mapped operator: 1e-9 (collocation), 1e-10 (FD)
    e = rel(r["omega"], w_ref)
    assert e.max() < (2e-9 if disc == ob.DISC_CHEB else 1e-10), e
    v = r["evec"][1]
    assert np.max(np.abs(v - ref[1]["evec"])) < 1e-6
    if disc == ob.DISC_CHEB:
        sh = engine.shoot_temporal(osb.shoot_defaults(osb.FLOW_CONTEBL),
                                   engine.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0),
                                   alpha[:3].real, np.full(3, Re), r["omega"][:3] / alpha[:3])
        assert (sh["status"] == 0).all()
        assert np.abs(r["omega"][:3] / alpha[:3] - sh["c"]).max() < 2e-7
    prof.free()
    assert oD2.shape == (N + 1, N + 1)


@pytest.mark.parametrize("n", [4001, 6001])
def test_banded_kernels_large_batch(engine, monkeypatch, n):
    """More than 3552 instances select the <8-row chunk, 3 CTAs/SM> build of the nine-lane kernel, more than
    5328 the <8, 4> build (128 registers); both must agree with the one-thread kernel exactly like the
    low-latency builds do (ragged tail included)."""
    alpha = np.linspace(0.76, 1.16, n)
    ch = engine.chan(ob.DISC_FD, 64)
    coarse = ch.solve(np.linspace(0.76, 1.16, 9), 1.0e4)
    good = coarse["omega"].real > 0
    xs = np.linspace(0.76, 1.16, 9)[good]
    sig = np.interp(alpha, xs, coarse["omega"].real[good]) + 1j * np.interp(alpha, xs, coarse["omega"].imag[good])
    monkeypatch.delenv("OSB_FD_THREAD", raising=False)
    a = ch.solve(alpha, 1.0e4, sigma0=sig)
    monkeypatch.setenv("OSB_FD_THREAD", "1")
    b = ch.solve(alpha, 1.0e4, sigma0=sig)
    monkeypatch.delenv("OSB_FD_THREAD")
    ch.free()
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    assert rel(a["omega"], b["omega"]).max() < 1e-12
    assert np.abs(a["iters"].astype(int) - b["iters"].astype(int)).max() <= 1


def test_orrfdchan_maximum_size(engine, oracle):
    """N = 1024, the largest grid the C ABI accepts (orrcolchan.f:22 idim = 1024; orrfdchan itself stops at
    512): banded path against the oracle's ZGEEV on the 1023 x 1023 operator."""
    N = 1024
    alpha = np.array([0.9, 1.0])
    ch = engine.chan(ob.DISC_FD, N)
    r = ch.solve(alpha, 1.0e4)
    ch.free()
    c = ob.chan(ob.DISC_FD, N)
    ref = np.array([c.solve_fd(a, 1.0e4)["omega"] for a in alpha])
    c.close()
    assert (r["status"] == 0).all()
    assert rel(r["omega"], ref).max() < 1e-9, rel(r["omega"], ref)


def _match(a, b):
    """Greedy nearest-neighbour matching of two spectra; returns relative distances (per entry of a)."""
    b = np.array(b)
    used = np.zeros(b.size, bool)
    out = np.zeros(a.size)
    for i, z in enumerate(a):
        d = np.where(used, np.inf, np.abs(b - z))
        j = int(np.argmin(d))
        used[j] = True
        out[i] = d[j] / max(abs(b[j]), 1e-300)
    return out


@pytest.mark.parametrize("disc,N", [(ob.DISC_FD, 64), (ob.DISC_CHEB, 64), (ob.DISC_FD, 128), (ob.DISC_FD, 300)])
def test_full_spectrum_on_device(engine, oracle, disc, N):
    """osb_chan_spectrum: inv(B)A + Hessenberg + shifted QR on the GPU against the oracle's ZGEEV spectrum
    (orrfdchan.f:792-811), and the reference's selection rule applied to the device spectrum against the
    mode osb_chan_eig isolates without it."""
    alpha = np.array([0.76, 1.0, 1.16, 1.0 + 0.02j]) if N <= 128 else np.array([1.0])
    ch = engine.chan(disc, N)
    sp = ch.spectrum(alpha, 1.0e4)
    kept = ch.solve(alpha, 1.0e4)
    ch.free()
    assert (sp["status"] == 0).all()
    c = ob.chan(disc, N)
    for k, a in enumerate(alpha):
        ref = c.solve_fd(a, 1.0e4, want_spectrum=True)
        w = sp["spectrum"][k]
        assert w.shape == (N - 1,) and np.isfinite(w).all()
        assert (np.diff(w.imag) >= 0).all()                       # sorted like the reference's listing
        e = _match(w, ref["spectrum"])
        # The spectrum of this non-normal operator is sensitive: a 1e-16 relative perturbation of inv(B)A
        # moves ZGEEV's own eigenvalues by up to 4e-6 (FD, N=200) / 4e-4 (N=300), median 1e-8.  Two QR
        # implementations can only agree to that level; the kept (least stable) mode is well conditioned.
        loose = disc != ob.DISC_FD or N > 128
        assert np.median(e) < (1e-5 if loose else 1e-9), np.median(e)
        # (the worst eigenvalues sit at the junction of the three branches, where the 1e-16-pseudospectrum is ~1e-3 wide:
        # any change of rounding -- e.g. the order of operations of the triangular solves that form inv(B)A -- moves the
        # maximum between 2e-3 and 6e-3 at N = 300 while the median and the well-conditioned modes do not move)
        assert e.max() < (2e-2 if loose else 1e-5), e.max()
        assert (e < 1e-3).mean() > 0.97, (e < 1e-3).mean()
        # the reference's rule on the device spectrum (orrfdchan.f:815-823) picks the mode osb_chan_eig returns
        ok = (np.abs(w.imag) < 10.0) & (w.real.astype(int) != 999)
        pick = w[ok][np.argmax(w[ok].imag)]
        if disc == ob.DISC_FD:
            assert abs(pick - kept["omega"][k]) < 1e-6 * abs(pick), (a, pick, kept["omega"][k])
            assert abs(pick - ref["omega"]) < 1e-6 * abs(pick)
    c.close()


def test_dense_path_size_limit_is_reported(engine, osb):
    """Chebyshev at N = 800 does not fit the dense kernels' shared-memory panel: a clear error, no crash
    (collocation at such N is numerically meaningless in fp64 anyway); N = 640 still runs."""
    ch = engine.chan(ob.DISC_CHEB, 800)
    with pytest.raises(osb.OsbError, match="limit N = "):
        ch.solve([1.0], 1.0e4, sigma0=[0.2375 + 0.0037j])
    with pytest.raises(osb.OsbError, match="limit N = "):
        ch.spectrum([1.0], 1.0e4)
    ch.free()
    ch = engine.chan(ob.DISC_CHEB, 640)
    r = ch.solve([1.0], 1.0e4, sigma0=[0.2375 + 0.0037j])
    ch.free()
    assert np.isfinite(r["omega"][0])


@pytest.mark.parametrize("disc", [ob.DISC_FD, ob.DISC_CHEB])
def test_spectrum_fallback_when_the_scan_misses(engine, disc, monkeypatch):
    """If no scanned shift settles on an acceptable mode, osb_chan_eig takes its shift from the device
    spectrum under the same selection rule: same answers as the scan on the reference's deck points."""
    alpha = np.array([0.76, 0.9, 1.0, 1.16])
    ch = engine.chan(disc, 64)
    monkeypatch.delenv("OSB_CHAN_SCAN_MISS", raising=False)
    a = ch.solve(alpha, 1.0e4)
    monkeypatch.setenv("OSB_CHAN_SCAN_MISS", "1")
    b = ch.solve(alpha, 1.0e4)
    monkeypatch.delenv("OSB_CHAN_SCAN_MISS")
    ch.free()
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    assert rel(a["omega"], b["omega"]).max() < 1e-10


@pytest.mark.parametrize("npts", [1, 2, 4, 5, 13, 25])
def test_banded_nine_lane_kernel_ragged_groups(engine, npts, monkeypatch):
    """Three instances share a warp and twelve a CTA in the nine-lane kernel: sweeps that leave groups or
    warps partly empty must give every instance the result it gets alone."""
    alpha = np.linspace(0.8, 1.1, npts) + 0.003j * (np.arange(npts) % 3)
    ch = engine.chan(ob.DISC_FD, 96)
    monkeypatch.delenv("OSB_FD_THREAD", raising=False)
    batch = ch.solve(alpha, 1.0e4, want_evec=True)
    singles = [ch.solve(alpha[k:k + 1], 1.0e4, sigma0=batch["omega"][k:k + 1], want_evec=True) for k in range(npts)]
    again = ch.solve(alpha, 1.0e4, sigma0=batch["omega"], want_evec=True)
    ch.free()
    assert (batch["status"] == 0).all() and (again["status"] == 0).all()
    for k in range(npts):
        # same kernel, same arithmetic per instance: bit-identical whatever the neighbours in the warp are
        assert singles[k]["omega"][0] == again["omega"][k], k
        assert np.array_equal(singles[k]["evec"][0], again["evec"][k]), k
    assert rel(again["omega"], batch["omega"]).max() < 1e-11


@pytest.mark.parametrize("N", [8, 9, 12, 17])
def test_smallest_grids(engine, oracle, N):
    """The smallest grids the C ABI accepts (interior size 7..16 < one window / one chunk of the banded
    kernels, < one panel of the dense ones): given the oracle's eigenvalue as the shift, every path must
    converge onto it."""
    alpha = np.array([0.9, 1.0])
    for disc in (ob.DISC_FD, ob.DISC_CHEB):
        c = ob.chan(disc, N)
        ref = np.array([c.solve_fd(a, 1.0e3)["omega"] for a in alpha])
        c.close()
        ch = engine.chan(disc, N)
        r = ch.solve(alpha, 1.0e3, sigma0=ref * (1 + 1e-4))
        sp = ch.spectrum(alpha, 1.0e3)
        ch.free()
        assert (r["status"] == 0).all(), (disc, N, r["status"])
        assert rel(r["omega"], ref).max() < 1e-10, (disc, N, rel(r["omega"], ref))
        assert (sp["status"] == 0).all() and sp["spectrum"].shape == (2, N - 1)
        for k in range(2):
            assert np.abs(sp["spectrum"][k] - ref[k]).min() < 1e-8 * abs(ref[k])


def test_banded_batches_are_chunked(engine, monkeypatch):
    """The banded path walks large batches in chunks of bounded device scratch: same bits as one batch."""
    alpha = np.linspace(0.8, 1.1, 50)
    ch = engine.chan(ob.DISC_FD, 64)
    first = ch.solve(alpha, 1.0e4)
    whole = ch.solve(alpha, 1.0e4, sigma0=first["omega"], want_evec=True)
    monkeypatch.setenv("OSB_CHAN_CHUNK", "12")
    parts = ch.solve(alpha, 1.0e4, sigma0=first["omega"], want_evec=True)
    monkeypatch.delenv("OSB_CHAN_CHUNK")
    ch.free()
    for k in ("omega", "evec", "iters", "status"):
        assert np.array_equal(whole[k], parts[k]), k


@pytest.mark.parametrize("disc,N,route", [(ob.DISC_FD, 128, "banded"), (ob.DISC_FD, 128, "thread"), (ob.DISC_FD, 96, "dense"),
                                          (ob.DISC_CHEB, 64, "warp"), (ob.DISC_FD, 200, "dense"), (ob.DISC_CHEB, 96, "dense")])
def test_early_stop_of_the_inverse_iteration(engine, disc, N, route, monkeypatch):
    """The inverse iteration may stop one solve early when two agreeing contraction ratios predict a remaining
    error below the tolerance (conv_push).  With the extrapolation switched off (OSB_NO_EARLY_STOP=1) every
    point iterates until the MEASURED change is below 1e-14 |omega|; the early-stopped results must agree with
    those to the tolerance the status word certifies (1e-12; the collocation pencil at N = 96 is itself
    conditioned to ~1e-11), and must not take more solves."""
    if route == "thread":
        monkeypatch.setenv("OSB_FD_THREAD", "1")
    if route == "dense" and disc == ob.DISC_FD:
        monkeypatch.setenv("OSB_FD_DENSE", "1")
    alpha = np.linspace(0.76, 1.16, 64)
    ch = engine.chan(disc, N)
    seed = ch.solve(alpha[::9], 1.0e4)["omega"]
    sig = np.interp(alpha, alpha[::9], seed.real) + 1j * np.interp(alpha, alpha[::9], seed.imag)
    fast = ch.solve(alpha, 1.0e4, sigma0=sig)
    monkeypatch.setenv("OSB_NO_EARLY_STOP", "1")
    full = ch.solve(alpha, 1.0e4, sigma0=sig)
    ch.free()
    assert (fast["status"] == 0).all() and (full["status"] == 0).all()
    d = rel(fast["omega"], full["omega"])
    print(f"{route} N={N}: early-stop vs full max rel diff {d.max():.2e}; solves {fast['iters'].mean():.2f} vs {full['iters'].mean():.2f}")
    assert d.max() <= (1e-12 if disc == ob.DISC_FD or N <= 64 else 2e-11), d.max()
    assert fast["iters"].sum() <= full["iters"].sum()


@pytest.mark.parametrize("N,env", [(20, {"OSB_CHAN_NO_WARP": "1"}), (33, {"OSB_CHAN_NO_WARP": "1"}), (64, {"OSB_CHAN_NO_WARP": "1"}),
                                   (150, {"OSB_CHAN_NO_SMALL": "1"}), (180, {"OSB_CHAN_NO_SMALL": "1"}), (200, {"OSB_CHAN_BIG_MIN": "100000"}),
                                   (361, {"OSB_CHAN_BIG": "0"}), (361, {"OSB_CHAN_BIG": "1"})])
def test_dense_kernel_shapes_agree(engine, N, env, monkeypatch):
    """The dense path picks a kernel by size (warp per instance for n <= 64, 128- and 256-thread CTAs, two CTAs per SM with
    the two-level LU from n = 351).  Forcing the neighbouring shape at the same N -- CTA kernels on warp-sized operators
    (n = 19, 32, 63: fewer rows than one 16-column block step, exactly two, a ragged last block), the 256-thread kernel on
    small ones, the single-level LU where the two-level one is the default, the single-level and the one-CTA two-level LU
    on a large one -- must reproduce the
    default shape's eigenvalues: they share pivots and row arithmetic and differ in summation grouping at most."""
    alpha = np.linspace(0.8, 1.12, 25)
    ch = engine.chan(ob.DISC_CHEB, N)
    # shifts 1e-5 away from the eigenvalue (a first solve without shifts finds it): both shapes then converge in a few
    # solves, far from the stall of a slowly contracting iteration at the rounding floor of a large collocation pencil
    sig = ch.solve(alpha, 1.0e4)["omega"] * (1.0 + 1.0e-5)
    base = ch.solve(alpha, 1.0e4, sigma0=sig)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    other = ch.solve(alpha, 1.0e4, sigma0=sig)
    ch.free()
    assert (base["status"] == 0).all() and (other["status"] == 0).all()
    d = rel(base["omega"], other["omega"])
    print(f"N={N} {env}: max rel diff between kernel shapes {d.max():.2e}")
    # Same kernel arithmetic, different thread count (OSB_CHAN_NO_SMALL): identical pivots and row operations, 1e-15.
    # Different LU shapes group the sums of the trailing update differently (warp / single-level k = 16 / two-level
    # k = 32 or 64): the difference is rounding amplified by the collocation pencil, whose conditioning grows like N^4
    # -- 1e-12 at N = 64, 3e-10 at N = 200, 1e-9 at N = 361, where single points next to a near-degenerate pair reach
    # 1e-6 (the reference's own ZGEGV is no better there).
    if "OSB_CHAN_NO_SMALL" in env:
        assert d.max() < 1e-13, d.max()
    else:
        floor = 5e-12 * max(1.0, N / 64.0) ** 4
        assert np.median(d) < floor and d.max() < (floor if N <= 200 else 5e-6), (np.median(d), d.max(), floor)


def _pencil(t, alpha, Re):
    """A and B of the interior pencil from the tables osb_chan_tables returns (channel metrics)."""
    n = t["eta"].size - 2
    D2, D4 = t["D2"][1:-1, 1:-1], t["D4"][1:-1, 1:-1]
    u, d2u = t["u"][1:-1], t["d2u"][1:-1]
    a2 = alpha * alpha
    A = D4 + np.diag(-2.0 * a2 - 1j * alpha * Re * u) @ D2 + np.diag(a2 * a2 + 1j * alpha ** 3 * Re * u + 1j * alpha * Re * d2u)
    B = -1j * Re * D2 + 1j * a2 * Re * np.eye(n)
    return A, B


@pytest.mark.parametrize("disc,N,route", [(ob.DISC_CHEB, 64, "warp"), (ob.DISC_CHEB, 128, "cta"), (ob.DISC_FD, 64, "banded"),
                                          (ob.DISC_FD, 128, "banded"), (ob.DISC_FD, 128, "dense"), (ob.DISC_FD, 512, "banded"),
                                          (ob.DISC_FD, 400, "dense")])
def test_adjoint_eigenvector(engine, osb, oracle, disc, N, route, monkeypatch):
    """The left eigenvector the reference asks LAPACK for (JOBVL = 'V', orrfdchan.f:796, orrcolchan.f:461) in both
    forms of the ABI, on every kernel that computes it (warp, CTA, one-thread banded): against the oracle's
    ZGEEV / ZGEGV vectors where those are well conditioned, and -- at every size -- by the defining properties
    u^H (A - w B) = 0, u^H evec = 1, |B^H u| = 1 with its largest component real and positive."""
    if route == "dense" and disc == ob.DISC_FD:
        monkeypatch.setenv("OSB_FD_DENSE", "1")
    alpha = np.array([0.9, 1.0 + 0.01j, 1.1])
    ch = engine.chan(disc, N)
    t = ch.tables()
    rp = ch.solve(alpha, 1.0e4, want_evec=True, want_adj=osb.ADJ_PENCIL)
    ri = ch.solve(alpha, 1.0e4, sigma0=rp["omega"] * (1 + 1e-5), want_adj=osb.ADJ_INVBA)
    ch.free()
    assert (rp["status"] == 0).all() and (ri["status"] == 0).all()
    for k, al in enumerate(alpha):
        A, B = _pencil(t, al, 1.0e4)
        w, x, u, vl = rp["omega"][k], rp["evec"][k][1:-1], rp["avec"][k][1:-1], ri["avec"][k][1:-1]
        assert rp["avec"][k][0] == 0 and rp["avec"][k][-1] == 0
        scale = np.linalg.norm(A, 2) * np.linalg.norm(u)
        assert np.linalg.norm(u.conj() @ (A - w * B)) / scale < 1e-9      # a left eigenvector of the pencil
        assert np.linalg.norm((A - w * B) @ x) / (np.linalg.norm(A, 2) * np.linalg.norm(x)) < 1e-9
        assert abs(np.vdot(u, x) - 1.0) < 1e-10                            # bi-orthonormalised (orrcolchan.f:545)
        assert abs(np.linalg.norm(vl) - 1.0) < 1e-12                      # ZGEEV: unit 2-norm ...
        kmax = np.argmax(np.abs(vl))
        assert abs(vl[kmax].imag) < 1e-14 and vl[kmax].real > 0            # ... largest component real
        bu = B.conj().T @ u
        # (numpy's own B^H u carries cond(B) eps: the FD operator's entries grow like N^4)
        assert np.linalg.norm(bu / np.linalg.norm(bu) * np.exp(-1j * np.angle(bu[kmax])) - vl) < (1e-8 if N <= 128 else 1e-6)
    c = ob.chan(disc, N)
    if disc == ob.DISC_FD and N <= 128:
        for k, al in enumerate(alpha):
            ref = c.solve_fd(al, 1.0e4, want_adj=True)
            if abs(ref["omega"] - ri["omega"][k]) < 1e-8 * abs(ref["omega"]):   # the same mode (not a spurious w_r < 0 one)
                assert np.max(np.abs(ri["avec"][k][1:-1] - ref["avec"][: N - 1])) < 1e-8
    if disc == ob.DISC_CHEB and N == 64:
        ref = c.solve_col(1.0, 1.0e4, which=62)
        one = engine.chan(disc, N)
        r1 = one.solve([1.0], 1.0e4, want_evec=True, want_adj=osb.ADJ_PENCIL)
        one.free()
        assert np.max(np.abs(r1["avec"][0] - ref["avec"])) < 2e-6 * np.abs(ref["avec"]).max()   # QZ's own accuracy here (H7)
    c.close()


def test_mode_choice_is_checked_against_the_full_spectrum(engine, osb, oracle, monkeypatch):
    """The reference keeps the largest Im(w) of the WHOLE spectrum (orrfdchan.f:815-823).  For alpha < 0.5 the FD
    operator's spurious branch sits at w_r ~ -2/alpha, outside the windows the shift scan covers (N = 32,
    Re = 1e4: w = -9.867 - 0.00028i at alpha = 0.2, -4.736 - 0.00115i at 0.4): the scan alone would report a
    different mode as converged.  osb_chan_eig cross-checks the scan against the device spectrum and must
    return the reference's mode; sweeps too long to re-check in full report OSB_ST_UNVERIFIED."""
    alpha = np.array([0.2, 0.3, 0.4, 0.9, 1.0])
    ch = engine.chan(ob.DISC_FD, 32)
    r = ch.solve(alpha, 1.0e4)
    c = ob.chan(ob.DISC_FD, 32)
    ref = np.array([c.solve_fd(a, 1.0e4)["omega"] for a in alpha])
    c.close()
    assert ref[0].real < -4.0 and ref[2].real < -4.0            # outside the scanned windows
    assert (r["status"] == 0).all()
    assert rel(r["omega"], ref).max() < 1e-8, rel(r["omega"], ref)
    monkeypatch.setenv("OSB_CHAN_NO_VERIFY", "1")               # what the scan alone would have returned
    blind = ch.solve(alpha, 1.0e4)
    monkeypatch.delenv("OSB_CHAN_NO_VERIFY")
    assert rel(blind["omega"][:3], ref[:3]).max() > 1e-3
    many = np.linspace(0.2, 0.45, 600)
    rm = ch.solve(many, 1.0e4)
    ch.free()
    assert set(np.unique(rm["status"])) <= {osb.ST_CONVERGED, osb.ST_UNVERIFIED}
    assert (rm["status"] == osb.ST_UNVERIFIED).sum() > 500      # flagged, not silently "converged"
    assert rm["status"][0] == osb.ST_CONVERGED and rm["status"][-1] == osb.ST_CONVERGED
