"""Host-side mirror of the reference's operator interface over libosb.so.

The reference (sscollis/os-stab) is compiled Fortran: its operators are
``SOLVE_BL`` (contebl.f:386), ``CONTE_IC`` (shoot.f:391), ``CONTE`` (shoot.f:26)
and the private ``CONTE`` copies of orrsom.f:165 / orrspace.f:209.  The product
boundary is the C ABI in ``include/osb.h``; the C++ twin drivers under
``host/`` and the Fortran ``bind(C)`` module in INTEGRATION.md sit directly on
it.  This module is a thin ctypes binding of the same ABI, used by ``tests/``
and ``bench.py`` (which need Python for pytest / torch.distributed plumbing).
It contains no numerics: every call goes to the CUDA library, and importing it
on a machine where ``libosb.so`` has not been built raises immediately.

Argument names follow the reference: ``nstep, testalpha, ymin(to), ymax(tf),
alpha, Re, c`` for the shooter; ``nbl, f2p, beta`` for the profile.  As in the
reference, ``alpha``/``Re`` passed to the shooter are in the solver's units
(after the driver's sqrt(2) rescale, contebl.f:181-182); :func:`contebl_units`
applies that rescale.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libosb.so"

FLOW_CONTEBL, FLOW_CONTE, FLOW_ORRSOM, FLOW_ORRSPACE = 0, 1, 2, 3
INT_RK4, INT_CK45 = 0, 1
ST_CONVERGED, ST_MAXCOUNT, ST_NONFINITE, ST_UNVERIFIED = 0, 1, 2, 3
E_ARG, E_CUDA, E_ALLOC, E_STATE = -1, -2, -3, -4

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class OsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libosb error {code}: {msg}")
        self.code = code


class ShootOpts(C.Structure):
    """``osb_shoot_opts`` (include/osb.h)."""

    _fields_ = [
        ("flow", C.c_int32),
        ("integrator", C.c_int32),
        ("nstep", C.c_int32),
        ("maxcount", C.c_int32),
        ("testalpha", C.c_double),
        ("ymin", C.c_double),
        ("ymax", C.c_double),
        ("errtol", C.c_double),
        ("eigtol", C.c_double),
    ]


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    p = Path(path) if path else Path(os.environ.get("OSB_LIB", LIB_PATH))  # OSB_LIB: experiment builds
    if not p.exists():
        raise ImportError(
            f"{p} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C os-stab_b200). There is no CPU fallback."
        )
    lib = C.CDLL(str(p))
    vp = C.c_void_p
    i32, i64, dbl = C.c_int32, C.c_int64, C.c_double
    sig = {
        "osb_version": (C.c_int, []),
        "osb_create": (C.c_int, [C.POINTER(vp), C.c_int]),
        "osb_create_multi": (C.c_int, [C.POINTER(vp), C.c_int, C.POINTER(C.c_int)]),
        "osb_destroy": (C.c_int, [vp]),
        "osb_device_count": (C.c_int, [vp]),
        "osb_last_error": (C.c_char_p, [vp]),
        "osb_stream": (vp, [vp]),
        "osb_launch_count": (i64, [vp]),
        "osb_profile_blasius": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, dbl, dbl, dbl, dbl, C.POINTER(vp)]),
        "osb_profile_blasius_batch": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, dbl, dbl, i64, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip]),
        "osb_profile_upload": (C.c_int, [vp, C.c_int, C.c_int, dbl, dbl, _dp, _dp, _dp, _dp, C.POINTER(vp)]),
        "osb_profile_shift": (C.c_int, [vp, vp, dbl, C.POINTER(vp)]),
        "osb_profile_channel": (C.c_int, [vp, C.POINTER(vp)]),
        "osb_profile_get": (C.c_int, [vp, vp, _dp, _dp, _dp, _dp, _dp, _ip, _dp]),
        "osb_profile_nbl": (C.c_int, [vp]),
        "osb_profile_free": (C.c_int, [vp, vp]),
        "osb_shoot_defaults": (C.c_int, [C.c_int, C.POINTER(ShootOpts)]),
        "osb_shoot_temporal": (C.c_int, [vp, C.POINTER(ShootOpts), vp, i64, _dp, _dp, _dp, _ip, _ip, _ip, _dp, _dp, i32]),
        "osb_shoot_temporal_dev": (C.c_int, [vp, C.POINTER(ShootOpts), vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, i32]),
        "osb_shoot_spatial_lines": (C.c_int, [vp, C.POINTER(ShootOpts), vp, i64, i32, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _ip, _dp, i32]),
        "osb_shoot_neutral": (C.c_int, [vp, C.POINTER(ShootOpts), vp, i64, _dp, _dp, _dp, dbl, dbl, i32, _dp, _dp, _ip, _ip]),
        "osb_shoot_grid_temporal": (C.c_int, [vp, C.POINTER(ShootOpts), vp, i32, _dp, i32, _dp, dbl, dbl, _dp, i32, _dp, _ip, _ip, _ip, _dp, _ip]),
        "osb_eigfun": (C.c_int, [vp, C.POINTER(ShootOpts), vp, i64, _dp, _dp, _dp, _dp, _dp, _dp, _ip]),
        "osb_measure_fp64_peak": (C.c_int, [vp, dbl, _dp, _dp]),
        "osb_measure_dmma_peak": (C.c_int, [vp, dbl, _dp]),
        "osb_chan_create": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp)]),
        "osb_chan_create_bl": (C.c_int, [vp, C.c_int, C.c_int, dbl, _dp, C.c_int, C.POINTER(vp)]),
        "osb_chan_bl_points": (C.c_int, [C.c_int, C.c_int, dbl, _dp]),
        "osb_chan_free": (C.c_int, [vp, vp]),
        "osb_chan_tables": (C.c_int, [vp, vp, _dp, _dp, _dp, _dp, _dp]),
        "osb_chan_eig": (C.c_int, [vp, vp, dbl, i64, _dp, _dp, _dp, _dp, _dp, C.c_int, _ip, _ip]),
        "osb_chan_spectrum": (C.c_int, [vp, vp, dbl, i64, _dp, _dp, _ip]),
        "osb_chan_neutral": (C.c_int, [vp, vp, i64, _dp, _dp, dbl, dbl, i32, _dp, _dp, _ip, _ip, _dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


EXPORTED_SYMBOLS = (
    "osb_version osb_create osb_create_multi osb_destroy osb_device_count osb_last_error osb_stream "
    "osb_launch_count osb_profile_blasius osb_profile_blasius_batch osb_profile_upload "
    "osb_profile_shift osb_profile_channel osb_profile_get osb_profile_nbl osb_profile_free osb_shoot_defaults "
    "osb_shoot_temporal osb_shoot_temporal_dev osb_shoot_spatial_lines osb_shoot_neutral osb_shoot_grid_temporal osb_eigfun osb_measure_fp64_peak osb_measure_dmma_peak "
    "osb_chan_create osb_chan_create_bl osb_chan_bl_points osb_chan_free osb_chan_tables osb_chan_eig osb_chan_spectrum osb_chan_neutral"
).split()

DISC_FD, DISC_CHEB, DISC_CHEB_NUM = 0, 1, 2
ADJ_PENCIL, ADJ_INVBA = 0, 1


def contebl_units(alpha, Re):
    """The driver-side rescale of contebl.f:181-182 / orrsom.f:129-130."""
    s2 = math.sqrt(2.0)
    return np.asarray(alpha) * s2, np.asarray(Re) * s2


def shoot_defaults(flow: int, lib: C.CDLL | None = None) -> ShootOpts:
    lib = lib or load_library()
    o = ShootOpts()
    rc = lib.osb_shoot_defaults(flow, C.byref(o))
    if rc:
        raise OsbError(rc, "osb_shoot_defaults")
    return o


def _f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} doubles, got {a.size}")
    return a


def _cplx_pairs(a, n):
    """complex or (n,2) float array -> contiguous float64 view of 2n doubles."""
    a = np.asarray(a)
    if np.iscomplexobj(a) or a.size == n:  # complex, or real values (imaginary part 0)
        a = np.ascontiguousarray(a, dtype=np.complex128).view(np.float64)
    return _f64(a.reshape(-1), 2 * n)


class Profile:
    """Device-resident mean profile (the reference's COMMON tables U, Uspl, d2U, d2Uspl)."""

    def __init__(self, engine, handle):
        self.engine = engine
        self.handle = handle

    @property
    def nbl(self):
        return self.engine.lib.osb_profile_nbl(self.handle)

    def tables(self):
        n = self.nbl + 1
        out = {k: np.zeros(n) for k in ("ydat", "U", "Uspl", "d2U", "d2Uspl")}
        npass = C.c_int32()
        f2p = C.c_double()
        self.engine._check(
            self.engine.lib.osb_profile_get(
                self.engine.ctx, self.handle, *[out[k].ctypes.data_as(_dp) for k in ("ydat", "U", "Uspl", "d2U", "d2Uspl")],
                C.byref(npass), C.byref(f2p)))
        out["npass"] = npass.value
        out["f2p"] = f2p.value
        return out

    def shifted(self, uc) -> "Profile":
        """SHIFT_BL (conteucbl.f:924): the profile in the frame moving with u_c."""
        h = C.c_void_p()
        self.engine._check(self.engine.lib.osb_profile_shift(self.engine.ctx, self.handle, float(uc), C.byref(h)))
        return Profile(self.engine, h)

    def free(self):
        if self.handle:
            self.engine.lib.osb_profile_free(self.engine.ctx, self.handle)
            self.handle = None


class Engine:
    """One ``osb_ctx``: a CUDA device (or several) plus its stream(s)."""

    def __init__(self, device: int | None = 0, devices=None, lib: C.CDLL | None = None):
        self.lib = lib or load_library()
        self.ctx = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.osb_create_multi(C.byref(self.ctx), len(devices), arr)
        else:
            rc = self.lib.osb_create(C.byref(self.ctx), int(device))
        if rc:
            raise OsbError(rc, "osb_create failed (no usable CUDA device? there is no CPU fallback)")

    def _check(self, rc):
        if rc:
            raise OsbError(rc, self.lib.osb_last_error(self.ctx).decode())

    def close(self):
        if self.ctx:
            self.lib.osb_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def launches(self) -> int:
        return int(self.lib.osb_launch_count(self.ctx))

    @property
    def stream(self) -> int:
        return int(self.lib.osb_stream(self.ctx) or 0)

    # ---- SOLVE_BL -----------------------------------------------------
    def solve_bl(self, variant=FLOW_CONTEBL, integrator=INT_CK45, nbl=1000, ymin=0.0, ymax=17.0, f2p=0.5, beta=0.0) -> Profile:
        h = C.c_void_p()
        self._check(self.lib.osb_profile_blasius(self.ctx, variant, integrator, nbl, ymin, ymax, f2p, beta, C.byref(h)))
        return Profile(self, h)

    def solve_bl_batch(self, f2p, beta, variant=FLOW_CONTEBL, integrator=INT_CK45, nbl=1000, ymin=0.0, ymax=17.0):
        f2p = _f64(f2p)
        beta = _f64(beta, f2p.size)
        n = f2p.size
        out = {k: np.zeros((n, nbl + 1)) for k in ("U", "Uspl", "d2U", "d2Uspl")}
        f2o = np.zeros(n)
        npass = np.zeros(n, dtype=np.int32)
        self._check(self.lib.osb_profile_blasius_batch(
            self.ctx, variant, integrator, nbl, ymin, ymax, n, f2p.ctypes.data_as(_dp), beta.ctypes.data_as(_dp),
            *[out[k].ctypes.data_as(_dp) for k in ("U", "Uspl", "d2U", "d2Uspl")], f2o.ctypes.data_as(_dp), npass.ctypes.data_as(_ip)))
        out["f2p"] = f2o
        out["npass"] = npass
        return out

    def upload_profile(self, variant, nbl, ymin, ymax, U, Uspl, d2U=None, d2Uspl=None) -> Profile:
        h = C.c_void_p()
        U = _f64(U, nbl + 1)
        Uspl = _f64(Uspl, nbl + 1)
        d2U = _f64(d2U, nbl + 1) if d2U is not None else None
        d2Uspl = _f64(d2Uspl, nbl + 1) if d2Uspl is not None else None
        self._check(self.lib.osb_profile_upload(
            self.ctx, variant, nbl, ymin, ymax, U.ctypes.data_as(_dp), Uspl.ctypes.data_as(_dp),
            d2U.ctypes.data_as(_dp) if d2U is not None else None,
            d2Uspl.ctypes.data_as(_dp) if d2Uspl is not None else None, C.byref(h)))
        return Profile(self, h)

    def channel_profile(self) -> Profile:
        h = C.c_void_p()
        self._check(self.lib.osb_profile_channel(self.ctx, C.byref(h)))
        return Profile(self, h)

    # ---- CONTE / CONTE_IC ---------------------------------------------
    def shoot_temporal(self, opts: ShootOpts, prof: Profile, alpha, Re, c, want_resid=False, hist_cap=0):
        """Batched CONTE_IC/CONTE.  alpha, c: complex arrays [npts]; Re: [npts].
        Returns dict(c, passes, qend, status[, resid, history])."""
        Re = _f64(Re)
        n = Re.size
        al = _cplx_pairs(alpha, n)
        cc = _cplx_pairs(c, n).copy()
        passes = np.zeros(n, dtype=np.int32)
        qend = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        resid = np.zeros(2 * n) if want_resid else None
        hist = np.zeros(n * hist_cap * 4) if hist_cap else None
        self._check(self.lib.osb_shoot_temporal(
            self.ctx, C.byref(opts), prof.handle, n, al.ctypes.data_as(_dp), Re.ctypes.data_as(_dp),
            cc.ctypes.data_as(_dp), passes.ctypes.data_as(_ip), qend.ctypes.data_as(_ip), status.ctypes.data_as(_ip),
            resid.ctypes.data_as(_dp) if want_resid else None, hist.ctypes.data_as(_dp) if hist_cap else None, hist_cap))
        out = dict(c=cc.view(np.complex128), passes=passes, qend=qend, status=status)
        if want_resid:
            out["resid"] = resid.reshape(n, 2)
        if hist_cap:
            out["history"] = hist.reshape(n, hist_cap, 4)
        return out

    def shoot_temporal_dev(self, opts: ShootOpts, prof: Profile, npts, d_alpha, d_Re, d_c, d_passes, d_qend, d_status,
                           d_resid=0, d_history=0, hist_cap=0):
        """Device-pointer form (ints are raw device addresses, e.g. torch ``data_ptr()``).
        Asynchronous on ``self.stream``."""
        self._check(self.lib.osb_shoot_temporal_dev(
            self.ctx, C.byref(opts), prof.handle, int(npts), d_alpha, d_Re, d_c, d_passes, d_qend, d_status,
            d_resid or None, d_history or None, hist_cap))

    def shoot_spatial_lines(self, opts: ShootOpts, prof: Profile, Re, omega0, domega, alpha0, niter, hist_cap=0):
        Re = _f64(Re)
        n = Re.size
        om = _cplx_pairs(omega0, n)
        dom = _f64(domega, n)
        a0 = _cplx_pairs(alpha0, n)
        out = np.zeros(n * niter * 2)
        passes = np.zeros(n * niter, dtype=np.int32)
        qend = np.zeros(n * niter, dtype=np.int32)
        status = np.zeros(n * niter, dtype=np.int32)
        hist = np.zeros(n * niter * hist_cap * 4) if hist_cap else None
        self._check(self.lib.osb_shoot_spatial_lines(
            self.ctx, C.byref(opts), prof.handle, n, niter, Re.ctypes.data_as(_dp), om.ctypes.data_as(_dp),
            dom.ctypes.data_as(_dp), a0.ctypes.data_as(_dp), out.ctypes.data_as(_dp), passes.ctypes.data_as(_ip),
            qend.ctypes.data_as(_ip), status.ctypes.data_as(_ip), hist.ctypes.data_as(_dp) if hist_cap else None, hist_cap))
        r = dict(alpha=out.view(np.complex128).reshape(n, niter), passes=passes.reshape(n, niter),
                 qend=qend.reshape(n, niter), status=status.reshape(n, niter))
        if hist_cap:
            r["history"] = hist.reshape(n, niter, hist_cap, 4)
        return r

    def shoot_neutral(self, opts: ShootOpts, prof: Profile, Re, alpha0, c0, dalpha_rel=1.0e-3, tol=1.0e-10, maxit=20):
        """Neutral points Im c(alpha, Re) = 0 by a batched secant iteration on alpha.
        Returns dict(alpha, c, steps, status)."""
        Re = _f64(Re)
        n = Re.size
        a0 = _f64(alpha0, n)
        cc = _cplx_pairs(c0, n)
        alpha = np.zeros(n)
        cout = np.zeros(2 * n)
        steps = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        self._check(self.lib.osb_shoot_neutral(
            self.ctx, C.byref(opts), prof.handle, n, Re.ctypes.data_as(_dp), a0.ctypes.data_as(_dp), cc.ctypes.data_as(_dp),
            float(dalpha_rel), float(tol), int(maxit), alpha.ctypes.data_as(_dp), cout.ctypes.data_as(_dp),
            steps.ctypes.data_as(_ip), status.ctypes.data_as(_ip)))
        return dict(alpha=alpha, c=cout.view(np.complex128), steps=steps, status=status)

    def shoot_grid_temporal(self, opts: ShootOpts, prof: Profile, alpha, Re, anchor_alpha, anchor_Re, anchor_c, coarse=0,
                            want_seed=False):
        """Self-seeding sweep over the tensor grid alpha[na] x Re[nre] from one anchor guess (include/osb.h).
        Returns dict(c[nre, na], passes, qend, status, info[, seed])."""
        alpha = _f64(alpha)
        Re = _f64(Re)
        na, nre = alpha.size, Re.size
        n = na * nre
        ac = _cplx_pairs(np.atleast_1d(anchor_c), 1)
        c = np.zeros(2 * n)
        passes = np.zeros(n, dtype=np.int32)
        qend = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        seed = np.zeros(2 * n) if want_seed else None
        info = np.zeros(4, dtype=np.int32)
        self._check(self.lib.osb_shoot_grid_temporal(
            self.ctx, C.byref(opts), prof.handle, na, alpha.ctypes.data_as(_dp), nre, Re.ctypes.data_as(_dp),
            float(anchor_alpha), float(anchor_Re), ac.ctypes.data_as(_dp), int(coarse), c.ctypes.data_as(_dp),
            passes.ctypes.data_as(_ip), qend.ctypes.data_as(_ip), status.ctypes.data_as(_ip),
            seed.ctypes.data_as(_dp) if want_seed else None, info.ctypes.data_as(_ip)))
        out = dict(c=c.view(np.complex128).reshape(nre, na), passes=passes.reshape(nre, na), qend=qend.reshape(nre, na),
                   status=status.reshape(nre, na), info=dict(zip(("coarse_solves", "rounds", "bisections", "given_up"), info.tolist())))
        if want_seed:
            out["seed"] = seed.view(np.complex128).reshape(nre, na)
        return out

    def eigfun(self, opts: ShootOpts, prof: Profile, Re, eig, alpha=None, omega=None):
        Re = _f64(Re)
        n = Re.size
        ev = _cplx_pairs(eig, n)
        al = _cplx_pairs(alpha, n) if alpha is not None else None
        om = _cplx_pairs(omega, n) if omega is not None else None
        y = np.zeros(n * (opts.nstep + 1) * 8)
        vmax = np.zeros(2 * n)
        qend = np.zeros(n, dtype=np.int32)
        self._check(self.lib.osb_eigfun(
            self.ctx, C.byref(opts), prof.handle, n, al.ctypes.data_as(_dp) if al is not None else None,
            Re.ctypes.data_as(_dp), ev.ctypes.data_as(_dp), om.ctypes.data_as(_dp) if om is not None else None,
            y.ctypes.data_as(_dp), vmax.ctypes.data_as(_dp), qend.ctypes.data_as(_ip)))
        return dict(y=y.view(np.complex128).reshape(n, opts.nstep + 1, 4), vmax=vmax.view(np.complex128), qend=qend)

    # ---- MAKE_MATRIX + SOLVE_ORR_SOM (channel matrix path) ---------------
    def chan(self, disc, N) -> "Chan":
        h = C.c_void_p()
        self._check(self.lib.osb_chan_create(self.ctx, disc, N, C.byref(h)))
        return Chan(self, h, N)

    def chan_bl(self, disc, N, lmap, prof: Profile = None, U=None, merge=None) -> "Chan":
        """Mapped boundary-layer operator (orrcolbl.f / orrfdbl.f) on y = lmap (1+eta)/(1-eta).
        The mean flow is either a Blasius/Falkner-Skan `prof` of this engine (evaluated at the grid points
        with the cubic spline of its tables, free stream beyond its ymax) or explicit values U[N+1]
        with `merge` = the last index of the free stream (U[0..merge] = 1)."""
        y = np.zeros(N + 1)
        if self.lib.osb_chan_bl_points(disc, N, float(lmap), y.ctypes.data_as(_dp)):
            raise ValueError("bad (disc, N, lmap)")
        if U is None:
            if prof is None:
                raise ValueError("chan_bl needs a profile or U values")
            t = prof.tables()
            yd, Ut, Us = t["ydat"], t["U"], t["Uspl"]
            h = yd[1] - yd[0]
            inside = np.isfinite(y) & (y <= yd[-1])
            merge = int(np.nonzero(~inside)[0].max()) if (~inside).any() else 0
            yy = np.where(inside, y, yd[-1])
            i = np.clip(((yy - yd[0]) / h).astype(np.int64), 0, yd.size - 2)
            A = (yd[i + 1] - yy) / h
            B = (yy - yd[i]) / h
            U = A * Ut[i] + B * Ut[i + 1] + ((A ** 3 - A) * Us[i] + (B ** 3 - B) * Us[i + 1]) * h * h / 6.0
            U = np.where(inside, U, 1.0)
        U = _f64(U, N + 1)
        if merge is None:
            raise ValueError("merge is required with explicit U")
        hnd = C.c_void_p()
        self._check(self.lib.osb_chan_create_bl(self.ctx, disc, N, float(lmap), U.ctypes.data_as(_dp), int(merge), C.byref(hnd)))
        ch = Chan(self, hnd, N)
        ch.y, ch.U, ch.merge = y, U, int(merge)
        return ch

    def measure_fp64_peak(self, ms_target=200.0):
        tf = C.c_double()
        mhz = C.c_double()
        self._check(self.lib.osb_measure_fp64_peak(self.ctx, ms_target, C.byref(tf), C.byref(mhz)))
        return tf.value, mhz.value

    def measure_dmma_peak(self, ms_target=200.0):
        tf = C.c_double()
        self._check(self.lib.osb_measure_dmma_peak(self.ctx, ms_target, C.byref(tf)))
        return tf.value


class Chan:
    """Grid + derivative tables of one (discretisation, N): the reference's COMMON
    /data/ /deriv/ after MAKE_CHANNEL_METRICS, MAKE_DERIVATIVES, INIT_CHANNEL_PROFILE."""

    def __init__(self, engine, handle, N):
        self.engine, self.handle, self.N = engine, handle, N

    def tables(self):
        n1 = self.N + 1
        out = dict(eta=np.zeros(n1), u=np.zeros(n1), d2u=np.zeros(n1), D2=np.zeros((n1, n1)), D4=np.zeros((n1, n1)))
        self.engine._check(self.engine.lib.osb_chan_tables(
            self.engine.ctx, self.handle, *[out[k].ctypes.data_as(_dp) for k in ("eta", "u", "d2u", "D2", "D4")]))
        return out

    def solve(self, alpha, Re, sigma0=None, want_evec=False, want_adj=None):
        """SOLVE_ORR_SOM for a batch of alpha (orrfdchan.f:116-121 hoisted into one call).
        want_adj = ADJ_PENCIL | ADJ_INVBA also returns the left eigenvector in that form (include/osb.h)."""
        alpha = np.atleast_1d(np.asarray(alpha, dtype=np.complex128))
        n = alpha.size
        al = _cplx_pairs(alpha, n)
        sg = _cplx_pairs(np.atleast_1d(sigma0), n) if sigma0 is not None else None
        om = np.zeros(2 * n)
        ev = np.zeros(2 * n * (self.N + 1)) if want_evec else None
        av = np.zeros(2 * n * (self.N + 1)) if want_adj is not None else None
        iters = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        self.engine._check(self.engine.lib.osb_chan_eig(
            self.engine.ctx, self.handle, float(Re), n, al.ctypes.data_as(_dp),
            sg.ctypes.data_as(_dp) if sg is not None else None, om.ctypes.data_as(_dp),
            ev.ctypes.data_as(_dp) if want_evec else None, av.ctypes.data_as(_dp) if av is not None else None,
            int(want_adj or 0), iters.ctypes.data_as(_ip), status.ctypes.data_as(_ip)))
        r = dict(omega=om.view(np.complex128), iters=iters, status=status)
        if av is not None:
            r["avec"] = av.view(np.complex128).reshape(n, self.N + 1)
        if want_evec:
            r["evec"] = ev.view(np.complex128).reshape(n, self.N + 1)
        return r

    def spectrum(self, alpha, Re):
        """All N-1 eigenvalues of the interior pencil at each alpha, ascending in Im(omega)
        (what SOLVE_ORR_SOM sorts before it keeps one mode).  Returns dict(spectrum[npts, N-1], status)."""
        alpha = np.atleast_1d(np.asarray(alpha, dtype=np.complex128))
        n = alpha.size
        al = _cplx_pairs(alpha, n)
        out = np.zeros(2 * n * (self.N - 1))
        status = np.zeros(n, dtype=np.int32)
        self.engine._check(self.engine.lib.osb_chan_spectrum(
            self.engine.ctx, self.handle, float(Re), n, al.ctypes.data_as(_dp), out.ctypes.data_as(_dp),
            status.ctypes.data_as(_ip)))
        return dict(spectrum=out.view(np.complex128).reshape(n, self.N - 1), status=status)

    def neutral(self, Re, alpha0, sfac=0.01, tol=1.0e-8, maxit=200, want_trace=False):
        """The Newton iteration on alpha of orrncbl.f:109-116, one search per Re."""
        Re = _f64(np.atleast_1d(Re))
        n = Re.size
        a0 = _f64(np.broadcast_to(np.asarray(alpha0, dtype=np.float64), (n,)).copy(), n)
        aout = np.zeros(n)
        om = np.zeros(2 * n)
        steps = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        tr = np.zeros(n * maxit * 7) if want_trace else None
        self.engine._check(self.engine.lib.osb_chan_neutral(
            self.engine.ctx, self.handle, n, Re.ctypes.data_as(_dp), a0.ctypes.data_as(_dp), sfac, tol, maxit,
            aout.ctypes.data_as(_dp), om.ctypes.data_as(_dp), steps.ctypes.data_as(_ip), status.ctypes.data_as(_ip),
            tr.ctypes.data_as(_dp) if want_trace else None))
        r = dict(alpha=aout, omega=om.view(np.complex128), steps=steps, status=status)
        if want_trace:
            r["trace"] = tr.reshape(n, maxit, 7)
        return r

    def free(self):
        if self.handle:
            self.engine.lib.osb_chan_free(self.engine.ctx, self.handle)
            self.handle = None


# ---- sharding of a sweep over ranks: host logic, no numerics ----
def shard_bounds(npts: int, world: int, rank: int, tile: int = 4096):
    """Cyclic-by-tile partition of the flattened point index: tile t goes to rank
    t % world.  Returns the index array owned by ``rank`` (sorted)."""
    ntile = (npts + tile - 1) // tile
    mine = np.arange(rank, ntile, world, dtype=np.int64)
    idx = (mine[:, None] * tile + np.arange(tile, dtype=np.int64)[None, :]).reshape(-1)
    return idx[idx < npts]
