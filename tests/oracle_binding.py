"""ctypes binding of oracle/_build/libos_oracle.so -- TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product package."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "_build" / "libos_oracle.so"
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

FLOW_CONTEBL, FLOW_CONTE, FLOW_ORRSOM, FLOW_ORRSPACE = 0, 1, 2, 3
INT_RK4, INT_CK45 = 0, 1


PMATH_LIB = ROOT / "tests" / "_build" / "libpmath_host.so"


def build():
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle")], check=True)


def build_pmath():
    """Host build (no FMA contraction) of the product's portable elementary functions, for osr_set_math."""
    src = ROOT / "tests" / "pmath_host.c"
    hdr = ROOT / "os-stab_b200" / "csrc" / "osb_pmath.h"
    if (not PMATH_LIB.exists() or PMATH_LIB.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime)):
        PMATH_LIB.parent.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                        "-o", str(PMATH_LIB), str(src), "-lm"], check=True)
    return C.CDLL(str(PMATH_LIB))


class _OsrMath(C.Structure):
    _fields_ = [("exp", C.c_void_p), ("sincos", C.c_void_p), ("hypot", C.c_void_p), ("acos", C.c_void_p)]


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        i, d, i64 = C.c_int, C.c_double, C.c_int64
        lib.osr_solve_bl.restype = i
        lib.osr_solve_bl.argtypes = [i, i, i, d, d, d, d, _dp, _dp, _dp, _dp, _dp, _dp]
        lib.osr_shoot_temporal.restype = i
        lib.osr_shoot_temporal.argtypes = [i, i, i, d, d, d, i, _dp, d, d, d, d, d, _dp, _ip, _dp, i, _dp, _dp]
        lib.osr_shoot_temporal_batch.restype = i
        lib.osr_shoot_temporal_batch.argtypes = [i, i, i, d, d, d, i, _dp, i64, _dp, _dp, _dp, _ip, _ip, _ip, _dp]
        lib.osr_shoot_spatial_line.restype = i
        lib.osr_shoot_spatial_line.argtypes = [i, i, d, d, d, i, _dp, d, d, d, d, d, d, i, _dp, _ip, _ip, _ip, _dp]
        lib.osr_shoot_spatial_point.restype = i
        lib.osr_shoot_spatial_point.argtypes = [i, i, d, d, d, i, _dp, d, d, d, d, d, _dp, _ip, _dp, _dp]

    def set_math(self, kind=None):
        """Which elementary functions the shooting restatement uses: None = libm's complex cabs/csqrt/cexp
        and acos, as the reference's build; "libm" = glibc's complex formulas restated over libm's real
        exp/sincos/hypot/acos (must equal None bit for bit); "portable" = the same formulas over the
        IEEE-only functions of osb_pmath.h that the CUDA strict build uses."""
        if kind is None:
            self.lib.osr_set_math(None)
        elif kind == "libm":
            self.lib.osr_set_math_libm()
        elif kind == "portable":
            pm = build_pmath()
            self._pm = pm  # keep the library alive
            m = _OsrMath(*[C.cast(getattr(pm, n), C.c_void_p).value for n in ("pmh_exp", "pmh_sincos", "pmh_hypot", "pmh_acos")])
            self.lib.osr_set_math.argtypes = [C.POINTER(_OsrMath)]
            self.lib.osr_set_math(C.byref(m))
        else:
            raise ValueError(kind)

    def solve_bl(self, variant=FLOW_CONTEBL, integrator=INT_CK45, nbl=1000, ymin=0.0, ymax=17.0, f2p=0.5, beta=0.0):
        """Returns (prof[5, nbl+2] = ydat,U,Uspl,d2U,d2Uspl zero padded, npass, f2p_last)."""
        prof = np.zeros((5, nbl + 2))
        f2 = C.c_double()
        n = self.lib.osr_solve_bl(variant, integrator, nbl, ymin, ymax, f2p, beta,
                                  *[prof[k].ctypes.data_as(_dp) for k in range(5)], C.byref(f2))
        return prof, n, f2.value

    def shoot_temporal(self, flow, integrator, nstep, testalpha, ymin, ymax, prof, alpha, Re, c,
                       hist_cap=50, eigfun=False):
        out = np.zeros(6)
        iout = np.zeros(3, dtype=np.int32)
        hist = np.zeros(4 * hist_cap)
        nbl = prof.shape[1] - 2 if prof is not None else 0
        y = np.zeros((nstep + 1) * 8) if eigfun else None
        vmax = np.zeros(2) if eigfun else None
        alpha = complex(alpha)
        c = complex(c)
        self.lib.osr_shoot_temporal(
            flow, integrator, nstep, testalpha, ymin, ymax, nbl,
            prof.ctypes.data_as(_dp) if prof is not None else None, alpha.real, alpha.imag, Re, c.real, c.imag,
            out.ctypes.data_as(_dp), iout.ctypes.data_as(_ip), hist.ctypes.data_as(_dp), hist_cap,
            y.ctypes.data_as(_dp) if eigfun else None, vmax.ctypes.data_as(_dp) if eigfun else None)
        r = dict(c=complex(out[0], out[1]), next=complex(out[2], out[3]), abserr=out[4], absdc=out[5],
                 passes=int(iout[0]), qend=int(iout[1]), status=int(iout[2]),
                 history=hist.reshape(hist_cap, 4)[: int(iout[0])])
        if eigfun:
            r["y"] = y.view(np.complex128).reshape(nstep + 1, 4)
            r["vmax"] = complex(vmax[0], vmax[1])
        return r

    def shoot_temporal_batch(self, flow, integrator, nstep, testalpha, ymin, ymax, prof, alpha, Re, c):
        Re = np.ascontiguousarray(Re, dtype=np.float64)
        n = Re.size
        al = np.ascontiguousarray(alpha, dtype=np.complex128).view(np.float64)
        cc = np.ascontiguousarray(c, dtype=np.complex128).copy().view(np.float64)
        passes = np.zeros(n, dtype=np.int32)
        qend = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        nbl = prof.shape[1] - 2 if prof is not None else 0
        absdc = np.zeros(n)
        self.lib.osr_shoot_temporal_batch(
            flow, integrator, nstep, testalpha, ymin, ymax, nbl,
            prof.ctypes.data_as(_dp) if prof is not None else None, n, al.ctypes.data_as(_dp),
            Re.ctypes.data_as(_dp), cc.ctypes.data_as(_dp), passes.ctypes.data_as(_ip), qend.ctypes.data_as(_ip),
            status.ctypes.data_as(_ip), absdc.ctypes.data_as(_dp))
        return dict(c=cc.view(np.complex128), passes=passes, qend=qend, status=status, absdc=absdc)

    def shoot_spatial_line(self, integrator, nstep, testalpha, ymin, ymax, prof, Re, omega, domega, alpha, niter):
        rows = np.zeros(4 * niter)
        ipass = np.zeros(niter, dtype=np.int32)
        iq = np.zeros(niter, dtype=np.int32)
        ist = np.zeros(niter, dtype=np.int32)
        absdc = np.zeros(niter)
        omega = complex(omega)
        alpha = complex(alpha)
        self.lib.osr_shoot_spatial_line(
            integrator, nstep, testalpha, ymin, ymax, prof.shape[1] - 2, prof.ctypes.data_as(_dp), Re,
            omega.real, omega.imag, domega, alpha.real, alpha.imag, niter, rows.ctypes.data_as(_dp),
            ipass.ctypes.data_as(_ip), iq.ctypes.data_as(_ip), ist.ctypes.data_as(_ip), absdc.ctypes.data_as(_dp))
        rows = rows.reshape(niter, 4)
        return dict(omega=rows[:, 0] + 1j * rows[:, 1], alpha=rows[:, 2] + 1j * rows[:, 3],
                    passes=ipass, qend=iq, status=ist, absdc=absdc)


def _spatial_point(self, integrator, nstep, testalpha, ymin, ymax, prof, Re, omega, alpha):
    """orrspace with niter = 1 and eigfun = 1 (space.inp)."""
    out = np.zeros(2)
    iout = np.zeros(3, dtype=np.int32)
    y = np.zeros((nstep + 1) * 8)
    vmax = np.zeros(2)
    omega = complex(omega)
    alpha = complex(alpha)
    self.lib.osr_shoot_spatial_point(integrator, nstep, testalpha, ymin, ymax, prof.shape[1] - 2,
                                     prof.ctypes.data_as(_dp), Re, omega.real, omega.imag, alpha.real, alpha.imag,
                                     out.ctypes.data_as(_dp), iout.ctypes.data_as(_ip), y.ctypes.data_as(_dp),
                                     vmax.ctypes.data_as(_dp))
    return dict(alpha=complex(out[0], out[1]), passes=int(iout[0]), qend=int(iout[1]), status=int(iout[2]),
                y=y.view(np.complex128).reshape(nstep + 1, 4), vmax=complex(vmax[0], vmax[1]))


Oracle.shoot_spatial_point = _spatial_point


def _profile_shift(self, prof, uc):
    """SHIFT_BL of conteucbl.f:924-968 applied to the tables of solve_bl()."""
    self.lib.osr_profile_shift.argtypes = [C.c_int, C.c_double, _dp, _dp]
    out = np.zeros_like(prof)
    self.lib.osr_profile_shift(prof.shape[1] - 2, float(uc), np.ascontiguousarray(prof).ctypes.data_as(_dp), out.ctypes.data_as(_dp))
    return out


Oracle.profile_shift = _profile_shift


def load(rebuild=True):
    if rebuild or not LIB.exists():
        build()
    return Oracle(C.CDLL(str(LIB)))


# ---------------------------------------------------------------- matrix path
DISC_FD, DISC_CHEB, DISC_CHEB_NUM = 0, 1, 2


def _find_openblas():
    import glob
    import scipy

    base = Path(scipy.__file__).resolve().parent.parent / "scipy.libs"
    cands = sorted(glob.glob(str(base / "libscipy_openblas-*.so"))) or sorted(glob.glob(str(base / "libscipy_openblas*.so")))
    if not cands:
        raise RuntimeError("scipy's bundled OpenBLAS not found")
    return cands[0]


class ChanOracle:
    """orrfdchan / orrcolchan / orrncbl restatement (oracle/os_chan_oracle.c)."""

    def __init__(self, lib, disc, n, bl=None):
        self.lib = lib
        self.n = n
        i, d = C.c_int, C.c_double
        lib.osr_lapack_open.argtypes = [C.c_char_p]
        lib.osr_chan_create.restype = C.c_void_p
        lib.osr_chan_create.argtypes = [i, i]
        lib.osr_chan_destroy.argtypes = [C.c_void_p]
        lib.osr_chan_get.argtypes = [C.c_void_p, i, _dp]
        lib.osr_chan_solve_fdstyle.argtypes = [C.c_void_p, d, d, d, _dp, _dp, _dp, _dp]
        lib.osr_chan_solve_col.argtypes = [C.c_void_p, d, d, d, _dp, i, _dp, _dp, _dp]
        lib.osr_chan_neutral.restype = i
        lib.osr_chan_neutral.argtypes = [C.c_void_p, d, d, d, d, i, _dp]
        if lib.osr_lapack_open(_find_openblas().encode()) != 0:
            raise RuntimeError("cannot open LAPACK")
        if bl is None:
            self.h = lib.osr_chan_create(disc, n)
        else:  # mapped boundary layer: bl = (lmap, U[N+1], merge)
            lib.osr_chan_create_bl.restype = C.c_void_p
            lib.osr_chan_create_bl.argtypes = [i, i, d, _dp, i]
            U = np.ascontiguousarray(bl[1], dtype=np.float64)
            self.h = lib.osr_chan_create_bl(disc, n, float(bl[0]), U.ctypes.data_as(_dp), int(bl[2]))

    def close(self):
        if self.h:
            self.lib.osr_chan_destroy(self.h)
            self.h = None

    def table(self, which):
        n1 = self.n + 1
        out = np.zeros(n1 if which < 4 else n1 * n1)
        self.lib.osr_chan_get(self.h, which, out.ctypes.data_as(_dp))
        return out if which < 4 else out.reshape(n1, n1)

    def solve_fd(self, alpha, Re, want_spectrum=False, want_evec=False, want_adj=False):
        alpha = complex(alpha)
        om = np.zeros(2)
        spec = np.zeros(2 * (self.n - 1)) if want_spectrum else None
        ev = np.zeros(2 * (self.n + 1)) if want_evec else None
        av = np.zeros(2 * (self.n + 1)) if want_adj else None
        self.lib.osr_chan_solve_fdstyle(self.h, alpha.real, alpha.imag, Re, om.ctypes.data_as(_dp),
                                        spec.ctypes.data_as(_dp) if want_spectrum else None,
                                        ev.ctypes.data_as(_dp) if want_evec else None,
                                        av.ctypes.data_as(_dp) if want_adj else None)
        r = dict(omega=complex(om[0], om[1]))
        if want_adj:
            r["avec"] = av.view(np.complex128)
        if want_spectrum:
            r["spectrum"] = spec.view(np.complex128)
        if want_evec:
            r["evec"] = ev.view(np.complex128)
        return r

    def solve_col(self, alpha, Re, which=-1):
        alpha = complex(alpha)
        spec = np.zeros(2 * (self.n + 1))
        ev = np.zeros(2 * (self.n + 1)) if which >= 0 else None
        av = np.zeros(2 * (self.n + 1)) if which >= 0 else None
        dots = np.zeros(4)
        self.lib.osr_chan_solve_col(self.h, alpha.real, alpha.imag, Re, spec.ctypes.data_as(_dp), which,
                                    ev.ctypes.data_as(_dp) if which >= 0 else None,
                                    av.ctypes.data_as(_dp) if which >= 0 else None, dots.ctypes.data_as(_dp))
        r = dict(spectrum=spec.view(np.complex128))
        if which >= 0:
            r["evec"] = ev.view(np.complex128)
            r["avec"] = av.view(np.complex128)
            r["dots"] = dots.view(np.complex128)
        return r

    def arbiter(self, alpha, Re, sigma, iters=4):
        """Quad-precision eigenvalue of the same discrete pencil nearest `sigma` (os_chan_arbiter.c)."""
        self.lib.osr_chan_arbiter.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double,
                                              C.c_double, C.c_double, C.c_int, _dp]
        D2 = np.ascontiguousarray(self.table(4)); D4 = np.ascontiguousarray(self.table(5))
        u = self.table(1); d2u = self.table(3)
        out = np.zeros(4)
        alpha = complex(alpha); sigma = complex(sigma)
        self.lib.osr_chan_arbiter(self.n, D2.ctypes.data_as(_dp), D4.ctypes.data_as(_dp), u.ctypes.data_as(_dp),
                                  d2u.ctypes.data_as(_dp), alpha.real, alpha.imag, Re, sigma.real, sigma.imag, iters,
                                  out.ctypes.data_as(_dp))
        return complex(out[0], out[1])

    def neutral(self, Re, sfac, alpha_r, tol=1.0e-8, maxit=200):
        tr = np.zeros(7 * maxit)
        it = self.lib.osr_chan_neutral(self.h, Re, sfac, alpha_r, tol, maxit, tr.ctypes.data_as(_dp))
        return tr.reshape(maxit, 7)[:it]


def chan(disc, n, rebuild=False, bl=None):
    if rebuild or not LIB.exists():
        build()
    return ChanOracle(C.CDLL(str(LIB)), disc, n, bl=bl)
