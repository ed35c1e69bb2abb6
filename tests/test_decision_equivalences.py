"""CPU checks of two device-side reformulations that must not change a single decision (they are emulated here in
IEEE double / integer arithmetic, which is what the kernels execute):

* the orthonormality test of the analytic inner products (orrspace.f:357-358) decided in squared form outside a
  1e-12-wide band around the threshold (csrc/shoot_kernels.cu, need_norm<true>);
* the pivot search of the warp-per-instance LU by three integer reductions (csrc/chan_kernels.cu, w_lu) against the
  sequential IZAMAX rule (largest |re| + |im|, lowest row on ties)."""
import numpy as np


def _sqrt_form(lhs, a2, b2, thr, strict):
    rhs = thr * (np.sqrt(a2) * np.sqrt(b2))
    return lhs > rhs if strict else lhs >= rhs


def _quick_form(lhs, a2, b2, thr):
    """Returns +1 / -1 where the squared form decides, 0 where the kernel falls back to the square roots."""
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        l2 = lhs * lhs
        r2 = (thr * thr) * (a2 * b2)
        ok = (r2 > 1.0e-250) & (r2 < 1.0e250)
        yes = ok & (l2 > r2 * 1.000000000001)
        no = ok & ~yes & (l2 < r2 * 0.999999999999)
    return yes.astype(int) - no.astype(int)


def test_squared_orthonormality_test_never_contradicts_the_square_root_form():
    rng = np.random.default_rng(11)
    n = 2_000_000
    thr = np.where(rng.random(n) < 0.5, 0.81, 0.04) * (1.0 + 0.1 * (rng.random(n) - 0.5))
    a2 = 10.0 ** rng.uniform(-120.0, 120.0, n)
    b2 = 10.0 ** rng.uniform(-120.0, 120.0, n)
    rhs = thr * (np.sqrt(a2) * np.sqrt(b2))
    # left-hand sides from far below to far above the threshold, most of them within 1e-16 .. 1e-9 of it
    eps = np.where(rng.random(n) < 0.8, 10.0 ** rng.uniform(-16.5, -9.0, n), 10.0 ** rng.uniform(-9.0, 0.5, n))
    lhs = rhs * (1.0 + np.where(rng.random(n) < 0.5, eps, -np.minimum(eps, 0.999)))
    lhs[: n // 50] = rhs[: n // 50]  # exactly on the threshold
    q = _quick_form(lhs, a2, b2, thr)
    for strict in (0, 1):
        ref = _sqrt_form(lhs, a2, b2, thr, strict)
        assert not np.any((q == 1) & ~ref), "squared form said 'normalise' where the reference form does not"
        assert not np.any((q == -1) & ref), "squared form said 'no' where the reference form normalises"
    # the quick form must decide all but the immediate neighbourhood of the threshold ...
    far = np.abs(lhs / rhs - 1.0) > 2.0e-12
    assert np.all(q[far] != 0)
    # ... and must leave the band, and everything outside the normal range, to the square roots
    assert np.all(q[np.abs(lhs / rhs - 1.0) < 1.0e-13] == 0)
    big = np.array([1.0e200, 1.0e-200, np.inf, np.nan, 0.0])
    assert np.all(_quick_form(big, big, big, np.full(5, 0.81)) == 0)


def _izamax_rule(v, k):
    best, bi = -1.0, k
    for r in range(k, len(v)):
        a = abs(v[r].real) + abs(v[r].imag)
        if a > best:
            best, bi = a, r
    return bi


def _three_reductions(v, k):
    n = len(v)
    keys, rows = [], []
    for lane in range(32):
        best, bi = -1, k
        for r in (lane, lane + 32):
            if k <= r < n:
                a = np.float64(abs(v[r].real) + abs(v[r].imag)).view(np.int64) & 0x7FFFFFFFFFFFFFFF
                if a > best:
                    best, bi = int(a), r
        keys.append(best + 1)
        rows.append(bi)
    hi = [x >> 32 for x in keys]
    lo = [x & 0xFFFFFFFF for x in keys]
    mh = max(hi)
    ml = max(l if h == mh else 0 for h, l in zip(hi, lo))
    return min(r if (h == mh and l == ml) else 0xFFFFFFFF for h, l, r in zip(hi, lo, rows))


def test_pivot_search_by_integer_reductions_follows_izamax():
    rng = np.random.default_rng(5)
    for trial in range(400):
        n = int(rng.integers(2, 65))
        v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        if trial % 4 == 1:   # ties: repeated magnitudes, also across the two rows of a lane
            v[rng.integers(0, n, size=n // 2)] = v[rng.integers(0, n)]
        if trial % 4 == 2:   # tiny and huge entries, exact zeros
            v *= 10.0 ** rng.uniform(-300, 300, n)
            v[rng.integers(0, n, size=max(1, n // 3))] = 0.0
        if trial % 4 == 3:   # an all-zero column below the diagonal
            v[:] = 0.0
        for k in sorted(set(int(x) for x in rng.integers(0, n, size=6)) | {0, n - 1}):
            assert _three_reductions(v, k) == _izamax_rule(v, k), (trial, n, k)
