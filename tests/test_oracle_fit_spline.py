"""fitSpline (fit_splie.d:26-85) restated in the oracle, pinned by the reference's own golden vectors (fit_splie.d:94-143)
and, for the spline underneath (mir.interpolate.spline, un-vendored), cross-checked against scipy's not-a-knot cubic spline.
CPU only."""
import ctypes as C

import numpy as np
import pytest

from mir_optim_b200._abi import LeastSquaresResultD, LeastSquaresSettingsD

# fit_splie.d:98-121
X = np.array([-1.0, 2, 4, 5, 8, 10, 12, 15, 19, 22])
Y0 = np.array([17.0, 0, 16, 4, 10, 15, 19, 5, 18, 6])
PY = np.array([-0.68361541, 7.28568719, 10.490694, 0.36192032, 11.91572713, 16.44546433, 17.66699525, 4.52730869, 19.22825394, -2.3242592])
PT = X + 0.5
# fit_splie.d:128-139 (lambda = 1e-3)
Y1 = np.array([15.898984945597563, 0.44978154774119194, 15.579636654078188, 4.028312405287987, 9.945895290402778,
               15.07778815727665, 18.877926155854535, 5.348699237978274, 16.898507797404278, 22.024920998359942])


def should_approx(a, b, rel=2.0 ** -20, abs_=2.0 ** -20):
    """mir.test.shouldApprox defaults (maxRelDiff = maxAbsDiff = 0x1p-20)."""
    a = np.asarray(a); b = np.asarray(b)
    return np.all((np.abs(a - b) <= abs_) | (np.abs(a - b) <= rel * np.maximum(np.abs(a), np.abs(b))))


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def fit(lib, lam, pt=PT, py=PY, x=X, l=None, u=None):
    lib.oracle_fit_spline_d.restype = C.c_int
    lib.oracle_fit_spline_d.argtypes = [C.POINTER(LeastSquaresSettingsD), C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_size_t,
                                        C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double,
                                        C.POINTER(C.c_double), C.POINTER(LeastSquaresResultD)]
    s = LeastSquaresSettingsD(); lib.mir_least_squares_init_d(C.byref(s))
    n = len(x)
    l = np.full(n, -np.inf) if l is None else l; u = np.full(n, np.inf) if u is None else u
    v = np.empty(n); r = LeastSquaresResultD()
    pt = np.ascontiguousarray(pt, dtype=np.float64); py = np.ascontiguousarray(py, dtype=np.float64); x = np.ascontiguousarray(x, dtype=np.float64)
    rc = lib.oracle_fit_spline_d(C.byref(s), len(pt), dp(pt), dp(py), n, dp(x), dp(l), dp(u), lam, dp(v), C.byref(r))
    return rc, v, r


def spline_eval(lib, x, v, t):
    lib.oracle_spline_eval_d.restype = None
    lib.oracle_spline_eval_d.argtypes = [C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]
    t = np.ascontiguousarray(t, dtype=np.float64); ov = np.empty_like(t); od = np.empty_like(t)
    lib.oracle_spline_eval_d(len(x), dp(np.ascontiguousarray(x)), dp(np.ascontiguousarray(v)), len(t), dp(t), dp(ov), dp(od))
    return ov, od


@pytest.mark.parametrize("n", [2, 3, 4, 5, 10, 37])
def test_restated_spline_is_the_not_a_knot_cubic_spline(oracle_lib, n):
    from scipy.interpolate import CubicSpline
    rng = np.random.default_rng(n)
    x = np.cumsum(rng.uniform(0.3, 2.0, n)); v = rng.normal(0, 5, n)
    t = np.concatenate([x, rng.uniform(x[0] - 1.0, x[-1] + 1.0, 200)])        # knots, interior points, extrapolation on both sides
    ov, od = spline_eval(oracle_lib, x, v, t)
    if n == 2:
        ref_v = v[0] + (v[1] - v[0]) / (x[1] - x[0]) * (t - x[0]); ref_d = np.full_like(t, (v[1] - v[0]) / (x[1] - x[0]))
    else:
        cs = CubicSpline(x, v, bc_type="not-a-knot"); ref_v = cs(t); ref_d = cs(t, 1)
    scale = np.abs(v).max()
    assert np.max(np.abs(ov - ref_v)) < 1e-11 * scale * 10
    assert np.max(np.abs(od - ref_d)) < 1e-10 * scale * 10
    assert np.max(np.abs(ov[:n] - v)) < 1e-13 * scale              # interpolates


def test_golden_vector_lambda_zero(oracle_lib):
    """fit_splie.d:123-126: ten points at knot + 0.5 and lambda = 0 recover the spline through y."""
    rc, v, r = fit(oracle_lib, 0.0)
    assert rc == 0 and r.status >= 0
    assert should_approx(v, Y0), (v, Y0)
    ov, _ = spline_eval(oracle_lib, X, v, X)                      # result.spline(x[i]), as the reference asserts it
    assert should_approx(ov, Y0)


def test_golden_vector_lambda_1e3(oracle_lib):
    """fit_splie.d:128-143 ("this case sensetive for numeric noise")."""
    rc, v, r = fit(oracle_lib, 1e-3)
    assert rc == 0 and r.status >= 0
    assert should_approx(v, Y1), (v - Y1)


def test_too_few_points_is_the_reference_exception(oracle_lib):
    rc, _, _ = fit(oracle_lib, 0.0, pt=PT[:5], py=PY[:5])          # fit_splie.d:45-49
    assert rc == -1
    rc, v, r = fit(oracle_lib, 1e-3, pt=PT[:5], py=PY[:5])         # allowed with lambda > 0
    assert rc == 0 and np.all(np.isfinite(v))


def test_bounds_are_honoured(oracle_lib):
    l = np.full(10, -1.0); u = np.full(10, 17.5)
    rc, v, r = fit(oracle_lib, 1e-3, l=l, u=u)
    assert rc == 0 and np.all(v >= l) and np.all(v <= u) and r.status >= -1
    assert np.any(v == u) or np.any(v == l)                       # (the unconstrained fit reaches 22 and 0.45)
    # the reference starts from zeros (fit_splie.d:55-56): bounds that exclude 0 are badBounds, the values stay untouched
    rc, v, r = fit(oracle_lib, 1e-3, l=np.full(10, 1.0), u=u)
    assert rc == 0 and r.status == -32 and np.all(v == 0)
