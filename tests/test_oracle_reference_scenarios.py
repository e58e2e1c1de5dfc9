"""Pins the CPU oracle (oracle/lm_oracle.cpp) against every assertion of the reference's own
unit tests: least_squares.d:217-434 (T1-T6) and boxcqp.d:381-402 (BQ1).  CPU only.

The reference draws its noise from mir.random `Random(12345)` (least_squares.d:352, 381), which
is not available here; the noisy scenarios (T4, T5) use numpy noise of the same sigma -- their
assertions are loose (0.05 / bounds only), exactly as in the reference.
"""
import ctypes as C

import numpy as np
import pytest

from mir_optim_b200.api import LeastSquaresStatus, LeastSquaresException

INF = np.inf


def nrm2(v):
    return float(np.sqrt(np.sum(np.asarray(v) ** 2)))


def test_t1_linear_with_jacobian(oracle):
    """least_squares.d:217-245"""
    s = oracle.settings()
    x = np.array([100.0, 100.0]); l = np.full(2, -INF); u = np.full(2, INF)

    def f(x, y): y[0] = x[0]; y[1] = 2 - x[1]
    def g(x, J): J[0, 0] = 1; J[0, 1] = 0; J[1, 0] = 0; J[1, 1] = -1
    r = oracle.optimize(s, 2, x, l, u, f, g)
    assert nrm2(x - [0, 2]) < 1e-8
    assert r.status == LeastSquaresStatus.fConverged


def rosen_f(x, y): y[0] = 10 * (x[1] - x[0] ** 2); y[1] = 1 - x[0]
def rosen_g(x, J): J[0, 0] = -20 * x[0]; J[0, 1] = 10; J[1, 0] = -1; J[1, 1] = 0


def test_t2_rosenbrock_fd_with_thread_manager(oracle):
    """least_squares.d:247-273 (taskPool thread manager -> serial task loop here)"""
    s = oracle.settings()
    x = np.array([-1.2, 1.0]); l = np.full(2, -INF); u = np.full(2, INF)
    calls = []

    def tm(count, task):
        calls.append(count)
        for i in range(count):
            task(1, 0, i)
    oracle.optimize(s, 2, x, l, u, rosen_f, None, tm)
    assert nrm2(x - [1, 1]) < 1e-6
    assert calls and all(c == 2 for c in calls)


def test_t2_rosenbrock_fd_default_tm(oracle):
    s = oracle.settings()
    x = np.array([-1.2, 1.0]); l = np.full(2, -INF); u = np.full(2, INF)
    oracle.optimize(s, 2, x, l, u, rosen_f)
    assert nrm2(x - [1, 1]) < 1e-6


def test_t3_rosenbrock_analytic_and_box(oracle):
    """least_squares.d:275-331"""
    s = oracle.settings()
    x = np.array([-1.2, 1.0]); l = np.full(2, -INF); u = np.full(2, INF)
    oracle.optimize(s, 2, x, l, u, rosen_f, rosen_g)
    assert nrm2(x - [1, 1]) < 1e-8
    s = oracle.settings()
    x[:] = [150.0, 150.0]; l[:] = [10.0, 10.0]; u[:] = [200.0, 200.0]
    oracle.optimize(s, 2, x, l, u, rosen_f, rosen_g)
    assert nrm2(x - [10, 100]) < 1e-5
    assert np.all(x >= 10)


def test_t4_exp_decay(oracle):
    """least_squares.d:333-363"""
    rng = np.random.default_rng(12345)
    xdata = np.linspace(0.0, 10.0, 20)
    model = lambda t, p: p[0] * np.exp(-t * p[1])
    ydata = model(xdata, [1.0, 2.0]) + 0.01 * rng.standard_normal(20)
    x = np.array([0.5, 0.5]); l = np.full(2, -INF); u = np.full(2, INF)

    def f(p, y): y[:] = model(xdata, p) - ydata
    oracle.optimize(oracle.settings(), 20, x, l, u, f)
    assert nrm2(x - [1.0, 2.0]) < 0.05


def test_t5_one_sided_bounds(oracle):
    """least_squares.d:365-411"""
    rng = np.random.default_rng(12345)
    xdata = np.arange(1, 101, dtype=np.float64)
    model = lambda t, p: p[0] * np.exp(-t / p[1]) + p[2]
    ydata = model(xdata, [10.0, 10.0, 10.0]) + 0.1 * rng.standard_normal(100)

    def f(p, y): y[:] = model(xdata, p) - ydata
    x = np.array([15.0, 15.0, 15.0]); l = np.array([5.0, 11.0, 5.0]); u = np.full(3, INF)
    oracle.optimize(oracle.settings(), 100, x, l, u, f)
    assert np.all(x >= l)
    x[:] = [5.0, 5.0, 5.0]; l = np.full(3, -INF); u = np.array([15.0, 9.0, 15.0])
    oracle.optimize(oracle.settings(), 100, x, l, u, f)
    assert np.all(x <= u)


def test_t6_degenerate_m_lt_n(oracle):
    """least_squares.d:413-434"""
    x = np.array([0.001, 0.0001]); l = np.array([-0.5, -0.5]); u = np.array([0.5, 0.5])

    def f(x, y): y[0] = np.sqrt(1 - (x[0] ** 2 + x[1] ** 2))
    oracle.optimize(oracle.settings(), 1, x, l, u, f)
    assert nrm2(x - u) < 1e-8


def test_bq1_boxqp(oracle_lib):
    """boxcqp.d:381-402"""
    P = np.array([[2.0, -1, 0], [-1.0, 2, -1], [0.0, -1, 2]])
    q = np.array([3.0, -7, 5]); l = np.array([-100.0, -2, 1]); u = np.array([100.0, 2, 1])
    x = np.zeros(3)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    it = C.c_uint(0)
    oracle_lib.oracle_solve_box_qp_d.restype = C.c_int
    st = oracle_lib.oracle_solve_box_qp_d(None, C.c_size_t(3), dp(P), dp(q), dp(l), dp(u), dp(x), C.byref(it))
    assert st == 0
    # mir.math.common.approxEqual defaults: maxRelDiff 1e-2, maxAbsDiff 1e-5 -- we hold it much tighter
    np.testing.assert_allclose(x, [-0.5, 2, 1], rtol=1e-12, atol=1e-12)


def test_defaults_and_layout(oracle):
    """sizes, offsets and default values of the PODs."""
    from mir_optim_b200 import _abi
    assert C.sizeof(_abi.LeastSquaresSettingsD) == 128 and C.sizeof(_abi.LeastSquaresSettingsS) == 68
    assert C.sizeof(_abi.LeastSquaresResultD) == 32 and C.sizeof(_abi.LeastSquaresResultS) == 24
    assert C.sizeof(_abi.BoxQPSettingsD) == 24 and C.sizeof(_abi.BoxQPSettingsS) == 12
    assert _abi.LeastSquaresSettingsD.qpSettings.offset == 104 and _abi.LeastSquaresSettingsS.qpSettings.offset == 56
    s = oracle.settings(np.float64)
    eps = np.finfo(np.float64).eps
    assert (s.maxIterations, s.maxAge) == (1000, 0)
    assert s.jacobianEpsilon == 2.0 ** -26 and s.absTolerance == eps and s.relTolerance == 0 and s.gradTolerance == eps
    assert s.maxGoodResidual == eps ** 2
    assert s.maxStep == np.sqrt(np.finfo(np.float64).max) / 16
    assert s.maxLambda == np.finfo(np.float64).max / 16 and s.minLambda == np.finfo(np.float64).tiny * 16
    assert (s.minStepQuality, s.goodStepQuality, s.lambdaIncrease) == (0.1, 0.5, 2.0)
    assert s.lambdaDecrease == pytest.approx(0.3090169943749474, rel=1e-15)
    assert s.qpSettings.relTolerance == 16 * eps and s.qpSettings.absTolerance == 16 * eps and s.qpSettings.maxIterations == 0
    f = oracle.settings(np.float32)
    feps = np.finfo(np.float32).eps
    assert f.jacobianEpsilon == np.float32(2.0 ** -11) and f.absTolerance == feps
    assert f.maxGoodResidual == np.float32(feps) * np.float32(feps)
    assert f.minStepQuality == np.float32(0.1) and f.lambdaDecrease == np.float32(0.3090169943749474)
    assert f.maxLambda == np.float32(np.finfo(np.float32).max / 16)


def test_work_lengths(oracle):
    """least_squares.d:642-656, boxcqp.d:36-50; SURVEY a3 values."""
    L = oracle.lib
    assert L.mir_box_qp_work_length(4) == 2 * 16 + 32
    assert L.mir_box_qp_iwork_length(4) == 5 and L.mir_box_qp_iwork_length(5) == 7
    assert L.mir_least_squares_work_length(1000, 3) == 5066
    assert L.mir_least_squares_work_length(64, 4) == 484
    assert L.mir_least_squares_work_length(128, 8) == 1576
    assert L.mir_least_squares_iwork_length(64, 4) == 5


def test_validation_statuses(oracle):
    """least_squares.d:930-943, first failure wins; `optimize` throws for status < 0 (:175-179)."""
    def f(x, y): y[:] = x[:2]
    base = lambda: (np.array([1.0, 1.0]), np.array([0.0, 0.0]), np.array([2.0, 2.0]))
    S = LeastSquaresStatus
    x, l, u = base(); x[0] = np.nan
    assert oracle.optimize_least_squares(oracle.settings(), 2, x, l, u, f).status == S.badGuess
    x, l, u = base(); x[1] = np.inf
    assert oracle.optimize_least_squares(oracle.settings(), 2, x, l, u, f).status == S.badGuess
    x, l, u = base()
    assert oracle.optimize_least_squares(oracle.settings(), 0, x, l, u, f).status == S.badGuess
    x, l, u = base(); l[0] = 1.5
    assert oracle.optimize_least_squares(oracle.settings(), 2, x, l, u, f).status == S.badBounds
    x, l, u = base(); u[1] = np.nan
    assert oracle.optimize_least_squares(oracle.settings(), 2, x, l, u, f).status == S.badBounds
    for field, val, st in (("minStepQuality", 1.0, S.badMinStepQuality), ("minStepQuality", -0.1, S.badMinStepQuality),
                           ("goodStepQuality", 1.5, S.badGoodStepQuality), ("goodStepQuality", 0.05, S.badStepQuality),
                           ("lambdaIncrease", 0.5, S.badLambdaParams), ("lambdaDecrease", 1.5, S.badLambdaParams),
                           ("lambdaDecrease", 1e-200, S.badLambdaParams)):
        s = oracle.settings(); setattr(s, field, val)
        x, l, u = base()
        r = oracle.optimize_least_squares(s, 2, x, l, u, f)
        assert r.status == st, (field, val)
        assert r.iterations == 0 and r.fCalls == 0 and r.residual == np.inf and r.lambda_ == 0
    s = oracle.settings(); s.lambdaIncrease = 0.5
    x, l, u = base()
    with pytest.raises(LeastSquaresException) as ei:
        oracle.optimize(s, 2, x, l, u, f)
    assert ei.value.status == S.badLambdaParams
    assert "lambdaIncrease" in str(ei.value)


def test_max_iterations_is_status_minus_one(oracle):
    s = oracle.settings(); s.maxIterations = 3
    x = np.array([-1.2, 1.0]); l = np.full(2, -INF); u = np.full(2, INF)
    r = oracle.optimize_least_squares(s, 2, x, l, u, rosen_f, rosen_g)
    assert r.status == LeastSquaresStatus.maxIterations and r.iterations == 3
    with pytest.raises(LeastSquaresException):
        oracle.optimize(s, 2, np.array([-1.2, 1.0]), l, u, rosen_f, rosen_g)


def test_provisional_goldens_counts(oracle):
    """Counts observed by the survey's scratch restatement (last paragraph)."""
    x = np.array([100.0, 100.0]); l = np.full(2, -INF); u = np.full(2, INF)
    def f(x, y): y[0] = x[0]; y[1] = 2 - x[1]
    def g(x, J): J[0, 0] = 1; J[0, 1] = 0; J[1, 0] = 0; J[1, 1] = -1
    r = oracle.optimize(oracle.settings(), 2, x, l, u, f, g)
    assert (r.status, r.iterations, r.fCalls, r.gCalls) == (3, 5, 6, 2)
    x = np.array([150.0, 150.0]); l = np.array([10.0, 10.0]); u = np.array([200.0, 200.0])
    r = oracle.optimize(oracle.settings(), 2, x, l, u, rosen_f, rosen_g)
    assert r.status == LeastSquaresStatus.furtherImprovement and r.residual == pytest.approx(81.0, rel=1e-12)
