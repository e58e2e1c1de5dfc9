"""CPU-only sanity of the oracle's batched driver and built-in model callbacks: the C model
functions must agree with straightforward numpy evaluations, analytic Jacobians with finite
differences, and fits must recover the generating parameters."""
import ctypes as C

import numpy as np
import pytest

from mir_optim_b200 import workloads
from mir_optim_b200._abi import ModelId
from mir_optim_b200.api import ReferenceAPI
from oracle_util import oracle_batched


class Ctx(C.Structure):
    _fields_ = [("model", C.c_int), ("t", C.c_void_p), ("y", C.c_void_p)]


def call_model(lib, model, p, t, y, m):
    n = len(p)
    r = np.zeros(m); J = np.zeros((m, n))
    ctx = Ctx(int(model), t.ctypes.data if t is not None else None, y.ctypes.data if y is not None else None)
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    for name, out in (("oracle_model_f_d", r), ("oracle_model_g_d", J)):
        fn = getattr(lib, name); fn.restype = None
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
        fn(C.addressof(ctx), m, n, dp(p), dp(out))
    return r, J


@pytest.mark.parametrize("model,n,m", [(ModelId.EXPDECAY2, 2, 20), (ModelId.EXPTAU3, 3, 50), (ModelId.EXPDECAY3, 3, 40),
                                       (ModelId.GAUSS4, 4, 64), (ModelId.SUMEXP, 8, 128), (ModelId.GAUSSMIX, 8, 100)])
def test_model_jacobians_match_finite_differences(oracle_lib, model, n, m):
    rng = np.random.default_rng(7)
    t = np.linspace(0.1, 3.0, m); y = rng.standard_normal(m)
    p = rng.uniform(0.5, 1.5, n)
    r, J = call_model(oracle_lib, model, p, t, y, m)
    h = 1e-6
    for k in range(n):
        pp = p.copy(); pm = p.copy(); pp[k] += h; pm[k] -= h
        rp, _ = call_model(oracle_lib, model, pp, t, y, m); rm, _ = call_model(oracle_lib, model, pm, t, y, m)
        np.testing.assert_allclose(J[:, k], (rp - rm) / (2 * h), rtol=2e-6, atol=2e-8)


def test_gauss4_residual_formula(oracle_lib):
    t = np.linspace(-4, 4, 64); y = np.zeros(64); p = np.array([2.0, 0.3, 0.8, 0.5])
    r, _ = call_model(oracle_lib, ModelId.GAUSS4, p, t, y, 64)
    np.testing.assert_allclose(r, p[0] * np.exp(-0.5 * ((t - p[1]) / p[2]) ** 2) + p[3], rtol=1e-14)


def test_batched_oracle_recovers_truth(oracle_lib, oracle):
    wl = workloads.c2_gauss4(64, rel_noise=1e-6)
    x, res, _ = oracle_batched(oracle_lib, oracle.settings(), wl.model, wl.x0, wl.l, wl.u, t=wl.t, y=wl.y)
    assert np.all(res["status"] >= 0)
    inside = np.all((wl.truth > wl.l + 1e-3) & (wl.truth < wl.u - 1e-3), axis=1)   # ~10 % of peaks sit on a bound by design
    assert inside.sum() > 40
    assert np.max(np.abs(x - wl.truth)[inside]) < 1e-4
    assert np.all(x >= wl.l) and np.all(x <= wl.u)
    wl = workloads.c3_sumexp8(16, noise=1e-7)
    x, res, _ = oracle_batched(oracle_lib, oracle.settings(), wl.model, wl.x0, wl.l, wl.u, t=wl.t, y=wl.y, fd_jacobian=True)
    assert np.all(res["status"] >= 0)
    assert np.median(np.abs(x - wl.truth)) < 1e-2


def test_batched_equals_single_calls(oracle_lib, oracle):
    """The OpenMP driver must return exactly what one extern(C) call per problem returns."""
    wl = workloads.c2_gauss4(8, noise=0.05)
    xb, rb, _ = oracle_batched(oracle_lib, oracle.settings(), wl.model, wl.x0, wl.l, wl.u, t=wl.t, y=wl.y, nthreads=2)
    for b in range(8):
        x = wl.x0[b].copy()
        ctx = Ctx(int(wl.model), wl.t.ctypes.data, wl.y[b].ctypes.data)
        L = oracle.lib
        wlen = L.mir_least_squares_work_length(wl.m, wl.n); iwlen = L.mir_least_squares_iwork_length(wl.m, wl.n)
        work = np.empty(wlen); iwork = np.empty(iwlen, dtype=np.int32)
        from mir_optim_b200 import _abi
        f = C.cast(oracle_lib.oracle_model_f_d, C.c_void_p); g = C.cast(oracle_lib.oracle_model_g_d, C.c_void_p)
        s = oracle.settings()
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        r = L.mir_optimize_least_squares_d(C.byref(s), wl.m, wl.n, dp(x), dp(wl.l), dp(wl.u), _abi.SliceD(wlen, dp(work)),
                                           _abi.SliceI(iwlen, iwork.ctypes.data_as(C.POINTER(C.c_int32))),
                                           C.addressof(ctx), f, C.addressof(ctx), g, None, None)
        assert np.array_equal(x, xb[b])
        assert (r.status, r.iterations, r.fCalls, r.gCalls, r.residual, r.lambda_) == tuple(rb[b])
