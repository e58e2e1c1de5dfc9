"""Residual models supplied as CUDA source at run time (the reference takes arbitrary f / g, LS:78-80):
compiled with NVRTC into the general batched kernel, checked against the built-in functor they restate (bit for bit),
against the CPU oracle driven by an ordinary host callback (the reference's own calling convention), and through the
reference's legacy entry point."""
import numpy as np
import pytest

from mir_optim_b200._abi import ModelId
from oracle_util import rel_err
import user_models

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import mir_optim_b200
    assert mir_optim_b200.engine.device_count() > 0, "no CUDA device"
    return mir_optim_b200.engine


def decay_data(B, m, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    t = np.linspace(0, 10, m)
    truth = np.stack([rng.uniform(2, 4, B), rng.uniform(0.4, 1.0, B), rng.uniform(0.5, 1.5, B)], axis=1)
    y = truth[:, 0, None] * np.exp(-truth[:, 1, None] * t[None, :]) + truth[:, 2, None] + 0.02 * rng.normal(size=(B, m))
    x0 = truth * rng.uniform(0.8, 1.2, truth.shape)
    return t.astype(dtype), y.astype(dtype), x0.astype(dtype), truth


@pytest.mark.parametrize("fd", [False, True])
def test_user_model_matches_the_builtin_it_restates(eng, fd, monkeypatch):
    mid = eng.compile_model(user_models.EXPDECAY3_REPRO)
    t, y, x0, truth = decay_data(300, 200, 1)                   # m = 200: the built-in goes through the general kernel too
    l = np.full(3, -np.inf); u = np.full(3, np.inf)
    # Same source through NVRTC and through nvcc: the two may contract different multiply-add pairs of the kernel's own
    # (not explicitly rounded) arithmetic.  k-step runs: identical decisions, x to rounding; full runs on noisy data fork
    # like any two 1-ulp-different runs of the reference algorithm, the fitted x agrees to ~1e-8.
    for k in (3, 0):
        s = eng.settings(np.float64)
        if k:
            s.maxIterations = k
        xa = x0.copy(); ra, _ = eng.optimize_batched(s, mid, xa, l, u, t=t, y=y, fd_jacobian=fd)
        xb = x0.copy(); rb, _ = eng.optimize_batched(s, ModelId.EXPDECAY3, xb, l, u, t=t, y=y, fd_jacobian=fd)
        if k:
            same = np.ones(len(xa), bool)
            for key in ("status", "iterations", "fCalls", "gCalls"):
                same &= ra[key] == rb[key]
            assert same.mean() >= 0.99, float(same.mean())
            assert np.max(rel_err(xa[same], xb[same])) < (1e-9 if fd else 1e-12)
        else:
            assert np.all(ra["status"] >= 0) and np.all(rb["status"] >= 0)
            assert np.quantile(rel_err(xa, xb), 0.99) < 1e-6 and np.max(rel_err(ra["residual"], rb["residual"])) < 1e-9
            assert np.max(rel_err(xa, truth)) < 0.2
    eng.release_model(mid)


def test_user_model_float(eng):
    mid = eng.compile_model(user_models.EXPDECAY3_REPRO)
    t, y, x0, truth = decay_data(64, 100, 2, np.float32)
    l = np.full(3, -np.inf, np.float32); u = np.full(3, np.inf, np.float32)
    s = eng.settings(np.float32); s.maxIterations = 3
    xa = x0.copy(); ra, _ = eng.optimize_batched(s, mid, xa, l, u, t=t, y=y)
    xb = x0.copy(); rb, _ = eng.optimize_batched(s, ModelId.EXPDECAY3, xb, l, u, t=t, y=y)       # (m = 100: the lane-group kernel)
    same = (ra["status"] == rb["status"]) & (ra["iterations"] == rb["iterations"])
    assert same.mean() > 0.9 and np.max(rel_err(xa[same], xb[same])) < 1e-4
    s = eng.settings(np.float32)
    xa = x0.copy(); ra, _ = eng.optimize_batched(s, mid, xa, l, u, t=t, y=y)
    assert np.all(ra["status"] >= 0) and np.max(rel_err(xa, truth)) < 0.2
    eng.release_model(mid)


def test_logistic_model_against_the_oracle_with_a_host_callback(eng, oracle):
    """A model the library does not ship, finite-difference Jacobian, box bounds: the GPU runs the CUDA source, the CPU
    oracle runs the reference algorithm with the same formula as an ordinary host callback (LS:78)."""
    mid = eng.compile_model(user_models.LOGISTIC)
    rng = np.random.default_rng(3)
    B, m = 40, 80
    t = np.linspace(0, 12, m)
    truth = np.stack([rng.uniform(5, 10, B), rng.uniform(0.6, 1.2, B), rng.uniform(4, 8, B)], axis=1)
    y = truth[:, 0, None] / (1 + np.exp(-truth[:, 1, None] * (t[None, :] - truth[:, 2, None]))) + 0.05 * rng.normal(size=(B, m))
    x0 = truth * rng.uniform(0.85, 1.15, truth.shape)
    l = np.array([0.0, 0.0, 0.0]); u = np.array([9.0, 5.0, 20.0])            # K capped at 9: active for some problems
    x0 = np.clip(x0, l, u)
    for k in (2, 0):
        s = eng.settings(np.float64)
        if k:
            s.maxIterations = k
        xg = x0.copy(); rg, _ = eng.optimize_batched(s, mid, xg, l, u, t=t, y=y, fd_jacobian=True)
        xo = x0.copy(); ro = []
        for b in range(B):
            def f(p, out, b=b):
                out[:] = p[0] / (1 + np.exp(-p[1] * (t - p[2]))) - y[b]
            ro.append(oracle.optimize_least_squares(s, m, xo[b], l, u, f))
        st = np.array([r.status for r in ro]); it = np.array([r.iterations for r in ro]); res = np.array([r.residual for r in ro])
        if k:
            assert np.mean((rg["status"] == st) & (rg["iterations"] == it)) > 0.95
            assert np.max(rel_err(xg, xo)) < 1e-6           # (libm's exp vs CUDA's exp, amplified by the finite differences)
        else:
            assert np.all(rg["status"] >= 0) and np.all(st >= 0)
            assert np.quantile(rel_err(xg, xo), 0.9) < 1e-5 and np.max(rel_err(rg["residual"], res)) < 1e-6
            assert np.any(xg[:, 0] == 9.0) and np.all(xg <= u) and np.all(xg >= l)
    eng.release_model(mid)


def test_aux_param_and_the_legacy_entry_point(eng, oracle):
    """aux / param reach the user's functor; the reference's own signature (mir_optimize_least_squares_d) runs a user model
    on the GPU when fContext carries its id (device-model mode)."""
    mid = eng.compile_model(user_models.RIDGE_POLY)
    rng = np.random.default_rng(5)
    n, P = 6, 50
    m = P + n
    t = np.zeros(m); y = np.zeros(m)
    t[:P] = np.linspace(-1, 1, P); coef = rng.normal(size=n); y[:P] = np.polyval(coef[::-1], t[:P]) + 0.01 * rng.normal(size=P)
    w = rng.uniform(0.5, 2.0, n); lam = 0.3
    l = np.full(n, -np.inf); u = np.full(n, np.inf)
    s = eng.settings(np.float64)
    x = np.zeros(n)
    r = eng.optimize_device_model(s, mid, x, l, u, t=t, y=y, aux=w, param=lam)
    assert r.status >= 0
    # closed form of the ridge problem
    A = np.vander(t[:P], n, increasing=True)
    xs = np.linalg.solve(A.T @ A + lam * np.diag(w ** 2), A.T @ y[:P])
    assert np.max(np.abs(x - xs)) < 1e-8
    # and the batched entry with per-problem aux
    B = 16
    W = rng.uniform(0.5, 2.0, (B, n)); X = np.zeros((B, n))
    rg, _ = eng.optimize_batched(s, mid, X, l, u, t=t, y=np.tile(y, (B, 1)), aux=W, param=lam)
    for b in range(B):
        xb = np.linalg.solve(A.T @ A + lam * np.diag(W[b] ** 2), A.T @ y[:P])
        assert np.max(np.abs(X[b] - xb)) < 1e-8
    eng.release_model(mid)
