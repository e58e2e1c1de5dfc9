"""GPU parity of the batched warp-per-problem LM kernel against the CPU oracle, following the
protocol of SURVEY section 8c (P0-P5).  Every call goes through the C ABI
(mir_optimize_least_squares_batched_{d,s}).  Tolerances: 1e-10 relative (double) and 1e-4 (float)
on parameters and residual, as BASELINE.json's north_star states; trajectory-prefix tests are
held tighter.  Measured maxima are appended to gpurun_out/parity_report.jsonl.
"""
import json
import os

import numpy as np
import pytest

from mir_optim_b200 import workloads
from mir_optim_b200._abi import LeastSquaresStatus as S, ModelId
from oracle_util import oracle_batched, rel_err, self_sensitivity, assert_within_conditioning

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eng():
    import mir_optim_b200
    assert mir_optim_b200.engine.device_count() > 0, "no CUDA device"
    return mir_optim_b200.engine


def report(name, **kv):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"), "a") as f:
        f.write(json.dumps({"test": name, **{k: (float(v) if isinstance(v, (np.floating, float)) else v) for k, v in kv.items()}}) + "\n")


def run_both(eng, oracle_lib, wl, settings_mut=None, dtype=np.float64, fd=None):
    fd = wl.fd_jacobian if fd is None else fd
    sg = eng.settings(dtype); so = eng.settings(dtype)
    if settings_mut:
        settings_mut(sg); settings_mut(so)
    xg = wl.x0.astype(dtype).copy()
    rg, stats = eng.optimize_batched(sg, wl.model, xg, wl.l.astype(dtype), wl.u.astype(dtype), t=wl.t.astype(dtype),
                                     y=wl.y.astype(dtype), fd_jacobian=fd, want_stats=True)
    xo, ro, _ = oracle_batched(oracle_lib, so, wl.model, wl.x0.astype(dtype), wl.l.astype(dtype), wl.u.astype(dtype),
                               t=wl.t.astype(dtype), y=wl.y.astype(dtype), fd_jacobian=fd)
    return xg, rg, xo, ro, stats


# ---- P0: argument validation statuses are exact -------------------------------------------------
def test_p0_validation_statuses(eng, oracle_lib):
    wl = workloads.c2_gauss4(8, noise=0.05)
    x0 = wl.x0.copy()
    x0[1, 0] = np.nan; x0[2, 3] = np.inf; x0[3, 1] = 5.0     # badGuess, badGuess, badBounds
    l = np.tile(wl.l, (8, 1)); u = np.tile(wl.u, (8, 1))
    l[4, 2] = np.nan                                          # badBounds (NaN bound)
    s = eng.settings()
    xg = x0.copy()
    rg, _ = eng.optimize_batched(s, wl.model, xg, l, u, t=wl.t, y=wl.y)
    xo, ro, _ = oracle_batched(oracle_lib, eng.settings(), wl.model, x0, l, u, t=wl.t, y=wl.y)
    assert list(rg["status"][1:5]) == [S.badGuess, S.badGuess, S.badBounds, S.badBounds]
    assert np.array_equal(rg["status"] < -1, ro["status"] < -1) and np.array_equal(rg["status"][1:5], ro["status"][1:5])
    for b in (1, 2, 3, 4):      # rejected problems: Result.init fields, x untouched
        assert rg["iterations"][b] == 0 and rg["fCalls"][b] == 0 and np.isinf(rg["residual"][b]) and rg["lambda"][b] == 0
        assert np.array_equal(xg[b], x0[b], equal_nan=True)
    good = [0, 5, 6, 7]                                       # neighbours of failing problems are unaffected
    assert np.all(rg["status"][good] >= 0)
    assert np.max(rel_err(xg[good], xo[good])) < 1e-6
    for field, val, st in (("minStepQuality", 1.0, S.badMinStepQuality), ("goodStepQuality", 1.5, S.badGoodStepQuality),
                           ("goodStepQuality", 0.05, S.badStepQuality), ("lambdaIncrease", 0.5, S.badLambdaParams),
                           ("lambdaDecrease", 1e-200, S.badLambdaParams)):
        s = eng.settings(); setattr(s, field, val)
        rg, _ = eng.optimize_batched(s, wl.model, wl.x0.copy(), wl.l, wl.u, t=wl.t, y=wl.y)
        assert np.all(rg["status"] == st), field


# ---- P1: k-step trajectory parity ---------------------------------------------------------------
@pytest.mark.parametrize("config,fd", [("c2", False), ("c2", True), ("c3", True)])
def test_p1_k_step_trajectories_double(eng, oracle_lib, config, fd):
    """maxIterations = k: after k accepted steps both sides must sit on the same trajectory point.
    Analytic Jacobian: <= 1e-12 up to k = 4.  Finite differences divide rounding noise of the residuals
    by 2*jacobianEpsilon = 3e-8, so the (bit-identical-model) FD runs are held to 1e-9.  From k = 6 a
    rounding-level accept/reject decision may fork a trajectory: counted, must stay < 1 %."""
    wl = workloads.c2_gauss4(512, noise=0.05) if config == "c2" else workloads.c3_sumexp8(256, noise=0.01)
    worst = {}
    for k in (1, 2, 3, 4, 6):
        def mut(s, k=k): s.maxIterations = k
        xg, rg, xo, ro, _ = run_both(eng, oracle_lib, wl, mut, fd=fd)
        same = ((rg["status"] == ro["status"]) & (rg["iterations"] == ro["iterations"]) & (rg["fCalls"] == ro["fCalls"])
                & (rg["gCalls"] == ro["gCalls"]))
        if k <= 4:
            assert same.all(), f"k={k}: {np.where(~same)[0][:5]}"
        assert same.mean() >= 0.99, f"k={k}"
        ex = np.max(rel_err(xg[same], xo[same])); er = np.max(rel_err(rg["residual"][same], ro["residual"][same]))
        el = np.max(rel_err(rg["lambda"][same], ro["lambda"][same]))
        worst[k] = (ex, er, el, float(same.mean()))
        tol = (1e-9 if fd else 1e-12) * (1 if k <= 4 else 20)
        assert ex < tol and er < tol * 20 and el < 1e-12, (k, ex, er, el)
    report(f"p1_{config}_fd{int(fd)}_double", **{f"k{k}": list(map(float, v)) for k, v in worst.items()})


def test_p1_k_step_trajectories_float(eng, oracle_lib):
    wl = workloads.c2_gauss4(256, noise=0.05)
    agree = {}
    for k in (1, 2, 3):
        def mut(s, k=k): s.maxIterations = k
        xg, rg, xo, ro, _ = run_both(eng, oracle_lib, wl, mut, dtype=np.float32)
        same = (rg["status"] == ro["status"]) & (rg["fCalls"] == ro["fCalls"])
        agree[k] = float(same.mean())
        assert same.mean() > 0.97                  # float trajectories may fork on a rounding-level accept/reject
        ex = np.max(rel_err(xg[same], xo[same])); er = np.max(rel_err(rg["residual"][same], ro["residual"][same]))
        assert ex < 1e-4 and er < 1e-4, (k, ex, er)
    report("p1_c2_float", agree=agree)


# ---- P2: low-noise fits with default settings -----------------------------------------------------
def test_p2_low_noise_defaults_double(eng, oracle_lib):
    """Default settings run to the reference's own termination (the lambda-overflow tail).  The residual must
    agree to 1e-10.  Final parameters are compared against the reference's own conditioning: the oracle re-run on
    inputs perturbed by one ulp moves its parameters by up to ~1e-8, so that spread -- not 1e-10 --
    is the resolvable accuracy of a full run; the GPU must sit inside it (and the median well below 1e-10)."""
    wl = workloads.c2_gauss4(4096, rel_noise=1e-4)
    xg, rg, xo, ro, stats = run_both(eng, oracle_lib, wl)
    assert np.all(rg["status"] >= 0) and np.all(ro["status"] >= 0)
    _, _, sens_x, sens_r = self_sensitivity(oracle_lib, eng.settings(), wl)
    ex = rel_err(xg, xo); er = rel_err(rg["residual"], ro["residual"])
    report("p2_c2_double", max_x=ex.max(), median_x=float(np.median(ex)), q99_x=float(np.quantile(ex, 0.99)), max_res=er.max(),
           oracle_1ulp_sensitivity_max_x=sens_x.max(), oracle_1ulp_sensitivity_q99_x=float(np.quantile(sens_x, 0.99)),
           same_status=float((rg["status"] == ro["status"]).mean()), passes_per_fit=stats["passes"] / 4096)
    assert er.max() < 1e-10
    assert np.median(ex) < 1e-10
    assert_within_conditioning(ex, sens_x)


def test_p2_low_noise_defaults_float(eng, oracle_lib):
    wl = workloads.c2_gauss4(4096, rel_noise=1e-3)
    xg, rg, xo, ro, _ = run_both(eng, oracle_lib, wl, dtype=np.float32)
    assert np.all(rg["status"] >= 0)
    ex = rel_err(xg, xo); er = rel_err(rg["residual"], ro["residual"])
    report("p2_c2_float", max_x=ex.max(), max_res=er.max(), p99_x=float(np.quantile(ex, 0.99)))
    assert np.quantile(ex, 0.99) < 1e-4 and np.quantile(er, 0.99) < 1e-4
    assert ex.max() < 2e-3            # float-vs-float oracle; a few fits fork at rounding level


# ---- P3: noise-free data, robust termination threshold --------------------------------------------
@pytest.mark.parametrize("config", ["c2", "c3"])
def test_p3_noise_free_fconverged(eng, oracle_lib, config):
    if config == "c2":
        wl = workloads.c2_gauss4(2048, noise=0.0)
        wl["l"] = np.array([0.0, -2.0, 0.3, -1.0]); wl["u"] = np.array([20.0, 2.0, 2.0, 2.0])   # truth strictly inside
    else:
        wl = workloads.c3_sumexp8(512, noise=0.0)
    def mut(s): s.maxGoodResidual = 1e-20
    xg, rg, xo, ro, _ = run_both(eng, oracle_lib, wl, mut)
    ok = ro["status"] == S.fConverged
    assert ok.mean() > (0.99 if config == "c2" else 0.5)
    same = float((rg["status"] == ro["status"]).mean())
    ex = rel_err(xg[ok], xo[ok])
    report(f"p3_{config}", same_status=same, max_x=ex.max(), frac_fconverged=float(ok.mean()))
    assert np.array_equal(rg["status"][ok], ro["status"][ok])
    assert ex.max() < (1e-10 if config == "c2" else 1e-6)


# ---- P4: realistic noise -------------------------------------------------------------------------
def test_p4_realistic_noise(eng, oracle_lib):
    wl = workloads.c2_gauss4(4096, noise=0.05)
    xg, rg, xo, ro, stats = run_both(eng, oracle_lib, wl)
    assert np.all(rg["status"] >= 0)
    er = rel_err(rg["residual"], ro["residual"]); ex = rel_err(xg, xo)
    same = float((rg["status"] == ro["status"]).mean())
    report("p4_c2_double", max_res=er.max(), max_x=ex.max(), same_status=same, passes_per_fit=stats["passes"] / 4096,
           oracle_mean_iterations=float(ro["iterations"].mean()), gpu_mean_iterations=float(rg["iterations"].mean()))
    _, _, sens_x, _ = self_sensitivity(oracle_lib, eng.settings(), wl)
    assert er.max() < 1e-10
    assert_within_conditioning(ex, sens_x)
    assert np.all(np.isin(rg["status"], (S.furtherImprovement, S.xConverged, S.gConverged, S.fConverged)))
    assert np.all(xg >= wl.l) and np.all(xg <= wl.u)


def test_p4_c3_fd_realistic(eng, oracle_lib):
    wl = workloads.c3_sumexp8(1024, noise=0.01)
    xg, rg, xo, ro, stats = run_both(eng, oracle_lib, wl)
    # (a fit that needs ~1000 accepted steps may end as maxIterations on either side: the oracle itself has one here)
    assert np.mean(rg["status"] < 0) <= 0.005 and np.all(rg["status"] >= -1), (np.unique(rg["status"], return_counts=True), np.unique(ro["status"], return_counts=True))
    er = rel_err(rg["residual"], ro["residual"])
    report("p4_c3_double", max_res=er.max(), median_res=float(np.median(er)), passes_per_fit=stats["passes"] / 1024,
           same_status=float((rg["status"] == ro["status"]).mean()))
    assert np.quantile(er, 0.95) < 1e-8          # ill-conditioned sum-of-exponentials: residuals agree, parameters wander


def test_c5b_active_bounds(eng, oracle_lib):
    """configs[4]b: LM with at least one bound active at the solution (BoxQP active-set iterations)."""
    wl = workloads.c2_gauss4(2048, rel_noise=1e-4, tight_bounds=True)
    xg, rg, xo, ro, stats = run_both(eng, oracle_lib, wl)
    on_bound = np.any((xo == wl.l) | (xo == wl.u), axis=1)
    assert on_bound.mean() > 0.5 and stats["qp_iterations"] > 0
    assert np.all(xg >= wl.l) and np.all(xg <= wl.u)
    ex = rel_err(xg, xo); er = rel_err(rg["residual"], ro["residual"])
    report("c5b_active_bounds", max_x=ex.max(), max_res=er.max(), frac_on_bound=float(on_bound.mean()),
           same_status=float((rg["status"] == ro["status"]).mean()), qp_iterations=stats["qp_iterations"])
    assert np.array_equal((xg == wl.l) | (xg == wl.u), (xo == wl.l) | (xo == wl.u))
    _, _, sens_x, _ = self_sensitivity(oracle_lib, eng.settings(), wl)
    assert er.max() < 1e-10
    assert_within_conditioning(ex, sens_x)


# ---- P5: the reference's own unit tests, on the GPU -------------------------------------------------
def test_p5_reference_unit_tests_on_gpu(eng):
    INF = np.inf
    s = eng.settings()
    x = np.array([[100.0, 100.0]])                                   # T1, least_squares.d:217-245
    r, _ = eng.optimize_batched(s, ModelId.LINEAR2, x, np.full(2, -INF), np.full(2, INF), m=2)
    assert np.linalg.norm(x[0] - [0, 2]) < 1e-8
    assert tuple(r[0])[:4] == (S.fConverged, 5, 6, 2)
    x = np.array([[-1.2, 1.0]])                                      # T2 (FD), :247-273
    r, _ = eng.optimize_batched(s, ModelId.ROSENBROCK, x, np.full(2, -INF), np.full(2, INF), m=2, fd_jacobian=True)
    assert np.linalg.norm(x[0] - [1, 1]) < 1e-6 and r["status"][0] >= 0
    x = np.array([[-1.2, 1.0]])                                      # T3a analytic, :275-317
    r, _ = eng.optimize_batched(s, ModelId.ROSENBROCK, x, np.full(2, -INF), np.full(2, INF), m=2)
    assert np.linalg.norm(x[0] - [1, 1]) < 1e-8
    x = np.array([[150.0, 150.0]])                                   # T3b box, :321-330
    r, _ = eng.optimize_batched(s, ModelId.ROSENBROCK, x, np.array([10.0, 10.0]), np.array([200.0, 200.0]), m=2)
    assert np.linalg.norm(x[0] - [10, 100]) < 1e-5 and np.all(x >= 10)
    assert r["status"][0] == S.furtherImprovement and abs(r["residual"][0] - 81.0) < 1e-9
    rng = np.random.default_rng(12345)                               # T4, :333-363
    t = np.linspace(0.0, 10.0, 20)
    y = (1.0 * np.exp(-t * 2.0) + 0.01 * rng.standard_normal(20))[None, :]
    x = np.array([[0.5, 0.5]])
    eng.optimize_batched(s, ModelId.EXPDECAY2, x, np.full(2, -INF), np.full(2, INF), t=t, y=y, fd_jacobian=True)
    assert np.linalg.norm(x[0] - [1.0, 2.0]) < 0.05
    rng = np.random.default_rng(12345)                               # T5, :365-411
    t = np.arange(1, 101, dtype=np.float64)
    y = (10.0 * np.exp(-t / 10.0) + 10.0 + 0.1 * rng.standard_normal(100))[None, :]
    x = np.array([[15.0, 15.0, 15.0]]); l = np.array([5.0, 11.0, 5.0])
    eng.optimize_batched(s, ModelId.EXPTAU3, x, l, np.full(3, INF), t=t, y=y, fd_jacobian=True)
    assert np.all(x[0] >= l)
    x = np.array([[5.0, 5.0, 5.0]]); u = np.array([15.0, 9.0, 15.0])
    eng.optimize_batched(s, ModelId.EXPTAU3, x, np.full(3, -INF), u, t=t, y=y, fd_jacobian=True)
    assert np.all(x[0] <= u)
    x = np.array([[0.001, 0.0001]]); u = np.array([0.5, 0.5])        # T6, :413-434
    eng.optimize_batched(s, ModelId.SQRTCIRCLE, x, -u, u, m=1, fd_jacobian=True)
    assert np.linalg.norm(x[0] - u) < 1e-8


# ---- full-size properties (BASELINE configs[1]: 2^20 fits) -------------------------------------------
def test_full_size_properties_c2(eng, oracle_lib):
    B = 1 << 20
    wl = workloads.c2_gauss4(B, noise=0.05)
    x = wl.x0.copy()
    res, stats = eng.optimize_batched(eng.settings(), wl.model, x, wl.l, wl.u, t=wl.t, y=wl.y, want_stats=True)
    assert stats["problems"] == B and stats["accepted"] == int(res["iterations"].sum())
    assert np.all(res["status"] >= 0)
    assert np.all(x >= wl.l) and np.all(x <= wl.u) and np.all(np.isfinite(x))
    z0 = (wl.t[None, :] - wl.x0[:, 1:2]) / wl.x0[:, 2:3]
    r0 = np.sum((wl.x0[:, 0:1] * np.exp(-0.5 * z0 * z0) + wl.x0[:, 3:4] - wl.y) ** 2, axis=1)
    assert np.all(res["residual"] <= r0 * (1 + 1e-12))              # LM never accepts an uphill step
    z = (wl.t[None, :] - x[:, 1:2]) / x[:, 2:3]
    r1 = np.sum((x[:, 0:1] * np.exp(-0.5 * z * z) + x[:, 3:4] - wl.y) ** 2, axis=1)
    np.testing.assert_allclose(res["residual"], r1, rtol=1e-10)     # reported residual is ||f(x_out)||^2
    idx = np.random.default_rng(0).choice(B, 32768, replace=False)  # idempotence of batching: sub-batch == full batch (same kernel: >= 16384)
    xs = wl.x0[idx].copy()
    rs, _ = eng.optimize_batched(eng.settings(), wl.model, xs, wl.l, wl.u, t=wl.t, y=wl.y[idx])
    assert np.array_equal(xs, x[idx]) and np.array_equal(rs, res[idx])
    xo, ro, _ = oracle_batched(oracle_lib, eng.settings(), wl.model, wl.x0[idx], wl.l, wl.u, t=wl.t, y=wl.y[idx])
    er = rel_err(res["residual"][idx], ro["residual"])
    assert np.quantile(er, 0.999) < 1e-10 and er.max() < 1e-8      # ||r||^2 <= 1e-10 (rare fits land 1e-10..1e-9 apart)
    report("full_c2", passes_per_fit=stats["passes"] / B, accepted_per_fit=stats["accepted"] / B,
           evals_per_fit=stats["model_evals"] / B, qp_solves_per_fit=stats["qp_solves"] / B)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["c2", "c2_tight", "c2_fd", "c3", "c2_float"])
def test_tail_fast_forward_is_bit_identical(case):
    """The lambda-overflow tail fast-forward (lm_small.cuh, tail_is_inert) must not change a single bit of x, status,
    iterations, fCalls, gCalls, residual or the final lambda relative to executing every pass (MIR_MODEL_NO_TAIL_SHORTCUT)."""
    import mir_optim_b200 as mo
    from mir_optim_b200 import workloads
    eng = mo.engine
    dt = np.float32 if case.endswith("float") else np.float64
    if case.startswith("c2"):
        wl = workloads.c2_gauss4(4096, dtype=dt, noise=0.05, tight_bounds=(case == "c2_tight"))
        fd = case == "c2_fd"
    else:
        wl = workloads.c3_sumexp8(2048, dtype=dt); fd = True
    s = eng.settings(dt)
    out = []
    for shortcut in (True, False):
        x = wl.x0.copy()
        res, stats = eng.optimize_batched(s, wl.model, x, wl.l, wl.u, t=wl.t, y=wl.y, fd_jacobian=fd, want_stats=True, tail_shortcut=shortcut)
        out.append((x, res, stats))
    (xa, ra, sa), (xb, rb, sb) = out
    assert np.array_equal(xa.view(np.uint8), xb.view(np.uint8))
    assert ra.tobytes() == rb.tobytes()
    assert sa["passes"] == sb["passes"] and sa["qp_solves"] < sb["qp_solves"]
    frac_tail = np.mean(ra["status"] == 0)
    assert frac_tail > 0.3, "workload should exercise the lambda-overflow exit"
