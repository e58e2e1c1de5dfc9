"""Small invocations of the round-2 kernels for compute-sanitizer (memcheck / racecheck / synccheck): the four-problems-
per-warp LM kernel (lock-step CTA with the shared-memory task queue, and free-running), the general CTA-per-problem
kernel (built-in functor, spline, NVRTC user model), the warp-per-QP BoxQP kernel.
usage: compute-sanitizer --tool racecheck python scripts/sanitize_round2.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mir_optim_b200 as mo
from mir_optim_b200 import workloads
from mir_optim_b200._abi import ModelId
import user_models
eng = mo.engine
# mux: double / float, FD / analytic, bounds active; 203 problems (lock-step CTA) -- and 30 (free-running single-warp CTAs)
for dt in (np.float64, np.float32):
    for fd in (True, False):
        for B in (203, 30):
            w3 = workloads.c3_sumexp8(B, dtype=dt, m=100); s3 = eng.settings(dt); s3.maxIterations = 4
            l = np.tile(np.array([0.0, 0.0] * 4, dtype=dt), (B, 1)); u = np.tile(np.array([4.0, 12.0] * 4, dtype=dt), (B, 1))
            x = np.clip(w3.x0, l, u).astype(dt); r, st = eng.optimize_batched(s3, w3.model, x, l, u, t=w3.t, y=w3.y, fd_jacobian=fd, want_stats=True)
            print("mux", dt.__name__, fd, B, st["passes"], st["qp_iterations"])
# general kernel: sum of 5 exponentials (n = 10, m = 150), FD and analytic
rng = np.random.default_rng(1)
B, nc, m = 40, 5, 150
t = np.linspace(0, 5, m); a = rng.uniform(1, 5, (B, nc)); b = np.geomspace(0.3, 9, nc)[None] * rng.uniform(0.9, 1.1, (B, nc))
truth = np.empty((B, 2 * nc)); truth[:, 0::2] = a; truth[:, 1::2] = b
y = (a[:, :, None] * np.exp(-b[:, :, None] * t[None, None])).sum(1) + 0.01 * rng.normal(size=(B, m))
s = eng.settings(); s.maxIterations = 5
for fd in (True, False):
    x = truth * rng.uniform(0.9, 1.1, truth.shape)
    r, st = eng.optimize_batched(s, ModelId.SUMEXP, x, np.full(2 * nc, 0.0), np.full(2 * nc, 20.0), t=t, y=y, fd_jacobian=fd, want_stats=True)
    print("cta sumexp10", fd, st["passes"], st["qp_solves"])
# fitSpline, batched
knots = np.cumsum(rng.uniform(0.5, 1.5, 9)); px = np.sort(rng.uniform(knots[0], knots[-1], (24, 40)), axis=1); py = np.sin(px) + 0.05 * rng.normal(size=px.shape)
for lam in (0.0, 1e-2):
    v, r = eng.fit_spline_batched(s, px, py, knots, np.full(9, -0.9), np.full(9, 0.9), lam)
    print("spline", lam, int(np.sum(r["status"] >= -1)))
# NVRTC user model
mid = eng.compile_model(user_models.LOGISTIC)
tt = np.linspace(0, 12, 60); yy = 8.0 / (1 + np.exp(-0.9 * (tt - 6.0))) + 0.05 * rng.normal(size=(16, 60))
x = np.tile(np.array([7.0, 1.0, 5.0]), (16, 1)); r, _ = eng.optimize_batched(s, mid, x, np.zeros(3), np.array([9.0, 5.0, 20.0]), t=tt, y=yy, fd_jacobian=True)
print("user model", int(np.sum(r["status"] >= -1)))
# warp-per-QP BoxQP: n = 64, 37 (odd), 5
for n in (64, 37, 5):
    w5 = workloads.c5_boxqp(96, n=n, seed=n)
    xq, stq, it = eng.solve_box_qp_batched(w5.P, w5.q, w5.l, w5.u); print("boxqp warp", n, int(np.sum(stq == 0)), float(it.mean()))
