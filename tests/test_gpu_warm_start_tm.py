"""SURVEY 8f-3 and 8f-4 on the GPU path.

  * 8f-3  the finite-difference Jacobian through the C thread-manager interface (LS:567-578, 837-854) with a manager that
          really runs the tasks CONCURRENTLY: the tasks of this library use private scratch (the reference shares mBuffer
          between them, LS:1036-1041, a race), so any schedule gives the bits of the serial loop.
  * 8f-4  warm start: Result.lambda of a previous call as the initial damping (documented at LS:141-142, never read back by
          the reference, LS:966).
"""
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from mir_optim_b200 import workloads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import mir_optim_b200 as mo
    assert mo.engine.device_count() > 0
    return mo.engine


def test_fd_jacobian_with_a_concurrent_thread_manager(eng, oracle):
    rng = np.random.default_rng(2)
    m, n = 400, 6
    t = np.linspace(0, 5, m)
    truth = np.array([3.0, 0.4, 2.0, 1.5, 1.0, 6.0])
    ydata = sum(truth[2 * k] * np.exp(-truth[2 * k + 1] * t) for k in range(3)) + 0.01 * rng.normal(size=m)

    def f(p, out):
        out[:] = sum(p[2 * k] * np.exp(-p[2 * k + 1] * t) for k in range(3)) - ydata

    seen = set(); lock = threading.Lock()
    pool = ThreadPoolExecutor(4)

    def tm_parallel(count, call):
        gate = threading.Barrier(4)                        # four tasks in flight at once, on four threads

        def work(tid):
            gate.wait(timeout=30)
            for i in range(tid, count, 4):                 # the manager's contract: every i in [0, count) exactly once (LS:672-678)
                with lock:
                    seen.add(threading.get_ident())
                call(4, tid, i)
        list(pool.map(work, range(4)))

    def tm_serial(count, call):
        for i in range(count):
            call(1, 0, i)

    l = np.full(n, -np.inf); u = np.full(n, np.inf)
    x0 = truth * np.array([1.1, 0.9, 0.9, 1.1, 1.05, 0.95])
    out = []
    for tm in (None, tm_serial, tm_parallel):
        s = eng.settings(); s.maxIterations = 12
        x = x0.copy()
        r = eng.optimize_least_squares(s, m, x, l, u, f, None, tm)
        out.append((x.copy(), (r.status, r.iterations, r.fCalls, r.gCalls, r.residual, r.lambda_)))
    pool.shutdown()
    assert len(seen) > 1, "the parallel manager did not use several threads"
    for x, meta in out[1:]:
        assert np.array_equal(x, out[0][0]) and meta == out[0][1]
    # and the oracle agrees (its serial loop is the reference's semantics)
    s = oracle.settings(); s.maxIterations = 12
    xo = x0.copy(); ro = oracle.optimize_least_squares(s, m, xo, l, u, f)
    assert (ro.status, ro.iterations, ro.fCalls) == out[0][1][:3]
    np.testing.assert_allclose(out[0][0], xo, rtol=1e-8)


@pytest.mark.parametrize("config", ["c2", "c3", "general"])
def test_warm_start_uses_the_given_lambda_and_continues_a_fit(eng, config):
    if config == "c2":
        wl = workloads.c2_gauss4(20000, noise=0.05); fd = False            # thread-per-problem kernel
    elif config == "c3":
        wl = workloads.c3_sumexp8(2000); fd = True                         # four-problems-per-warp kernel
    else:
        wl = workloads.c2_gauss4(512, noise=0.05, m=200); fd = False       # m = 200: the general kernel (one CTA per problem)
    kw = dict(t=wl.t, y=wl.y, fd_jacobian=fd)
    # (1) cold run of k steps, then a warm-started continuation: it starts from the lambda the first call ended with
    s3 = eng.settings(np.float64); s3.maxIterations = 3
    x = wl.x0.copy(); r1, _ = eng.optimize_batched(s3, wl.model, x, wl.l, wl.u, **kw)
    lam1 = r1["lambda"].copy()
    s1 = eng.settings(np.float64); s1.maxIterations = 1
    xw = x.copy(); rw = r1.copy(); eng.optimize_batched(s1, wl.model, xw, wl.l, wl.u, results=rw, warm_start=True, **kw)
    xc = x.copy(); rc, _ = eng.optimize_batched(s1, wl.model, xc, wl.l, wl.u, **kw)
    # after one accepted step with a good gain ratio the damping is lambdaDecrease * lambda_in (LS:1158-1161): the warm run's
    # lambda must be tied to lam1, the cold run's (lambda_in = 0.001 max diag J'J, LS:1067-1072) must not
    good = (rw["iterations"] == 1) & (lam1 > 0)
    ratio = rw["lambda"][good] / lam1[good]
    tied = np.isclose(ratio, s1.lambdaDecrease, rtol=1e-12) | np.isclose(ratio, 1.0, rtol=1e-12) | np.isclose(ratio, s1.lambdaIncrease, rtol=1e-12)
    assert good.mean() > 0.5 and tied.mean() > 0.9, (float(good.mean()), float(tied.mean()))
    goodc = rc["iterations"] == 1
    assert np.mean(np.isclose(rc["lambda"][goodc] / lam1[goodc], s1.lambdaDecrease, rtol=1e-12)) < 0.2
    # (2) a warm-started full continuation reaches the same fit as one cold full run, with fewer evaluations than a cold restart
    sf = eng.settings(np.float64)
    xa = wl.x0.copy(); ra, _ = eng.optimize_batched(sf, wl.model, xa, wl.l, wl.u, **kw)
    xb = x.copy(); rb = r1.copy(); eng.optimize_batched(sf, wl.model, xb, wl.l, wl.u, results=rb, warm_start=True, **kw)
    xcr = x.copy(); rcr, _ = eng.optimize_batched(sf, wl.model, xcr, wl.l, wl.u, **kw)
    ok = (ra["status"] >= 0) & (rb["status"] >= 0)
    assert ok.mean() > 0.99
    rel = np.abs(rb["residual"][ok] - ra["residual"][ok]) / np.maximum(ra["residual"][ok], 1e-300)
    assert np.quantile(rel, 0.99) < 1e-6
    assert rb["fCalls"].mean() <= rcr["fCalls"].mean() * 1.02
