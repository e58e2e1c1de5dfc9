"""GPU parity of the CTA-per-QP BOXCQP kernel (mir_solve_box_qp*, boxcqp.d:85-102 / 122-379) against the CPU
oracle (restated solveBoxQP over real LAPACK posvx).  Through the C ABI."""
import numpy as np
import pytest

from mir_optim_b200 import workloads
from mir_optim_b200._abi import BoxQPStatus, BoxQPSettingsD
from oracle_util import oracle_box_qp_batched, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import mir_optim_b200
    assert mir_optim_b200.engine.device_count() > 0, "no CUDA device"
    return mir_optim_b200.engine


def test_bq1_reference_unit_test(eng):
    """boxcqp.d:381-402"""
    P = np.array([[2.0, -1, 0], [-1.0, 2, -1], [0.0, -1, 2]])
    q = np.array([3.0, -7, 5]); l = np.array([-100.0, -2, 1]); u = np.array([100.0, 2, 1]); x = np.zeros(3)
    assert eng.solve_box_qp(P, q, l, u, x) == BoxQPStatus.solved
    np.testing.assert_allclose(x, [-0.5, 2, 1], rtol=1e-12, atol=1e-12)


def kkt_violation(P, q, l, u, x):
    """max KKT residual of a box QP solution (size independent property)."""
    Ps = np.tril(P) + np.tril(P, -1).transpose(0, 2, 1)
    g = np.einsum("bij,bj->bi", Ps, x) + q
    free = (x > l) & (x < u)
    v = np.where(free, np.abs(g), 0.0)
    v = np.maximum(v, np.where(x == l, np.maximum(-g, 0), 0))       # at lower bound the gradient must be >= 0
    v = np.maximum(v, np.where(x == u, np.maximum(g, 0), 0))
    return v.max(axis=1)


@pytest.mark.parametrize("n", [1, 2, 3, 8, 17, 64, 100, 128])
def test_random_qps_double(eng, oracle_lib, n):
    wl = workloads.c5_boxqp(96 if n > 64 else 256, n=n, seed=50 + n)
    xg, sg, ig = eng.solve_box_qp_batched(wl.P, wl.q, wl.l, wl.u)
    xo, so, io = oracle_box_qp_batched(oracle_lib, wl.P, wl.q, wl.l, wl.u)
    assert np.array_equal(sg, so) and np.all(sg == BoxQPStatus.solved)
    assert np.array_equal((xg == wl.l), (xo == wl.l)) and np.array_equal((xg == wl.u), (xo == wl.u))   # same active set
    assert np.array_equal(ig, io)                                                                        # same BOXCQP iteration count
    assert np.max(rel_err(xg, xo)) < 1e-10 if n > 1 else True
    assert np.max(np.abs(xg - xo)) < 1e-12
    assert np.all(xg >= wl.l) and np.all(xg <= wl.u)
    assert kkt_violation(wl.P, wl.q, wl.l, wl.u, xg).max() < 1e-12


def test_c5a_config(eng, oracle_lib):
    """BASELINE configs[4]a shape (n = 64), parity-sized batch."""
    wl = workloads.c5_boxqp(4096, n=64)
    xg, sg, ig = eng.solve_box_qp_batched(wl.P, wl.q, wl.l, wl.u)
    xo, so, io = oracle_box_qp_batched(oracle_lib, wl.P, wl.q, wl.l, wl.u)
    active = ((xo == wl.l) | (xo == wl.u)).mean()
    assert 0.25 < active < 0.6
    assert np.array_equal(sg, so) and np.array_equal(ig, io)
    assert np.max(np.abs(xg - xo)) < 1e-12


def test_random_qps_float(eng, oracle_lib):
    wl = workloads.c5_boxqp(256, n=32, dtype=np.float32)
    xg, sg, ig = eng.solve_box_qp_batched(wl.P, wl.q, wl.l, wl.u)
    xo, so, io = oracle_box_qp_batched(oracle_lib, wl.P, wl.q, wl.l, wl.u)
    assert np.all(sg == 0) and np.all(so == 0)
    ok = np.all((xg == wl.l) == (xo == wl.l), axis=1) & np.all((xg == wl.u) == (xo == wl.u), axis=1)
    assert ok.mean() > 0.95            # float: a multiplier within rounding of zero may flip one variable
    assert np.max(np.abs(xg[ok] - xo[ok])) < 1e-4
    assert kkt_violation(wl.P.astype(np.float64), wl.q.astype(np.float64), wl.l.astype(np.float64), wl.u.astype(np.float64),
                         xg.astype(np.float64)).max() < 1e-4


def test_edge_cases(eng, oracle_lib):
    n = 6
    rng = np.random.default_rng(3)
    A = rng.standard_normal((5, 20, n)); P = np.einsum("brn,brk->bnk", A, A) / 20 + 0.1 * np.eye(n)
    q = rng.standard_normal((5, n))
    l = np.full((5, n), -np.inf); u = np.full((5, n), np.inf)
    l[1, 2] = u[1, 2] = 0.25                    # fixed variable (l == u)
    l[2] = -0.01; u[2] = 0.01                   # (nearly) everything active
    P[3] = -P[3]                                # not positive definite -> numericError (boxcqp.d:212-213)
    P[4, 1, 1] = np.nan                         # NaN pivot: OpenBLAS' potrf lets it through, every comparison fails,
                                                # all variables end up 'free' -> maxIterations (boxcqp.d:265-266, 378)
    xg, sg, ig = eng.solve_box_qp_batched(P, q, l, u)
    xo, so, io = oracle_box_qp_batched(oracle_lib, P, q, l, u)
    assert list(sg) == list(so) == [0, 0, 0, 1, 2]
    assert np.max(np.abs(xg[:3] - xo[:3])) < 1e-12
    assert xg[1, 2] == 0.25
    np.testing.assert_allclose(xg[0], np.linalg.solve(P[0], -q[0]), rtol=1e-12)     # unconstrained: x = -P^-1 q
    # empty batch, and custom settings (1 BOXCQP iteration only -> maxIterations status, boxcqp.d:378)
    x0, s0, _ = eng.solve_box_qp_batched(np.zeros((0, n, n)), np.zeros((0, n)), np.zeros((0, n)), np.zeros((0, n)))
    assert x0.shape == (0, n)
    wl = workloads.c5_boxqp(64, n=16, seed=9)
    st = BoxQPSettingsD(16 * np.finfo(float).eps, 16 * np.finfo(float).eps, 1)
    xg, sg, ig = eng.solve_box_qp_batched(wl.P, wl.q, wl.l, wl.u, settings=st)
    xo, so, io = oracle_box_qp_batched(oracle_lib, wl.P, wl.q, wl.l, wl.u, settings=st)
    assert np.array_equal(sg, so) and (sg == BoxQPStatus.maxIterations).any()
