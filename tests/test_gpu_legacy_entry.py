"""The reference's own unit-test scenarios (least_squares.d:217-434) and API-contract checks, run through the CUDA
library's implementation of the reference's extern(C) entry point mir_optimize_least_squares_d with HOST callbacks
(f / g / thread manager in Python): the LM state machine and all dense algebra run on the GPU, the callbacks run on
the host exactly when the reference would call them.  Same test bodies as test_oracle_reference_scenarios.py."""
import numpy as np
import pytest

import test_oracle_reference_scenarios as ref

pytestmark = pytest.mark.gpu

SCENARIOS = ["test_t1_linear_with_jacobian", "test_t2_rosenbrock_fd_with_thread_manager", "test_t2_rosenbrock_fd_default_tm",
             "test_t3_rosenbrock_analytic_and_box", "test_t4_exp_decay", "test_t5_one_sided_bounds", "test_t6_degenerate_m_lt_n",
             "test_validation_statuses", "test_max_iterations_is_status_minus_one", "test_work_lengths", "test_defaults_and_layout"]


@pytest.fixture(scope="module")
def cuda_api():
    import mir_optim_b200 as mo
    assert mo.engine.device_count() > 0
    return mo.engine


@pytest.mark.parametrize("name", SCENARIOS)
def test_reference_scenario_through_cuda_library(name, cuda_api):
    before = cuda_api.kernel_launches()
    getattr(ref, name)(cuda_api)
    if name.startswith("test_t"):
        assert cuda_api.kernel_launches() > before, "the CUDA path did not run"


def test_counts_match_oracle_on_reference_scenarios(cuda_api, oracle):
    """status / iterations / fCalls / gCalls / x of T1, T3a, T3b agree with the oracle (host-callback mode)."""
    INF = np.inf
    def lin_f(x, y): y[0] = x[0]; y[1] = 2 - x[1]
    def lin_g(x, J): J[0, 0] = 1; J[0, 1] = 0; J[1, 0] = 0; J[1, 1] = -1
    cases = [(lin_f, lin_g, [100.0, 100.0], [-INF, -INF], [INF, INF]),
             (ref.rosen_f, ref.rosen_g, [-1.2, 1.0], [-INF, -INF], [INF, INF]),
             (ref.rosen_f, ref.rosen_g, [150.0, 150.0], [10.0, 10.0], [200.0, 200.0]),
             (ref.rosen_f, None, [-1.2, 1.0], [-INF, -INF], [INF, INF])]
    for f, g, x0, l, u in cases:
        xa = np.array(x0); xb = np.array(x0)
        ra = cuda_api.optimize_least_squares(cuda_api.settings(), 2, xa, np.array(l), np.array(u), f, g)
        rb = oracle.optimize_least_squares(oracle.settings(), 2, xb, np.array(l), np.array(u), f, g)
        assert (ra.status, ra.iterations, ra.fCalls, ra.gCalls) == (rb.status, rb.iterations, rb.fCalls, rb.gCalls), (x0, ra, rb)
        np.testing.assert_allclose(xa, xb, rtol=1e-10, atol=1e-12)
        assert ra.residual == pytest.approx(rb.residual, rel=1e-9, abs=1e-25)


def test_float_entry_uses_real_m(cuda_api):
    """mir_optimize_least_squares_s runs with the real m (the reference passes 2, least_squares.d:629 -- not reproduced)."""
    t = np.linspace(0, 4, 50, dtype=np.float32)
    ydata = (2.0 * np.exp(-0.7 * t) + 0.5).astype(np.float32)
    def f(p, y): y[:] = p[0] * np.exp(-p[1] * t) + p[2] - ydata
    x = np.array([1.0, 1.0, 0.0], dtype=np.float32)
    r = cuda_api.optimize_least_squares(cuda_api.settings(np.float32), 50, x, np.full(3, -np.inf, np.float32), np.full(3, np.inf, np.float32), f)
    assert r.status >= 0
    np.testing.assert_allclose(x, [2.0, 0.7, 0.5], rtol=2e-3)
