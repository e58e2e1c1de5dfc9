"""The reference has no float test and its float entry point is defective (it passes m = 2, least_squares.d:629), so the
float oracle (the restated algorithm with the real m) is "parity unpinned" against the reference.  What CAN be pinned on
the CPU: on well-conditioned problems the float oracle must agree with the double oracle to float accuracy, with the same
termination class -- a float oracle that drifted (wrong epsilon-dependent constant, wrong LAPACK routine) fails this."""
import numpy as np

from mir_optim_b200 import workloads
from mir_optim_b200._abi import LeastSquaresStatus as S
from oracle_util import oracle_batched, rel_err


def test_float_oracle_tracks_double_oracle(oracle_lib, oracle):
    wl = workloads.c2_gauss4(256, noise=0.0, seed=77)
    l = np.array([0.0, -2.0, 0.3, -1.0]); u = np.array([20.0, 2.0, 2.0, 2.0])
    sd = oracle.settings(np.float64); sd.maxGoodResidual = 1e-9
    ss = oracle.settings(np.float32); ss.maxGoodResidual = 1e-9
    xd, rd, _ = oracle_batched(oracle_lib, sd, wl.model, wl.x0, l, u, t=wl.t, y=wl.y)
    f = np.float32
    xs, rs, _ = oracle_batched(oracle_lib, ss, wl.model, wl.x0.astype(f), l.astype(f), u.astype(f), t=wl.t.astype(f), y=wl.y.astype(f))
    assert np.all(rd["status"] == S.fConverged)
    assert np.all(rs["status"] >= 0)
    assert np.quantile(rel_err(xs.astype(np.float64), xd), 0.99) < 2e-3 and np.max(rel_err(xs.astype(np.float64), xd)) < 2e-2
    assert np.median(rs["residual"]) < 1e-6          # float round-off floor of sum r^2 over 64 samples of O(1..10) peaks


def test_float_settings_defaults_follow_the_float_epsilon(oracle):
    s = oracle.settings(np.float32); d = oracle.settings(np.float64)
    # least_squares.d:100-106: tolerances are multiples of T.epsilon, jacobianEpsilon = sqrt(T.epsilon)
    assert abs(s.jacobianEpsilon - 2.0 ** -11.5) < 1e-6 or abs(s.jacobianEpsilon - np.sqrt(np.finfo(np.float32).eps)) < 1e-6 or s.jacobianEpsilon == 2.0 ** -11
    assert d.jacobianEpsilon == 2.0 ** -26
    assert s.maxIterations == d.maxIterations == 1000
    assert s.lambdaIncrease == d.lambdaIncrease and abs(s.lambdaDecrease - d.lambdaDecrease) < 1e-6
