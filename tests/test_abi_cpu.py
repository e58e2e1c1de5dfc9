"""CPU-side checks of the drop-in boundary (no compute calls): the C-ABI library loads, exports every function
include/mir_optim_b200.h declares, its host-only entry points behave like the reference's (work lengths LS:642-656,
BQ:36-50; status strings LS:528-557; init/reset LS:761-792), and without a CUDA device every compute entry fails loudly
(no CPU fallback).  Plus the golden fixture: the oracle must still reproduce tests/golden/oracle_c2_k3.json."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mir_optim_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    names = set(re.findall(r"\b(mir_[a-z0-9_]+)\s*\(", text))
    return sorted(n for n in names if not n.endswith("_fn"))


def test_library_exports_every_declared_function():
    import mir_optim_b200
    lib = mir_optim_b200.engine.lib
    names = declared_functions()
    assert len(names) >= 30, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_host_only_entry_points_match_reference_formulas():
    import mir_optim_b200
    lib = mir_optim_b200.engine.lib
    for fn in ("mir_least_squares_work_length", "mir_least_squares_iwork_length"):
        getattr(lib, fn).argtypes = [C.c_size_t, C.c_size_t]; getattr(lib, fn).restype = C.c_size_t
    for fn in ("mir_box_qp_work_length", "mir_box_qp_iwork_length"):
        getattr(lib, fn).argtypes = [C.c_size_t]; getattr(lib, fn).restype = C.c_size_t
    for m, n in ((1000, 3), (64, 4), (128, 8), (1, 2)):
        qp = 2 * n * n + 8 * n                                                    # BQ:36-42
        assert lib.mir_box_qp_work_length(n) == qp
        assert lib.mir_box_qp_iwork_length(n) == n + (n + 3) // 4                 # BQ:45-50 (flags as bytes)
        assert lib.mir_least_squares_work_length(m, n) == qp + 5 * n + n * n + n * m + 2 * m      # LS:642-646
        assert lib.mir_least_squares_iwork_length(m, n) == max(n + (n + 3) // 4, n)               # LS:651-656
    lib.mir_least_squares_status_string.restype = C.c_char_p
    seen = {lib.mir_least_squares_status_string(st) for st in (3, 2, 1, 0, -1, -26, -27, -28, -29, -30, -31, -32)}
    assert len(seen) == 12 and all(s for s in seen)                               # twelve distinct non-empty strings
    eng = mir_optim_b200.engine
    for dt in (np.float64, np.float32):
        s = eng.settings(dt)
        assert s.maxIterations == 1000 and s.maxAge == 0                          # LS:85-123 defaults
        assert 0 < s.minStepQuality < s.goodStepQuality <= 1 and s.lambdaIncrease >= 1 and 0 < s.lambdaDecrease <= 1


def test_no_device_means_loud_failure_not_cpu_fallback():
    import mir_optim_b200
    from mir_optim_b200 import workloads
    eng = mir_optim_b200.engine
    if eng.device_count() > 0:
        pytest.skip("a CUDA device is present")
    wl = workloads.c2_gauss4(4)
    from mir_optim_b200.engine import B200Error
    with pytest.raises(B200Error) as ei:
        eng.optimize_batched(eng.settings(), wl.model, wl.x0.copy(), wl.l, wl.u, t=wl.t, y=wl.y)
    assert "no usable CUDA device" in str(ei.value) or "CPU fallback" in str(ei.value)
    P = np.eye(3); q = np.ones(3); x = np.zeros(3)
    with pytest.raises(B200Error):
        eng.solve_box_qp(P, q, -np.ones(3), np.ones(3), x)
    assert np.all(x == 0)                                                         # nothing was computed on the host


def test_oracle_reproduces_golden_fixture(oracle_lib, oracle):
    from mir_optim_b200 import workloads
    from oracle_util import oracle_batched
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_c2_k3.json")))
    k3 = g["k3"]
    wl = workloads.c2_gauss4(24, noise=k3["noise"], seed=k3["seed"])
    s = oracle.settings(); s.maxIterations = k3["maxIterations"]
    x, r, _ = oracle_batched(oracle_lib, s, wl.model, wl.x0, wl.l, wl.u, t=wl.t, y=wl.y)
    for f in ("status", "iterations", "fCalls", "gCalls"):
        assert r[f].tolist() == k3[f], f
    np.testing.assert_allclose(x, np.array(k3["x"]), rtol=1e-13)
    np.testing.assert_allclose(r["residual"], np.array(k3["residual"]), rtol=1e-13)
    rb = g["robust"]
    wl = workloads.c2_gauss4(24, noise=rb["noise"], seed=rb["seed"])
    s = oracle.settings(); s.maxGoodResidual = rb["maxGoodResidual"]
    x, r, _ = oracle_batched(oracle_lib, s, wl.model, wl.x0, np.array(rb["l"]), np.array(rb["u"]), t=wl.t, y=wl.y)
    assert r["status"].tolist() == rb["status"] and r["iterations"].tolist() == rb["iterations"]
    np.testing.assert_allclose(x, np.array(rb["x"]), rtol=1e-12)
