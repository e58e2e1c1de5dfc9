"""Host-side mirror of mir-optim's operator interface, over the C ABI.

Same names, argument meaning and error behaviour as the reference's D API
(least_squares.d:165-215 `optimize`, :459-519 `optimizeLeastSquares`, boxcqp.d:85-102
`solveBoxQP`), written against *a* bound shared object so that tests can drive the CUDA
library and the CPU oracle through identical code.  Nothing here computes: every call goes
through the C ABI declared in include/mir_optim_b200.h.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np

from . import _abi
from ._abi import (BoxQPStatus, LeastSquaresStatus, ModelDesc, ModelId,  # noqa: F401  (re-exported)
                   MODEL_FD_JACOBIAN, MODEL_GRID_PER_PROBLEM)


class LeastSquaresException(Exception):
    """Raised by :meth:`ReferenceAPI.optimize` for status < 0 (least_squares.d:48-70, 175-179)."""

    def __init__(self, status: LeastSquaresStatus, message: str):
        super().__init__("mir-optim Least Squares: " + message)
        self.status = status


def _types(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "d", C.c_double, _abi.LeastSquaresSettingsD, _abi.LeastSquaresResultD, _abi.SliceD, _abi.FunctionD, _abi.BoxQPSettingsD
    if dtype == np.float32:
        return "s", C.c_float, _abi.LeastSquaresSettingsS, _abi.LeastSquaresResultS, _abi.SliceS, _abi.FunctionS, _abi.BoxQPSettingsS
    raise TypeError("T must be float32 or float64 (least_squares.d:86)")


def _ptr(a, real):
    return a.ctypes.data_as(C.POINTER(real))


class ReferenceAPI:
    """The reference's extern(C) surface (header part 1) of one shared object."""

    def __init__(self, lib):
        self.lib = _abi.bind_reference_abi(lib)

    # -- settings ---------------------------------------------------------------------
    def settings(self, dtype=np.float64):
        """LeastSquaresSettings!T.init via mir_least_squares_init_{d,s} (least_squares.d:761-770)."""
        sfx, _, S, *_ = _types(dtype)
        s = S()
        getattr(self.lib, f"mir_least_squares_init_{sfx}")(C.byref(s))
        return s

    def status_string(self, status: int) -> str:
        return self.lib.mir_least_squares_status_string(int(status)).decode()

    # -- optimizeLeastSquares (nothrow, least_squares.d:459-519) -----------------------
    def optimize_least_squares(self, settings, m: int, x: np.ndarray, l: np.ndarray, u: np.ndarray,
                               f: Callable, g: Optional[Callable] = None, tm: Optional[Callable] = None,
                               zero_outputs: bool = True):
        """f(x, y) fills y (length m); g(x, J) fills row-major J (m x n).  Like the reference's
        template wrapper, y / J are zeroed before each user call (least_squares.d:469, 482)."""
        sfx, real, S, R, Sl, FT, _ = _types(x.dtype)
        assert isinstance(settings, S), "settings precision must match x"
        n = x.shape[0]
        assert x.flags.c_contiguous and l.dtype == x.dtype and u.dtype == x.dtype
        dt = x.dtype

        def f_c(_ctx, m_, n_, xp, yp):
            xv = np.ctypeslib.as_array(xp, shape=(n_,))
            yv = np.ctypeslib.as_array(yp, shape=(m_,))
            if zero_outputs:
                yv[:] = 0
            f(xv, yv)

        def g_c(_ctx, m_, n_, xp, jp):
            xv = np.ctypeslib.as_array(xp, shape=(n_,))
            jv = np.ctypeslib.as_array(jp, shape=(m_, n_))
            if zero_outputs:
                jv[:] = 0
            g(xv, jv)

        f_cb = FT(f_c)
        g_cb = FT(g_c) if g is not None else None
        tm_cb = None
        if tm is not None:
            def tm_c(_ctx, count, task, task_fn):
                tm(count, lambda total, tid, i: task_fn(task, total, tid, i))
            tm_cb = _abi.ThreadManager(tm_c)

        wl = self.lib.mir_least_squares_work_length(m, n)
        iwl = self.lib.mir_least_squares_iwork_length(m, n)
        work = np.empty(max(wl, 1), dtype=dt)
        iwork = np.empty(max(iwl, 1), dtype=np.int32)
        fn = getattr(self.lib, f"mir_optimize_least_squares_{sfx}")
        res = fn(C.byref(settings), m, n, _ptr(x, real), _ptr(l, real), _ptr(u, real),
                 Sl(wl, _ptr(work, real)), _abi.SliceI(iwl, iwork.ctypes.data_as(C.POINTER(C.c_int32))),
                 None, C.cast(f_cb, C.c_void_p),
                 None, C.cast(g_cb, C.c_void_p) if g_cb is not None else None,
                 None, C.cast(tm_cb, C.c_void_p) if tm_cb is not None else None)
        return res

    # -- optimize (throws, least_squares.d:165-181) ------------------------------------
    def optimize(self, settings, m, x, l, u, f, g=None, tm=None):
        res = self.optimize_least_squares(settings, m, x, l, u, f, g, tm)
        if res.status < 0:
            st = LeastSquaresStatus(res.status)
            raise LeastSquaresException(st, self.status_string(res.status))
        return res


__all__ = ["ReferenceAPI", "LeastSquaresException", "LeastSquaresStatus", "BoxQPStatus", "ModelId", "ModelDesc",
           "MODEL_FD_JACOBIAN", "MODEL_GRID_PER_PROBLEM"]
