"""Test-side driver for the CPU oracle's batched entry points (oracle/models_oracle.cpp).
TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs."""
import ctypes as C

import numpy as np

from mir_optim_b200._abi import (ModelDesc, MODEL_FD_JACOBIAN, MODEL_GRID_PER_PROBLEM, LeastSquaresSettingsD,
                                 LeastSquaresSettingsS)
from mir_optim_b200.engine import RESULT_DTYPES


def _vp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def oracle_batched(lib, settings, model, x0, l, u, t=None, y=None, m=None, fd_jacobian=False, nthreads=0, aux=None, param=0.0):
    """Runs the oracle LM (least_squares.d:877-1176 restated) on every problem.  Returns (x, results, threads)."""
    x = np.array(x0, copy=True, order="C")
    dt = x.dtype
    sfx = "d" if dt == np.float64 else "s"
    assert isinstance(settings, LeastSquaresSettingsD if sfx == "d" else LeastSquaresSettingsS)
    batch, n = x.shape
    l = np.ascontiguousarray(l, dtype=dt); u = np.ascontiguousarray(u, dtype=dt)
    bound_stride = 0 if l.ndim == 1 else n
    flags = MODEL_FD_JACOBIAN if fd_jacobian else 0
    if y is not None:
        y = np.ascontiguousarray(y, dtype=dt); m = y.shape[1] if m is None else m
    if t is not None:
        t = np.ascontiguousarray(t, dtype=dt)
        if t.ndim == 2:
            flags |= MODEL_GRID_PER_PROBLEM
    if aux is not None:
        aux = np.ascontiguousarray(aux, dtype=dt)
        if aux.ndim == 2:
            flags |= 8                         # MIR_MODEL_AUX_PER_PROBLEM
    desc = ModelDesc(int(model), flags, _vp(t), _vp(y), _vp(aux), float(param))
    results = np.empty(batch, dtype=RESULT_DTYPES[dt])
    fn = getattr(lib, f"oracle_batched_{sfx}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_size_t, C.c_void_p, C.c_int]
    threads = fn(C.addressof(settings), C.addressof(desc), batch, m, n, _vp(x), _vp(l), _vp(u), bound_stride,
                 _vp(results), nthreads)
    return x, results, threads


def oracle_box_qp_batched(lib, P, q, l, u, settings=None, nthreads=0):
    dt = P.dtype
    sfx = "d" if dt == np.float64 else "s"
    batch, n, _ = P.shape
    x = np.zeros((batch, n), dtype=dt)
    status = np.empty(batch, dtype=np.int32); iters = np.empty(batch, dtype=np.uint32)
    fn = getattr(lib, f"oracle_box_qp_batched_{sfx}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t] + [C.c_void_p] * 7 + [C.c_int]
    fn(C.addressof(settings) if settings is not None else None, batch, n, _vp(P), _vp(q), _vp(l), _vp(u), _vp(x),
       _vp(status), _vp(iters), nthreads)
    return x, status, iters


def rel_err(a, b):
    """max over problems of ||a-b||_inf / max(||b||_inf, tiny) per problem."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.ndim == 1:
        return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    return np.max(np.abs(a - b), axis=1) / np.maximum(np.max(np.abs(b), axis=1), 1e-300)


# ---- process-parallel driver -------------------------------------------------------------------
# OpenBLAS serialises concurrent callers on a global buffer lock (negative scaling with OpenMP
# threads: 2.5k fits/s on 1 thread, 0.45k on 8), so "all host cores" is done with forked worker
# processes, one single-threaded OpenBLAS each.
_MP = {}


def _mp_worker(span):
    lo, hi = span
    a = _MP
    x, res, _ = oracle_batched(a["lib"], a["settings"], a["model"], a["x0"][lo:hi], a["l"] if a["l"].ndim == 1 else a["l"][lo:hi],
                               a["u"] if a["u"].ndim == 1 else a["u"][lo:hi], t=a["t"] if (a["t"] is None or a["t"].ndim == 1) else a["t"][lo:hi],
                               y=None if a["y"] is None else a["y"][lo:hi], m=a["m"], fd_jacobian=a["fd"], nthreads=1)
    return lo, x, res


def oracle_batched_mp(lib, settings, model, x0, l, u, t=None, y=None, m=None, fd_jacobian=False, procs=0):
    """Same contract as oracle_batched, spread over `procs` forked processes (0 = all cores)."""
    import multiprocessing as mp
    import os
    procs = procs or (os.cpu_count() or 1)
    batch = len(x0)
    if procs <= 1 or batch < 4 * procs:
        return oracle_batched(lib, settings, model, x0, l, u, t=t, y=y, m=m, fd_jacobian=fd_jacobian, nthreads=1)
    _MP.update(lib=lib, settings=settings, model=model, x0=np.ascontiguousarray(x0), l=np.asarray(l), u=np.asarray(u),
               t=None if t is None else np.asarray(t), y=None if y is None else np.asarray(y), m=m, fd=fd_jacobian)
    chunk = max(1, min(512, (batch + 4 * procs - 1) // (4 * procs)))
    spans = [(s, min(batch, s + chunk)) for s in range(0, batch, chunk)]
    x = np.empty_like(np.ascontiguousarray(x0)); res = np.empty(batch, dtype=RESULT_DTYPES[x.dtype])
    with mp.get_context("fork").Pool(procs) as pool:
        for lo, xs, rs in pool.imap_unordered(_mp_worker, spans):
            x[lo:lo + len(xs)] = xs; res[lo:lo + len(rs)] = rs
    _MP.clear()
    return x, res, procs


def _mp_qp_worker(span):
    lo, hi = span
    a = _MP
    x, st, it = oracle_box_qp_batched(a["lib"], a["P"][lo:hi], a["q"][lo:hi], a["l"][lo:hi], a["u"][lo:hi], a["settings"], nthreads=1)
    return lo, x, st, it


def oracle_box_qp_batched_mp(lib, P, q, l, u, settings=None, procs=0):
    """oracle_box_qp_batched over forked processes (0 = all cores), one single-threaded OpenBLAS each."""
    import multiprocessing as mp
    import os
    procs = procs or (os.cpu_count() or 1)
    batch = len(q)
    if procs <= 1 or batch < 4 * procs:
        return oracle_box_qp_batched(lib, P, q, l, u, settings, nthreads=1) + (1,)
    _MP.update(lib=lib, settings=settings, P=P, q=q, l=l, u=u)
    chunk = max(1, min(256, (batch + 4 * procs - 1) // (4 * procs)))
    spans = [(s, min(batch, s + chunk)) for s in range(0, batch, chunk)]
    x = np.empty_like(q); st = np.empty(batch, dtype=np.int32); it = np.empty(batch, dtype=np.uint32)
    with mp.get_context("fork").Pool(procs) as pool:
        for lo, xs, ss, its in pool.imap_unordered(_mp_qp_worker, spans):
            x[lo:lo + len(xs)] = xs; st[lo:lo + len(ss)] = ss; it[lo:lo + len(its)] = its
    _MP.clear()
    return x, st, it, procs


def self_sensitivity(lib, settings, wl, dtype=np.float64, fd=None, seed=99):
    """The reference algorithm's own conditioning: run the oracle on the workload and on a copy whose samples are
    each moved by ONE ulp (up or down at random) -- rounding-level input noise, which is what a different
    summation order or exp implementation amounts to.  Returns (x, results, dx_rel, dres_rel) of the unperturbed
    run and the per-problem relative deviations between the two runs.  (a 1-ulp change
    moves final parameters by ~sigma * 1e-7 and flips 10-30 % of the termination statuses.)"""
    fd = wl.fd_jacobian if fd is None else fd
    a = dict(t=wl.t.astype(dtype), fd_jacobian=fd)
    y = wl.y.astype(dtype)
    x1, r1, _ = oracle_batched_mp(lib, settings, wl.model, wl.x0.astype(dtype), wl.l.astype(dtype), wl.u.astype(dtype), y=y, **a)
    up = np.random.default_rng(seed).random(y.shape) < 0.5
    y2 = np.where(up, np.nextafter(y, dtype(np.inf)), np.nextafter(y, dtype(-np.inf))).astype(dtype)
    x2, r2, _ = oracle_batched_mp(lib, settings, wl.model, wl.x0.astype(dtype), wl.l.astype(dtype), wl.u.astype(dtype), y=y2, **a)
    return x1, r1, rel_err(x1, x2), rel_err(r1["residual"], r2["residual"])


def assert_within_conditioning(err, sens, factor=10.0, floor=1e-12, what="x"):
    """Quantile-wise: the GPU-vs-oracle deviation must not exceed `factor` x the oracle's own 1-ulp sensitivity."""
    for q in (0.5, 0.9, 0.99):
        assert np.quantile(err, q) <= factor * np.quantile(sens, q) + floor, (what, q, np.quantile(err, q), np.quantile(sens, q))
    assert err.max() <= 3 * factor * sens.max() + floor, (what, "max", err.max(), sens.max())
