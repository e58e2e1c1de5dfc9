"""Deterministic synthetic workloads of BASELINE.json's configs.

One generator feeds the oracle, the CUDA library, the tests and bench.py, so CPU and GPU see
byte-identical inputs.  numpy's PCG64 streams are reproducible across machines.
"""
from __future__ import annotations

import numpy as np

from ._abi import ModelId


class Workload(dict):
    """dict with attribute access: model, m, n, t, y, x0, l, u, truth, fd_jacobian, name"""
    __getattr__ = dict.__getitem__


def c1_expdecay3(dtype=np.float64, seed=1, m=1000, noise=0.02):
    """configs[0]: y = a exp(-b x) + c, m=1000, n=3, FD Jacobian, unbounded."""
    rng = np.random.default_rng(seed)
    t = np.linspace(0.0, 10.0, m)
    truth = np.array([3.0, 0.7, 1.0])
    y = truth[0] * np.exp(-truth[1] * t) + truth[2] + noise * rng.standard_normal(m)
    x0 = np.array([[1.0, 1.0, 0.0]])
    return Workload(name="C1 exp-decay m=1000 n=3", model=ModelId.EXPDECAY3, m=m, n=3, t=t.astype(dtype),
                    y=y[None, :].astype(dtype), x0=x0.astype(dtype), l=np.full(3, -np.inf, dtype), u=np.full(3, np.inf, dtype),
                    truth=truth[None, :], fd_jacobian=True)


def c2_gauss4(batch, dtype=np.float64, seed=2, m=64, noise=0.05, rel_noise=None, tight_bounds=False):
    """configs[1]: batched 4-parameter Gaussian peak fits, m=64 samples, box bounds, analytic J.
    noise: absolute sigma; rel_noise (if given): sigma = rel_noise * A per problem (parity tier)."""
    rng = np.random.default_rng(seed)
    t = np.linspace(-4.0, 4.0, m)
    A = rng.uniform(1.0, 10.0, batch); mu = rng.uniform(-1.0, 1.0, batch)
    sg = rng.uniform(0.5, 1.5, batch); c = rng.uniform(0.0, 1.0, batch)
    truth = np.stack([A, mu, sg, c], axis=1)
    z = (t[None, :] - mu[:, None]) / sg[:, None]
    clean = A[:, None] * np.exp(-0.5 * z * z) + c[:, None]
    sigma = (rel_noise * A)[:, None] if rel_noise is not None else noise
    y = clean + sigma * rng.standard_normal((batch, m))
    l = np.array([0.0, -2.0, 0.3, 0.0]); u = np.array([20.0, 2.0, 1.4, 1.0])
    if tight_bounds:      # configs[4]b: at least one bound active at the solution
        l = np.array([0.0, -0.5, 0.3, 0.0]); u = np.array([20.0, 0.5, 1.0, 1.0])
    x0 = truth * rng.uniform(0.7, 1.3, truth.shape)
    x0[:, 1] = mu + rng.uniform(-0.3, 0.3, batch)
    x0 = np.clip(x0, l, u)
    return Workload(name=f"C2 gauss4 B={batch} m={m}", model=ModelId.GAUSS4, m=m, n=4, t=t.astype(dtype), y=y.astype(dtype),
                    x0=x0.astype(dtype), l=l.astype(dtype), u=u.astype(dtype), truth=truth, fd_jacobian=False)


def c3_sumexp8(batch, dtype=np.float64, seed=3, m=128, noise=0.01):
    """configs[2]: batched 8-parameter sum of 4 exponentials, FD Jacobian, unbounded."""
    rng = np.random.default_rng(seed)
    t = np.linspace(0.0, 5.0, m)
    a = rng.uniform(1.0, 5.0, (batch, 4))
    b = np.array([0.3, 1.0, 3.0, 9.0])[None, :] * rng.uniform(0.8, 1.2, (batch, 4))
    truth = np.empty((batch, 8)); truth[:, 0::2] = a; truth[:, 1::2] = b
    y = np.einsum("bk,bkm->bm", a, np.exp(-b[:, :, None] * t[None, None, :])) + noise * rng.standard_normal((batch, m))
    x0 = truth * rng.uniform(0.8, 1.2, truth.shape)
    return Workload(name=f"C3 sumexp8 B={batch} m={m}", model=ModelId.SUMEXP, m=m, n=8, t=t.astype(dtype), y=y.astype(dtype),
                    x0=x0.astype(dtype), l=np.full(8, -np.inf, dtype), u=np.full(8, np.inf, dtype), truth=truth,
                    fd_jacobian=True)


def c4_gaussmix(m=4_000_000, K=42, seed=4, noise=1e-3, dtype=np.float64, row_slice=None, width=(0.3, 0.5)):
    """configs[3]: one large problem, n = 3K+2 = 128: K Gaussians on a regular grid + linear baseline.
    row_slice=(lo, hi) generates only those rows (for row-sharded ranks) without materialising the rest.
    width: peak sigma range in units of the peak spacing; (0.3, 0.5) gives neighbouring peaks 2-3 sigma apart (overlapping,
    cond(J'J) ~ 1e4): with defaults the reference algorithm takes ~37 accepted steps and ends in the lambda-overflow tail."""
    n = 3 * K + 2
    rng = np.random.default_rng(seed)
    centers = (np.arange(K) + 0.5) / K
    amps = rng.uniform(0.5, 2.0, K)
    widths = rng.uniform(width[0], width[1], K) / K
    truth = np.empty(n); truth[0:3 * K:3] = amps; truth[1:3 * K:3] = centers; truth[2:3 * K:3] = widths
    truth[n - 2] = 0.3; truth[n - 1] = -0.2
    x0 = truth.copy()
    x0[0:3 * K:3] *= rng.uniform(0.9, 1.1, K)
    x0[1:3 * K:3] += rng.uniform(-0.1, 0.1, K) / K
    x0[2:3 * K:3] *= rng.uniform(0.9, 1.1, K)
    x0[n - 2:] += rng.uniform(-0.05, 0.05, 2)
    lo, hi = (0, m) if row_slice is None else row_slice
    t = (np.arange(lo, hi, dtype=np.float64) + 0.5) / m
    y = np.full(hi - lo, truth[n - 2]) + truth[n - 1] * t
    for k in range(K):
        z = (t - centers[k]) / widths[k]
        y += amps[k] * np.exp(-0.5 * z * z)
    # per-row noise from block-keyed counter streams, so every row shard sees the same global noise vector
    blk_len = 1 << 16
    noise_v = np.empty(hi - lo)
    pos = lo
    while pos < hi:
        blk = pos // blk_len
        chunk = np.random.Generator(np.random.Philox(key=[seed + 1000, blk])).standard_normal(blk_len)
        a0 = pos - blk * blk_len
        b0 = min(blk_len, a0 + (hi - pos))
        noise_v[pos - lo:pos - lo + (b0 - a0)] = chunk[a0:b0]
        pos += b0 - a0
    y += noise * noise_v
    return Workload(name=f"C4 gaussmix m={m} n={n}", model=ModelId.GAUSSMIX, m=m, n=n, t=t.astype(dtype), y=y.astype(dtype),
                    x0=x0[None, :].astype(dtype), l=np.full(n, -np.inf, dtype), u=np.full(n, np.inf, dtype),
                    truth=truth[None, :], fd_jacobian=False, rows=(lo, hi))


def c5_boxqp(batch, n=64, seed=5, dtype=np.float64, rows=256, bound_scale=2.0):
    """configs[4]a: batched box-constrained QPs, P = A'A/rows + 0.1 I, q ~ N(0,1), l = -U(0, s), u = +U(0, s).
    s = 2 puts 30-50 % of the variables on a bound at the solution (SURVEY 8d asks for that fraction; its
    s = 0.2 gives ~91 %)."""
    rng = np.random.default_rng(seed)
    P = np.empty((batch, n, n), dtype=dtype)
    chunk = 2048
    for s in range(0, batch, chunk):
        e = min(batch, s + chunk)
        A = rng.standard_normal((e - s, rows, n))
        P[s:e] = (np.einsum("brn,brk->bnk", A, A) / rows + 0.1 * np.eye(n)).astype(dtype)
    q = rng.standard_normal((batch, n)).astype(dtype)
    l = (-rng.uniform(0.0, bound_scale, (batch, n))).astype(dtype)
    u = (rng.uniform(0.0, bound_scale, (batch, n))).astype(dtype)
    return Workload(name=f"C5 boxqp B={batch} n={n}", P=P, q=q, l=l, u=u, n=n)
