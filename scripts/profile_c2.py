"""Small driver for ncu: a few device-resident launches of the configs[1] kernel (not a benchmark)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mir_optim_b200 as mo
from mir_optim_b200 import workloads

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1 << 18)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--config", default="c2")
ap.add_argument("--dtype", default="f64")
a = ap.parse_args()
dt = np.float64 if a.dtype == "f64" else np.float32
wl = workloads.c2_gauss4(a.batch, noise=0.05, dtype=dt) if a.config == "c2" else workloads.c3_sumexp8(a.batch, dtype=dt)
dev = torch.device("cuda", 0)
eng = mo.engine
s = eng.settings(dt)
T = lambda v: torch.from_numpy(v).to(dev)
t, y, x0, l, u = T(wl.t), T(wl.y), T(wl.x0), T(wl.l), T(wl.u)
x = torch.empty_like(x0)
stats = torch.zeros(8, dtype=torch.int64, device=dev)
for _ in range(a.launches):
    x.copy_(x0)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    res = eng.optimize_batched_device(s, wl.model, x, l, u, t=t, y=y, fd_jacobian=wl.fd_jacobian, stats=stats)
    e1.record(); torch.cuda.synchronize()
    print(f"{wl.name}: {e0.elapsed_time(e1):.2f} ms -> {a.batch / e0.elapsed_time(e1) * 1e3:.0f} fits/s")
if a.launches > 1: print(dict(zip(("problems", "passes", "accepted", "fresh", "broyden", "evals", "qp_solves", "qp_it"), stats.cpu().tolist())))
