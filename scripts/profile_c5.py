"""Driver for timing / ncu: configs[4]a, batched box-constrained QPs (n = 64), device resident."""
import argparse, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mir_optim_b200 as mo

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=100000)
ap.add_argument("--n", type=int, default=64)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--dtype", default="f64")
a = ap.parse_args()
dt = torch.float64 if a.dtype == "f64" else torch.float32
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(5)
B, n = a.batch, a.n
P = torch.empty(B, n, n, dtype=dt, device=dev)
for s in range(0, B, 4096):
    e = min(B, s + 4096)
    A = torch.randn(e - s, 256, n, dtype=dt, device=dev, generator=g)
    P[s:e] = torch.bmm(A.transpose(1, 2), A) / 256 + 0.1 * torch.eye(n, dtype=dt, device=dev)
q = torch.randn(B, n, dtype=dt, device=dev, generator=g)
l = -2.0 * torch.rand(B, n, dtype=dt, device=dev, generator=g); u = 2.0 * torch.rand(B, n, dtype=dt, device=dev, generator=g)
x = torch.zeros(B, n, dtype=dt, device=dev)
status = torch.empty(B, dtype=torch.int32, device=dev); iters = torch.empty(B, dtype=torch.int32, device=dev)
for _ in range(a.launches):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    mo.engine.solve_box_qp_batched_device(P, q, l, u, x, status, iters)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    low = n * (n + 1) // 2 * P.element_size() + 4 * n * P.element_size() + 4
    print(f"boxqp B={B} n={n} {a.dtype}: {ms:.2f} ms -> {B / ms * 1e3:.0f} QP/s, {B * low / ms / 1e6:.1f} GB/s algorithmic "
          f"(lower triangle), {B * n * n * P.element_size() / ms / 1e6:.1f} GB/s full P")
print("status histogram", torch.bincount(status).tolist(), "mean BOXCQP iterations", iters.float().mean().item(),
      "active fraction", ((x == l) | (x == u)).float().mean().item())
