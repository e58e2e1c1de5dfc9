"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck).
usage: compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mir_optim_b200 as mo
from mir_optim_b200 import workloads
eng = mo.engine
s = eng.settings(); s.maxIterations = 6
# thread-per-problem kernel: v-list (analytic), stored-J (maxAge 5), finite differences; host-staged with watermark
wl = workloads.c2_gauss4(16384 + 100, noise=0.05)
for max_age, fd in ((0, False), (5, False), (0, True)):
    s.maxAge = max_age
    x = wl.x0.copy(); r, _ = eng.optimize_batched(s, wl.model, x, wl.l, wl.u, t=wl.t, y=wl.y, fd_jacobian=fd)
    print("tpp", max_age, fd, int(np.sum(r["status"] >= -1)))
s.maxAge = 0
# lane-group kernel: C3 (n = 8, FD) double and float, small C2
for dt in (np.float64, np.float32):
    w3 = workloads.c3_sumexp8(512, dtype=dt); s3 = eng.settings(dt); s3.maxIterations = 4
    x = w3.x0.copy(); r, _ = eng.optimize_batched(s3, w3.model, x, w3.l, w3.u, t=w3.t, y=w3.y, fd_jacobian=True)
    print("group c3", dt.__name__, int(np.sum(r["status"] >= -1)))
w2 = workloads.c2_gauss4(1024, noise=0.05); x = w2.x0.copy()
r, _ = eng.optimize_batched(s, w2.model, x, w2.l, w2.u, t=w2.t, y=w2.y); print("group c2", int(np.sum(r["status"] >= -1)))
# batched BoxQP
w5 = workloads.c5_boxqp(256, n=64)
xq, st, it = eng.solve_box_qp_batched(w5.P, w5.q, w5.l, w5.u); print("boxqp", int(np.sum(st == 0)))
# single large problem: fresh + fused Broyden SYRK + control kernels + graph replay
w4 = workloads.c4_gaussmix(m=16384 + 7, K=10, noise=1e-4)
t = torch.from_numpy(w4.t).cuda(); y = torch.from_numpy(w4.y).cuda()
s4 = eng.settings(); s4.maxIterations = 6
x = w4.x0[0].copy(); r, stt = eng.optimize_sharded(s4, w4.model, x, w4.l, w4.u, t, y, want_stats=True)
print("large", r.status, r.iterations, stt["broyden_updates"], stt["fresh_jacobians"])
