"""Small single-problem solves through the large-problem path for compute-sanitizer (memcheck / racecheck / synccheck):
the one-CTA control kernels at n = 128 (136 register tiles, fifth warp), n = 100 (13 tile rows), n = 32, double and float.
usage: compute-sanitizer --tool racecheck python scripts/sanitize_large_ctl.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mir_optim_b200 as mo
from mir_optim_b200 import workloads
eng = mo.engine
for K, dt in ((42, np.float64), (33, np.float64), (10, np.float64), (42, np.float32)):
    w4 = workloads.c4_gaussmix(m=4096 + 7, K=K, noise=1e-4)
    t = torch.from_numpy(w4.t.astype(dt)).cuda(); y = torch.from_numpy(w4.y.astype(dt)).cuda()
    s4 = eng.settings(dt); s4.maxIterations = 3
    x = w4.x0[0].astype(dt).copy()
    if dt == np.float64:
        r, stt = eng.optimize_sharded(s4, w4.model, x, w4.l, w4.u, t, y, want_stats=True)
        print("large n =", 3 * K + 2, dt.__name__, "status", r.status, "iterations", r.iterations, "passes", int(stt["passes"]), "qp solves", int(stt["qp_solves"]), flush=True)
    else:       # float goes through the reference's own entry point (device-model mode, m > 128 => the large-problem engine)
        r = eng.optimize_device_model(s4, w4.model, x, w4.l.astype(dt), w4.u.astype(dt), t=w4.t.astype(dt), y=w4.y.astype(dt))
        print("large n =", 3 * K + 2, dt.__name__, "status", r.status, "iterations", r.iterations, flush=True)
