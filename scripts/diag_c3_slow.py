"""Why is configs[2] slower inside bench.py than alone?  Times C3 (double) alone, after device-resident C2, after host-pointer C2."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mir_optim_b200 as mo
from mir_optim_b200 import workloads
eng = mo.engine; dev = torch.device("cuda", 0)
T = lambda v: torch.from_numpy(v).to(dev)
def c3(tag, B=65536):
    wl = workloads.c3_sumexp8(B); s = eng.settings(np.float64)
    t, y, x0, l, u = T(wl.t), T(wl.y), T(wl.x0), T(wl.l), T(wl.u); x = torch.empty_like(x0)
    for i in range(3):
        x.copy_(x0); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); eng.optimize_batched_device(s, wl.model, x, l, u, t=t, y=y, fd_jacobian=True); e1.record(); torch.cuda.synchronize()
        print(tag, f"C3 {e0.elapsed_time(e1):.1f} ms", flush=True)
mode = sys.argv[1]
if mode in ("dev", "host"):
    wl = workloads.c2_gauss4(1 << 20); s = eng.settings(np.float64)
    if mode == "dev":
        t, y, x0, l, u = T(wl.t), T(wl.y), T(wl.x0), T(wl.l), T(wl.u); x = x0.clone()
        eng.optimize_batched_device(s, wl.model, x, l, u, t=t, y=y); torch.cuda.synchronize()
    else:
        x = wl.x0.copy(); eng.optimize_batched(s, wl.model, x, wl.l, wl.u, t=wl.t, y=wl.y)
c3(mode)
