"""Driver for timing / ncu: configs[3], one large problem m x 128 (GAUSSMIX K=42), single GPU.
  --syrk-only : time the FP64-tensor J^T J kernel alone on a random J (CUDA events, L2 flushed by size: J >> L2)
  default     : full LM solve through mir_optimize_least_squares_sharded_d (comm = NULL)"""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mir_optim_b200 as mo
from mir_optim_b200 import workloads

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=4_000_000)
ap.add_argument("--K", type=int, default=42)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--syrk-only", action="store_true")
ap.add_argument("--max-iterations", type=int, default=0)
a = ap.parse_args()
eng = mo.engine
n = 3 * a.K + 2
if a.syrk_only:
    rows = (a.m + 31) // 32 * 32
    J = torch.randn(rows, n + (n & 1), dtype=torch.float64, device="cuda")
    packed = torch.empty(n * (n + 1) // 2, dtype=torch.float64, device="cuda")
    dmma = eng.lib.mir_b200_measure_peak_tflops(1, 5)
    for i in range(a.reps + 2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); eng.syrk_lower_device(J, n, packed); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        fl = rows * n * (n + 1)
        print(json.dumps({"syrk_ms": ms, "rows": rows, "n": n, "algorithmic_tflops": fl / ms / 1e9, "dmma_peak_tflops": dmma,
                          "frac_of_dmma_peak": fl / ms / 1e9 / dmma, "J_read_GBps": rows * J.shape[1] * 8 / ms / 1e6}))
    sys.exit(0)
wl = workloads.c4_gaussmix(m=a.m, K=a.K)
t = torch.from_numpy(wl.t).cuda(); y = torch.from_numpy(wl.y).cuda()
s = eng.settings()
if a.max_iterations:
    s.maxIterations = a.max_iterations
for i in range(a.reps):
    x = wl.x0[0].copy()
    l0 = eng.kernel_launches()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r, st = eng.optimize_sharded(s, wl.model, x, wl.l, wl.u, t, y, want_stats=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"solve_s": dt, "status": r.status, "iterations": r.iterations, "fCalls": r.fCalls, "gCalls": r.gCalls,
                      "residual": r.residual, "passes": st["passes"], "iterations_per_s": r.iterations / dt, "passes_per_s": st["passes"] / dt,
                      "launches": eng.kernel_launches() - l0, "max_rel_err_vs_truth": float(np.max(np.abs(x - wl.truth[0]) / np.abs(wl.truth[0]))),
                      "stats": st}))
