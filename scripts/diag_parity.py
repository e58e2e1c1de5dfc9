"""Diagnostic (not a test): per-k trajectory agreement of the CUDA path vs the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mir_optim_b200 as mo
from mir_optim_b200 import workloads
from conftest import load_oracle
from oracle_util import oracle_batched_mp, rel_err
eng = mo.engine; olib = load_oracle()
for name, wl, fd, dt in (("c2", workloads.c2_gauss4(1024, noise=0.05), False, np.float64),
                         ("c2fd", workloads.c2_gauss4(1024, noise=0.05), True, np.float64),
                         ("c3", workloads.c3_sumexp8(512, noise=0.01), True, np.float64),
                         ("c2f32", workloads.c2_gauss4(1024, noise=0.05), False, np.float32)):
    for k in (1, 2, 3, 4, 6, 8, 12, 1000):
        sg = eng.settings(dt); sg.maxIterations = k
        xg = wl.x0.astype(dt).copy()
        rg, _ = eng.optimize_batched(sg, wl.model, xg, wl.l.astype(dt), wl.u.astype(dt), t=wl.t.astype(dt), y=wl.y.astype(dt), fd_jacobian=fd)
        xo, ro, _ = oracle_batched_mp(olib, sg, wl.model, wl.x0.astype(dt), wl.l.astype(dt), wl.u.astype(dt), t=wl.t.astype(dt), y=wl.y.astype(dt), fd_jacobian=fd)
        same = (rg["status"] == ro["status"]) & (rg["fCalls"] == ro["fCalls"]) & (rg["iterations"] == ro["iterations"]) & (rg["gCalls"] == ro["gCalls"])
        ex = rel_err(xg, xo); er = rel_err(rg["residual"], ro["residual"]); el = rel_err(rg["lambda"], ro["lambda"])
        bad = np.where(~same)[0]
        print(f"{name} k={k}: same {same.sum()}/{len(same)} status-eq {(rg['status']==ro['status']).mean():.3f} "
              f"max dx all {ex.max():.2e} same-only {ex[same].max() if same.any() else 0:.2e} dres {er.max():.2e} / {er[same].max() if same.any() else 0:.2e} dlam(same) {el[same].max() if same.any() else 0:.2e}")
        for b in bad[:3]:
            print("    prob", b, "gpu", tuple(rg[b]), "cpu", tuple(ro[b]))
