"""Generates tests/golden/oracle_c2_k3.json: the CPU oracle's results on 24 seeded configs[1] problems stopped after
k = 3 accepted steps (trajectories have not forked yet) and run to termination on noise-free data with
maxGoodResidual = 1e-20 (P3).  The reference itself cannot be run here (no D compiler): these are the ORACLE's outputs,
pinned so that neither the oracle nor the CUDA path can drift unnoticed.   python scripts/make_golden.py"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_oracle
from oracle_util import oracle_batched
from mir_optim_b200 import workloads
from mir_optim_b200.api import ReferenceAPI

lib = load_oracle(); api = ReferenceAPI(lib)
out = {}
wl = workloads.c2_gauss4(24, noise=0.05, seed=2024)
s = api.settings(); s.maxIterations = 3
x, r, _ = oracle_batched(lib, s, wl.model, wl.x0, wl.l, wl.u, t=wl.t, y=wl.y)
out["k3"] = {"seed": 2024, "noise": 0.05, "maxIterations": 3, "x": x.tolist(), "status": r["status"].tolist(), "iterations": r["iterations"].tolist(),
             "fCalls": r["fCalls"].tolist(), "gCalls": r["gCalls"].tolist(), "residual": r["residual"].tolist(), "lambda": r["lambda"].tolist()}
wl = workloads.c2_gauss4(24, noise=0.0, seed=2025)
l = np.array([0.0, -2.0, 0.3, -1.0]); u = np.array([20.0, 2.0, 2.0, 2.0])
s = api.settings(); s.maxGoodResidual = 1e-20
x, r, _ = oracle_batched(lib, s, wl.model, wl.x0, l, u, t=wl.t, y=wl.y)
out["robust"] = {"seed": 2025, "noise": 0.0, "maxGoodResidual": 1e-20, "l": l.tolist(), "u": u.tolist(), "x": x.tolist(), "status": r["status"].tolist(),
                 "iterations": r["iterations"].tolist(), "fCalls": r["fCalls"].tolist(), "gCalls": r["gCalls"].tolist(), "residual": r["residual"].tolist()}
with open(os.path.join(ROOT, "tests", "golden", "oracle_c2_k3.json"), "w") as f:
    json.dump(out, f)
print("written", {k: len(v["x"]) for k, v in out.items()})
