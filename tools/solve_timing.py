#!/usr/bin/env python
"""Fixed cost vs per-iteration cost of srb_cg_minimize at a configuration (default cfg3)."""
import os, sys, time
from importlib import import_module
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srb200 as srb
wl = import_module("super-resolution_b200.workloads")
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cf = wl.CONFIGS[cfg]
H, W, C, s = cf["H"], cf["W"], cf["C"], cf["s"]
psf = wl.gaussian_psf(cf["K"], cf["sigma"])
shifts = wl.default_shifts(cf["N"], s)
eng = srb.Engine((cf["N"], C, H // s, W // s), s, psf, shifts)
work = wl.make(cfg, forward=lambda k, plane: eng.forward(k, plane))
eng.set_observations(work["lr"])
eng.set_regularizer(work["reg_kind"], work["lam"], work["btv_range"], work["btv_decay"])
x0 = np.ascontiguousarray(work["x0"], dtype=np.float64)
srb.pin_host(x0) if hasattr(srb, "pin_host") else None
for its in (1, 1, 5, 20, 20, 40):
    x = x0.copy()
    l0 = eng.timing()["kernel_launches"]
    t0 = time.perf_counter()
    _, rep = eng.cg_minimize(x, maxits=its)
    dt = time.perf_counter() - t0
    print("maxits %3d: %.2f ms total, iterations %d, evaluations %d, launches %d, cost %.6f" %
          (its, dt * 1e3, rep["iterations"], rep["num_evaluations"], eng.timing()["kernel_launches"] - l0, rep["final_cost"]))
eng.close()
import torch
eng = srb.Engine((cf["N"], C, H // s, W // s), s, psf, shifts)
eng.set_observations(work["lr"])
eng.set_regularizer(work["reg_kind"], work["lam"], work["btv_range"], work["btv_decay"])
stream = torch.cuda.ExternalStream(eng.stream_handle())
with torch.cuda.stream(stream):
    xd0 = torch.from_numpy(x0.reshape(-1)).cuda()
    for its in (1, 1, 5, 20, 20, 40):
        xd = xd0.clone()
        torch.cuda.synchronize()
        l0 = eng.timing()["kernel_launches"]
        t0 = time.perf_counter()
        rep = eng.cg_minimize_dev(xd, maxits=its)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("dev maxits %3d: %.2f ms total, iterations %d, evaluations %d, launches %d, cost %.6f" %
              (its, dt * 1e3, rep["iterations"], rep["num_evaluations"], eng.timing()["kernel_launches"] - l0, rep["final_cost"]))
eng.close()
