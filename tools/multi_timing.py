#!/usr/bin/env python
"""srb_multi_eval (single process, G devices, host x in / host g out) against srb_eval at a configuration."""
import os, sys, time
from importlib import import_module
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srb200 as srb
wl = import_module("super-resolution_b200.workloads")
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cf = wl.CONFIGS[cfg]
H, W, C, s, N = cf["H"], cf["W"], cf["C"], cf["s"], cf["N"]
psf = wl.gaussian_psf(cf["K"], cf["sigma"])
shifts = wl.default_shifts(N, s)
rng = np.random.default_rng(1)
lr = rng.random((N, C, H // s, W // s))
x = rng.random(C * H * W)
g = np.empty_like(x)
srb.pin_host(x); srb.pin_host(g)
ref = None
for G in [1, 2, 4, 8]:
    if G > srb.device_count():
        break
    with srb.MultiEngine(lr.shape, s, psf, shifts, n_gpus=G) as me:
        me.set_observations(lr)
        me.set_regularizer(srb.REG_TV, 0.01)
        for _ in range(3):
            c, _ = me.eval(x, out=g)
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            c, _ = me.eval(x, out=g)
        dt = (time.perf_counter() - t0) / reps
        if ref is None:
            ref = (c, g.copy())
        rel = np.linalg.norm(g - ref[1]) / np.linalg.norm(ref[1])
        print("G=%d: %.3f ms per srb_multi_eval (wall), cost %.10g, rel L2 vs G=1 %.2e, dev0 span %.3f ms" %
              (G, dt * 1e3, c, rel, me.timing()["last_eval_kernel_ms"]))
