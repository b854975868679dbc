#!/usr/bin/env python
"""A/B of the default fused tile kernel against its Z-layout variant (SRB_ZLAYOUT=1) at cfg3
(2048 x 2048 RGB, 16 frames, 4x, 7x7 PSF, TV), device resident, torch-free: kernel time from the
library's own CUDA events (srb_set_profiling), plus the difference of the two gradients.
    python tools/ab_zlayout.py [iterations]"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
srb = importlib.import_module("super-resolution_b200")
wl = importlib.import_module("super-resolution_b200.workloads")

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
C, H, W, N, s, K, sigma = 3, 2048, 2048, 16, 4, 7, 1.5
rng = np.random.default_rng(1)
lr = rng.random((N, C, H // s, W // s))
x = rng.random(C * H * W)
n = x.size
out = {}
grads = {}
for name, flag in (("default", "0"), ("zlayout", "1"), ("default_again", "0")):
    os.environ["SRB_ZLAYOUT"] = flag
    with srb.Engine(lr.shape, s, wl.gaussian_psf(K, sigma), wl.default_shifts(N, s)) as e:
        e.set_observations(lr)
        e.set_regularizer(srb.REG_TV, 0.01)
        cost, _ = e.eval(x, want_grad=False)      # uploads x into the context's own buffer
        xd = e.dev_x_ptr()
        gd = srb.dev_alloc((n + 1) * 8)
        e.set_profiling(True)
        ms = []
        for _ in range(iters):
            e.eval_partial_dev(xd, gd)
            e.synchronize()
            ms.append(e.timing()["last_main_kernel_ms"])
        g = np.empty(n + 1)
        e.memcpy_d2h(g, gd, g.nbytes)
        srb.dev_free(gd)
        ms = np.array(ms[5:])
        out[name] = {"zlayout_active": e.zlayout_active, "kernel_ms_min": float(ms.min()),
                     "kernel_ms_median": float(np.median(ms)), "cost": float(g[n]), "cost_host_path": cost}
        grads[name] = g
d = grads["zlayout"] - grads["default"]
out["rel_l2_gradient_z_vs_default"] = float(np.linalg.norm(d[:n]) / np.linalg.norm(grads["default"][:n]))
out["rel_cost_z_vs_default"] = float(abs(d[n]) / abs(grads["default"][n]))
alg = wl.algorithmic_bytes(H, W, C, N, s, has_reg=True)
for k in ("default", "zlayout"):
    out[k]["algorithmic_GBps"] = alg / out[k]["kernel_ms_median"] / 1e6
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_zlayout.json"), "w"), indent=1)
