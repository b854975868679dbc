#!/usr/bin/env python
"""Device-resident evaluation time of cfg4- / cfg5-shaped data terms (random data, no regularizer):
how the tile kernel does with 2 and 4 frames per sub-pixel phase.  SRB_NO_TABLE=1 forces the generic
residual pass for an A/B comparison.
    python tools/bench_configs.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
srb = importlib.import_module("super-resolution_b200")
wl = importlib.import_module("super-resolution_b200.workloads")

for name, (C, H, W, N, s, K, sigma) in {"cfg4-shaped 1024x1024x16 N=8 s=2 K=5": (16, 1024, 1024, 8, 2, 5, 1.5),
                                        "cfg5-shaped 2048x2048x3 N=64 s=4 K=9": (3, 2048, 2048, 64, 4, 9, 2.5)}.items():
    rng = np.random.default_rng(1)
    lr = rng.random((N, C, H // s, W // s))
    x = rng.random(C * H * W)
    with srb.Engine(lr.shape, s, wl.gaussian_psf(K, sigma), wl.default_shifts(N, s)) as e:
        e.set_observations(lr)
        stream = torch.cuda.ExternalStream(e.stream_handle())
        with torch.cuda.stream(stream):
            xd = torch.from_numpy(x).cuda()
            gc = torch.zeros(x.size + 1, dtype=torch.float64, device="cuda")
            for _ in range(5):
                e.eval_partial_dev(xd, gc)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(50):
                e.eval_partial_dev(xd, gc)
            b.record(stream)
            stream.synchronize()
            ms = a.elapsed_time(b) / 50
            bytes_alg = wl.algorithmic_bytes(H, W, C, N, s, has_reg=False)
            print("%s: %.3f ms/eval, %.0f GB/s algorithmic, cost %.6e, table=%s" %
                  (name, ms, bytes_alg / ms / 1e6, float(gc[-1]), os.environ.get("SRB_NO_TABLE") is None), flush=True)
