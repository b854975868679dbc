#!/usr/bin/env python
"""Device-resident CG solve (20 iterations, pinned host x in / out once) on 1 .. G GPUs of this box:
srb_cg_minimize on one device against srb_multi_cg_minimize (row bands, one host thread + helper threads) --
the multi-GPU form of bench.py's `solve` block, outside torchrun.

    python tools/multi_solve_timing.py [--config 3] [--gpus 1,2,4,8] [--iters 20] [--shared] [--no-threads]

--shared places all G contexts on GPU 0 (SRB_MULTI_SHARE_DEVICES=1): no speed-up to expect, measures the overhead of
the multi-device machinery on a one-GPU box.  Prints one JSON line per run."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--shared", action="store_true")
    ap.add_argument("--no-threads", action="store_true")
    ap.add_argument("--repeat", type=int, default=3, help="timed solves per configuration (all are printed)")
    args = ap.parse_args()
    if args.shared:
        os.environ["SRB_MULTI_SHARE_DEVICES"] = "1"
    if args.no_threads:
        os.environ["SRB_MULTI_THREADS"] = "0"
    srb = importlib.import_module("super-resolution_b200")
    wl = importlib.import_module("super-resolution_b200.workloads")
    ndev = srb.device_count()
    assert ndev > 0, "needs a CUDA device"
    cf = wl.CONFIGS[args.config]
    H, W, C, s, N = cf["H"], cf["W"], cf["C"], cf["s"], cf["N"]
    psf = wl.gaussian_psf(cf["K"], cf["sigma"])
    shifts = wl.default_shifts(N, s)
    shape = (N, C, H // s, W // s)
    units = wl.work_units(H, W, C, N)
    with srb.Engine(shape, s, psf, shifts, device=0) as e:
        work = wl.make(args.config, forward=lambda k, plane: e.forward(k, plane), N=N)
        e.set_observations(work["lr"])
        e.set_regularizer(work["reg_kind"], work["lam"], work["btv_range"], work["btv_decay"])
        x0 = np.ascontiguousarray(work["x0"], dtype=np.float64).reshape(-1)
        xs = x0.copy()
        srb.pin_host(xs)
        e.cg_minimize_inplace(xs, maxits=2)
        xs[:] = x0
        t0 = time.perf_counter()
        rep = e.cg_minimize_inplace(xs, maxits=args.iters)
        dt = time.perf_counter() - t0
        x_single = xs.copy()
        srb.unpin_host(xs)
        print(json.dumps({"api": "srb_cg_minimize", "config": cf["name"], "gpus": 1, "seconds": dt,
                          "ms_per_iteration": dt * 1e3 / max(rep["iterations"], 1), "iterations": rep["iterations"],
                          "evaluations": rep["num_evaluations"], "final_cost": rep["final_cost"],
                          "value": units * rep["num_evaluations"] / dt}), flush=True)
    for G in [int(g) for g in args.gpus.split(",")]:
        if not args.shared and G > ndev:
            continue
        devices = [0] * G if args.shared else list(range(G))
        with srb.MultiEngine(shape, s, psf, shifts, n_gpus=G, devices=devices, partition=srb.PARTITION_ROWS) as me:
            me.set_observations(work["lr"])
            me.set_regularizer(work["reg_kind"], work["lam"], work["btv_range"], work["btv_decay"])
            xs = x0.copy()
            srb.pin_host(xs)
            me.cg_minimize_inplace(xs, maxits=2)
            times = []
            for _ in range(max(1, args.repeat)):
                xs[:] = x0
                t0 = time.perf_counter()
                rep = me.cg_minimize_inplace(xs, maxits=args.iters)
                times.append(time.perf_counter() - t0)
            dt = min(times)
            rel = float(np.linalg.norm(xs - x_single) / np.linalg.norm(x_single))
            srb.unpin_host(xs)
            print(json.dumps({"api": "srb_multi_cg_minimize", "config": cf["name"], "gpus": G,
                              "placement": "shared" if args.shared else "distinct",
                              "threads": not args.no_threads, "seconds": dt, "all_seconds": times,
                              "ms_per_iteration": dt * 1e3 / max(rep["iterations"], 1), "iterations": rep["iterations"],
                              "evaluations": rep["num_evaluations"], "final_cost": rep["final_cost"],
                              "value": units * rep["num_evaluations"] / dt, "rel_l2_vs_one_device": rel}), flush=True)


if __name__ == "__main__":
    main()
