#!/usr/bin/env python
"""A/B of builds of the library that differ only in semantics-preserving kernel options (cache hints, streaming
stores, load batching): the fused tile kernel at cfg3's shape (2048 x 2048 RGB, 16 frames, 4x, 7x7 PSF, TV), device
resident, torch-free; kernel time from the library's own CUDA events (srb_set_profiling) and a SHA-256 of gradient +
cost, which must be identical for every build.

    python tools/ab_variants.py base=super-resolution_b200/libsrb200.so name=path.so ...   (driver)
    python tools/ab_variants.py --one name path.so [iterations]                             (one build, one process)

The driver runs every build once in its own process (the library is loaded once per process), then the best two and
the base a second time, and prints one JSON line per run plus a summary."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one(name, path, iters):
    os.environ["SRB200_LIB"] = os.path.abspath(path)
    sys.path.insert(0, ROOT)
    import importlib
    import numpy as np
    srb = importlib.import_module("super-resolution_b200")
    wl = importlib.import_module("super-resolution_b200.workloads")
    C, H, W, N, s, K, sigma = 3, 2048, 2048, 16, 4, 7, 2.0
    rng = np.random.default_rng(1)
    lr = rng.random((N, C, H // s, W // s))
    x = rng.random(C * H * W)
    n = x.size
    with srb.Engine(lr.shape, s, wl.gaussian_psf(K, sigma), wl.default_shifts(N, s)) as e:
        e.set_observations(lr)
        e.set_regularizer(srb.REG_TV, 0.01)
        e.eval(x, want_grad=False)                # uploads x into the context's own buffer
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
        ms = np.array(ms[10:])
        print(json.dumps({"build": name, "kernel_ms_min": float(ms.min()), "kernel_ms_median": float(np.median(ms)),
                          "zlayout": bool(e.zlayout_active), "cost": float(g[n]),
                          "sha256": hashlib.sha256(g.tobytes()).hexdigest()[:16]}), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        one(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 80)
        return
    builds = [a.split("=", 1) for a in sys.argv[1:]]
    results = {}

    def run(name, path):
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name, path], capture_output=True, text=True, timeout=120)
        for line in res.stdout.splitlines():
            if line.startswith("{"):
                print(line, flush=True)
                results.setdefault(name, []).append(json.loads(line))
                return
        print(json.dumps({"build": name, "failed": (res.stderr or res.stdout)[-300:]}), flush=True)

    for name, path in builds:
        run(name, path)
    ranked = sorted((n for n in results if n != builds[0][0]), key=lambda n: results[n][0]["kernel_ms_median"])
    for name in [builds[0][0]] + ranked[:2]:
        run(name, dict(builds)[name])
    base = results.get(builds[0][0], [])
    summary = {"base": builds[0][0], "sha_all_equal": len({r["sha256"] for rs in results.values() for r in rs}) == 1,
               "median_ms": {n: [r["kernel_ms_median"] for r in rs] for n, rs in results.items()},
               "min_ms": {n: [r["kernel_ms_min"] for r in rs] for n, rs in results.items()}}
    print(json.dumps(summary), flush=True)


if __name__ == "__main__":
    main()
