#!/usr/bin/env python
"""Debug aid: per-tile comparison of the HOLES Z-layout kernel against the k_tile path and the oracle."""
import os, sys
from importlib import import_module
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srb200 as srb
from oracle import sr_oracle as oracle
wl = import_module("super-resolution_b200.workloads")

def engine(mode, lr, s, psf, shifts):
    os.environ["SRB_ZLAYOUT"] = str(mode)
    e = srb.Engine(lr.shape, s, psf, shifts)
    e.set_observations(lr)
    return e

K, s, sigma, h, w = 7, 4, 1.5, 48, 80
C = int(sys.argv[1]) if len(sys.argv) > 1 else 1
TV = len(sys.argv) > 2 and sys.argv[2] == "tv"
rng = np.random.default_rng(21)
N = s * s
psf = wl.gaussian_psf(K, sigma)
shifts = wl.default_shifts(N, s)
lr = rng.random((N, C, h, w))
x = rng.random((C, h * s, w * s))
for frames in ([3, 12], list(range(8))):
    sh, l = shifts[frames], np.ascontiguousarray(lr[frames])
    m = oracle.Model(s, psf, sh)
    obs = oracle.upsample_observations(m, l)
    cr, gr = oracle.evaluate(m, x, obs, reg_kind=oracle.REG_TV, lam=0.01) if TV else oracle.data_term(m, x, obs)
    with engine(2, l, s, psf, sh) as ez, engine(0, l, s, psf, sh) as ed:
        if TV:
            ez.set_regularizer(srb.REG_TV, 0.01); ed.set_regularizer(srb.REG_TV, 0.01)
        cz, gz = ez.eval(x)
        cd, gd = ed.eval(x)
        print("frames", frames, "zactive", ez.zlayout_active, ed.zlayout_active)
        print(" cost oracle %.6f  z %.6f  default %.6f" % (cr, cz, cd))
        for name, g in (("z", gz), ("default", gd)):
            d = np.abs(g - gr)[0]
            print(" ", name, "max abs diff", d.max())
            for ty in range(0, d.shape[0], 32):
                print("   ", " ".join("%8.1e" % d[ty:ty + 32, tx:tx + 64].max() for tx in range(0, d.shape[1], 64)))
        d = np.abs(gz - gr).max(axis=0)
        bad = np.argwhere(d > 1e-9)
        if len(bad):
            print("  bad rows", np.unique(bad[:, 0])[:40], "cols", np.unique(bad[:, 1])[:40], "count", len(bad))
