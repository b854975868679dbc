#!/usr/bin/env python
"""Generates tests/golden/cg_golden.npz: the results of the REFERENCE'S OWN ALGLIB mincg (oracle/_ref,
ref_mincg: configured as RunCGSolverAnalyticalDiff, alglib_objective.cpp:47-75) on the cases of
tests/cg_cases.py.  Needs oracle/_ref (build container only); the fixture travels.
    python tools/make_cg_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cg_cases  # noqa: E402
from oracle import sr_ref  # noqa: E402

out = {}
for name, fg, x0, kw in cg_cases.cases():
    x, rep, trace = cg_cases.run(sr_ref.lib().ref_mincg, x0, fg, **kw)
    out[name + "_x"] = x
    out[name + "_report"] = rep[:4]      # iterations, nfev, termination type, final f
    out[name + "_trace"] = trace
    print("%-24s iterations %4d nfev %4d termination %2d f %.17g" % (name, rep[0], rep[1], rep[2], rep[3]))
# the same objectives through ALGLIB's minlbfgs (RunLBFGSSolverAnalyticalDiff: m = 5 by default)
for name, fg, x0, kw in cg_cases.cases():
    for m in ((5, 3) if name.startswith("rosenbrock10") else (5,)):
        x, rep, trace = cg_cases.run(sr_ref.lib().ref_minlbfgs, x0, fg, lbfgs_m=m, **kw)
        key = "lbfgs%d_%s" % (m, name)
        out[key + "_x"] = x
        out[key + "_report"] = rep[:4]
        out[key + "_trace"] = trace
        print("%-32s iterations %4d nfev %4d termination %2d f %.17g" % (key, rep[0], rep[1], rep[2], rep[3]))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cg_golden.npz"), **out)
