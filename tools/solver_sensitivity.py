#!/usr/bin/env python
"""How reproducible is the REFERENCE's own solver?  Runs the CPU reference IRLS + ALGLIB solve of
BASELINE configuration 1 (oracle/_ref: the reference's unmodified solver sources) twice: as is, and
with the data-term gradient multiplied by (1 + eps * N(0,1)) per element, eps = 1e-16 .. 1e-13.
CPU only; results are quoted in DESIGN.md section 6 and tests/test_gpu_solver.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sr_oracle as o, sr_ref as ref  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "cv2_fixtures.npz"))
lr, x0, psf, shifts = g["cfg1_lr"], g["cfg1_x0"], g["cfg1_psf"], g["cfg1_shifts"]
truth = np.moveaxis(g["fb_bgr_u8"].astype(float) / 255, 2, 0)
m = o.Model(2, psf, shifts)
obs = o.upsample_observations(m, lr)
rng = np.random.default_rng(0)


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def perturbed(eps):
    def cb(xp, gp, c0, c1, user):
        n = (c1 - c0) * 28 * 28
        x = np.ctypeslib.as_array(xp, (n,)).reshape(c1 - c0, 28, 28)
        if gp:
            gacc = np.ctypeslib.as_array(gp, (n,)).reshape(c1 - c0, 28, 28)
            tmp = np.zeros_like(gacc)
            f, _ = o.data_term(m, x, obs, grad=tmp, channel_start=c0)
            gacc += tmp * (1 + eps * rng.standard_normal(tmp.shape))
        else:
            f, _ = o.data_term(m, x, obs, want_grad=False, channel_start=c0)
        return f
    return ref.Callbacks(ref.DATA_TERM_CB(cb), ref.REG_APPLY_CB(), ref.REG_APPLY_DIFF_CB(), None)


for irls, cg in [(20, 50), (1, 50), (2, 8), (1, 5)]:
    opt = ref.default_options()
    opt.max_num_irls_iterations, opt.max_num_solver_iterations = irls, cg
    base, st = ref.solve(m, lr, x0, reg_kind=o.REG_TV, lam=0.01, options=opt)
    for eps in (1e-16, 1e-15, 1e-13):
        out, _ = ref.solve(m, lr, x0, reg_kind=o.REG_TV, lam=0.01, options=opt, callbacks=perturbed(eps))
        print("IRLS<=%2d CG<=%2d eps=%.0e: rel L2 vs unperturbed %.3e; to truth %.5f vs %.5f (%d evals)" %
              (irls, cg, eps, rel(out, base), rel(out, truth), rel(base, truth), st.num_data_term_evals))
