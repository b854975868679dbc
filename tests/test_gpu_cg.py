"""GPU tests of the device-resident solver (srb_cg_minimize, srb_solve_irls).

The control flow is pinned on the CPU (tests/test_cg_restatement.py: bit-identical to the
reference's ALGLIB).  What needs a GPU is the CUDA vector backend (csrc/srb_cg_device.cuh), whose
reductions sum in a different order: iterates agree with ALGLIB driven by the same device objective
to rounding, not bit for bit.

First run on a B200 in round 2 (gpurun_out/r2_cg_tests.log): CG 8e-16 / 5e-16, L-BFGS 7e-16 relative L2
against ALGLIB on the same device objective, same iteration / evaluation counts and termination."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

import cg_cases

pytestmark = pytest.mark.gpu
wl = importlib.import_module("super-resolution_b200.workloads")
solver = importlib.import_module("super-resolution_b200.solver")
CG_REL_L2 = 1e-8        # same algorithm, same objective kernels; only the reduction order differs
SOLVER_REL_L2 = 1e-4    # north_star's bar for solver output


@pytest.fixture(scope="module")
def srb():
    import srb200
    assert srb200.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return srb200


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def _problem(oracle, C_=2, h=24, w=40, s=2, seed=3):
    rng = np.random.default_rng(seed)
    N = s * s
    psf = wl.gaussian_psf(3, 0.8)
    shifts = wl.default_shifts(N, s)
    m = oracle.Model(s, psf, shifts)
    truth = wl.ground_truth(h * s, w * s, C_, seed)
    lr = np.stack([[oracle.forward(m, k, truth[c]) for c in range(C_)] for k in range(N)])
    lr = lr + 0.005 * rng.standard_normal(lr.shape)
    x0 = wl.bilinear_upsample(lr[0], s)
    return psf, shifts, lr, x0


@pytest.mark.parametrize("reg", ["none", "tv"])
def test_device_cg_follows_alglib_on_the_same_objective(srb, oracle, ref, reg):
    """ALGLIB's mincg (the reference's own, oracle/_ref) with the device objective behind a host
    callback, against srb_cg_minimize: same thresholds, same start."""
    psf, shifts, lr, x0 = _problem(oracle)
    with srb.Engine(lr.shape, 2, psf, shifts) as e:
        e.set_observations(lr)
        if reg == "tv":
            e.set_regularizer(srb.REG_TV, 0.01)
        kw = dict(epsg=1e-7, epsf=1e-12, epsx=1e-10, maxits=30)

        def fg(x):
            f, g = e.eval(x)
            return f, g.ravel()
        xa, ra, _ = cg_cases.run(ref.lib().ref_mincg, x0.ravel(), fg, **kw)
        xd, rd = e.cg_minimize(x0, **kw)
    print("device CG (%s): rel L2 vs ALGLIB %.3e; iterations %d / %d, evaluations %d / %d, termination %d / %d"
          % (reg, rel_l2(xd.ravel(), xa), rd["iterations"], ra[0], rd["num_evaluations"], ra[1],
             rd["termination_type"], ra[2]))
    assert rel_l2(xd.ravel(), xa) <= CG_REL_L2
    assert rd["iterations"] == int(ra[0]) and rd["termination_type"] == int(ra[2])
    assert abs(rd["final_cost"] - ra[3]) <= 1e-10 * abs(ra[3])


def test_device_lbfgs_follows_alglib_on_the_same_objective(srb, oracle, ref):
    """ALGLIB's minlbfgs (5 pairs) with the device objective behind a host callback against
    srb_lbfgs_minimize."""
    psf, shifts, lr, x0 = _problem(oracle)
    with srb.Engine(lr.shape, 2, psf, shifts) as e:
        e.set_observations(lr)
        e.set_regularizer(srb.REG_TV, 0.01)
        kw = dict(epsg=1e-7, epsf=1e-12, epsx=1e-10, maxits=25)

        def fg(x):
            f, g = e.eval(x)
            return f, g.ravel()
        xa, ra, _ = cg_cases.run(ref.lib().ref_minlbfgs, x0.ravel(), fg, lbfgs_m=5, **kw)
        xd, rd = e.lbfgs_minimize(x0, corrections=5, **kw)
    print("device L-BFGS: rel L2 vs ALGLIB %.3e; iterations %d / %d, termination %d / %d"
          % (rel_l2(xd.ravel(), xa), rd["iterations"], ra[0], rd["termination_type"], ra[2]))
    assert rel_l2(xd.ravel(), xa) <= CG_REL_L2
    assert rd["iterations"] == int(ra[0]) and rd["termination_type"] == int(ra[2])


def test_device_irls_solve_matches_reference_loop(srb, oracle, ref):
    """IRLSMapSolver::Solve: the reference's loop + ALGLIB with the device objective (ref_solve_fused)
    against the fully device-resident solve, tie-free image, 2 IRLS rounds of 25 CG iterations."""
    psf, shifts, lr, x0 = _problem(oracle, C_=3, h=14, w=14)
    lam = 0.01
    opt = ref.default_options()
    opt.max_num_solver_iterations = 25
    opt.max_num_irls_iterations = 2
    mine = solver.IrlsMapSolverOptions(max_num_solver_iterations=25, max_num_irls_iterations=2)
    with srb.Engine(lr.shape, 2, psf, shifts) as e:
        e.set_observations(lr)
        e.set_regularizer(srb.REG_TV, lam)
        expect, _ = ref.solve_fused(e, x0, True, lam, options=opt)
        got, reports = solver.solve(e, x0, mine, regularization_parameter_sum=lam)
    print("device IRLS solve: rel L2 vs reference loop %.3e; reports %s" % (rel_l2(got, expect), reports))
    assert rel_l2(got, expect) <= SOLVER_REL_L2
    assert reports[0]["num_irls_iterations"] == 2


def test_device_cg_rejects_bad_options(srb, oracle):
    psf, shifts, lr, x0 = _problem(oracle, C_=1, h=12, w=16)
    with srb.Engine(lr.shape, 2, psf, shifts) as e:
        with pytest.raises(srb.SrbError):          # no observations yet
            e.cg_minimize(x0)
        e.set_observations(lr)
        with pytest.raises(srb.SrbError):          # mincgsetcond asserts non-negative thresholds
            e.cg_minimize(x0, epsg=-1.0)
