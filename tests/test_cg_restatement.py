"""The device-resident CG of SURVEY 8f / N1 is a restatement of ALGLIB's mincg
(super-resolution_b200/csrc/srb_cg.h) over a vector backend.  Here its control flow is pinned on
the CPU: instantiated over host arrays with ALGLIB's summation orders (oracle/cg_host_harness.cpp),
it must reproduce the reference's own ALGLIB bit for bit -- every objective value the line searches
ask for, the iterate, the iteration / evaluation counts and the termination type -- against the
committed fixture (tests/golden/cg_golden.npz, made by tools/make_cg_golden.py) and, where
oracle/_ref is present, against ALGLIB run live."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cg_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = cg_cases.cases()


@pytest.fixture(scope="module")
def host_cg():
    path = os.path.join(ROOT, "oracle", "_build", "libsrb_cg_host.so")
    src = os.path.join(ROOT, "oracle", "cg_host_harness.cpp")
    hdr = os.path.join(ROOT, "super-resolution_b200", "csrc", "srb_cg.h")
    if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_build/libsrb_cg_host.so"])
    return C.CDLL(path)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "cg_golden.npz"))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_restatement_reproduces_alglib_fixture(host_cg, golden, case):
    name, fg, x0, kw = case
    x, rep, trace = cg_cases.run(host_cg.srbcg_host_minimize, x0, fg, **kw)
    np.testing.assert_array_equal(rep[:4], golden[name + "_report"])
    np.testing.assert_array_equal(trace, golden[name + "_trace"])
    np.testing.assert_array_equal(x, golden[name + "_x"])
    assert rep[5] == len(trace)


def test_cases_cover_the_termination_rules_and_restarts(host_cg, golden):
    kinds = {int(golden[c[0] + "_report"][2]) for c in CASES}
    assert {1, 2, 4, 5} <= kinds
    restarts = 0
    for name, fg, x0, kw in CASES:
        restarts += cg_cases.run(host_cg.srbcg_host_minimize, x0, fg, **kw)[1][4]
    assert restarts > 0


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_restatement_reproduces_alglib_live(host_cg, case):
    from oracle import sr_ref
    if not sr_ref.available():
        pytest.skip("oracle/_ref (the reference's ALGLIB) is not built here")
    name, fg, x0, kw = case
    xr, rr, tr = cg_cases.run(sr_ref.lib().ref_mincg, x0, fg, **kw)
    xh, rh, th = cg_cases.run(host_cg.srbcg_host_minimize, x0, fg, **kw)
    np.testing.assert_array_equal(rh[:4], rr[:4])
    np.testing.assert_array_equal(th, tr)
    np.testing.assert_array_equal(xh, xr)


# ---- L-BFGS (LBFGS_SOLVER, alglib_objective.cpp:111-140) -----------------------------------------
LBFGS_CASES = [(c, m) for c in CASES for m in ((5, 3) if c[0].startswith("rosenbrock10") else (5,))]


@pytest.mark.parametrize("case,m", LBFGS_CASES, ids=["lbfgs%d_%s" % (m, c[0]) for c, m in LBFGS_CASES])
def test_lbfgs_restatement_reproduces_alglib(host_cg, golden, case, m):
    """lbfgs_minimize over host arrays against ALGLIB's minlbfgs: the committed fixture always, ALGLIB
    live where oracle/_ref is present."""
    name, fg, x0, kw = case
    x, rep, trace = cg_cases.run(host_cg.srbcg_host_lbfgs, x0, fg, lbfgs_m=m, **kw)
    key = "lbfgs%d_%s" % (m, name)
    np.testing.assert_array_equal(rep[:4], golden[key + "_report"])
    np.testing.assert_array_equal(trace, golden[key + "_trace"])
    np.testing.assert_array_equal(x, golden[key + "_x"])
    from oracle import sr_ref
    if sr_ref.available():
        xr, rr, tr = cg_cases.run(sr_ref.lib().ref_minlbfgs, x0, fg, lbfgs_m=m, **kw)
        np.testing.assert_array_equal(rep[:4], rr[:4])
        np.testing.assert_array_equal(trace, tr)
        np.testing.assert_array_equal(x, xr)


# ---- the IRLS loop and the round structure of IRLSMapSolver::Solve ------------------------------
RW = C.CFUNCTYPE(None, C.c_longlong, C.POINTER(C.c_double), C.c_void_p)


def _host_irls_round_solver(host_cg, oracle, model, obs_hr, reg_kind, lam, lbfgs_corrections=0):
    """round_solver for solver.solve_rounds: srb_cg.h's irls_solve over host arrays, the objective
    and the re-weighting being the oracle's (data term + IRLS-weighted regularizer)."""
    fn = host_cg.srbcg_host_irls
    fn.restype = C.c_int
    fn.argtypes = [C.c_longlong, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                   C.c_double, C.c_int, C.c_int, cg_cases.FG, RW, C.c_void_p, C.POINTER(C.c_double)]

    def round_solver(c0, c1, x0_slice, opt):
        Cn, H, W = x0_slice.shape
        state = {"w": np.ones((Cn, H, W))}                      # irls_map_solver.cpp:66-74
        obs = np.ascontiguousarray(obs_hr[:, c0:c1])

        def fg(n, xp, fp, gp, user):
            x = np.ctypeslib.as_array(xp, (n,)).reshape(Cn, H, W)
            f, g = oracle.evaluate(model, x, obs, reg_kind=reg_kind, lam=lam, weights=state["w"])
            fp[0] = f
            np.ctypeslib.as_array(gp, (n,))[:] = g.ravel()

        def rw(n, xp, user):
            x = np.ctypeslib.as_array(xp, (n,)).reshape(Cn, H, W)
            state["w"] = oracle.reweight(reg_kind, x)

        x = np.array(x0_slice, dtype=np.float64).reshape(-1)
        rep = np.zeros(5)
        fn(x.size, x.ctypes.data_as(C.POINTER(C.c_double)), opt.gradient_norm_threshold,
           opt.cost_decrease_threshold, opt.parameter_variation_threshold, opt.max_num_solver_iterations,
           opt.max_num_irls_iterations, opt.irls_cost_difference_threshold, 1 if lam > 0 else 0,
           lbfgs_corrections, cg_cases.FG(fg), RW(rw), None, rep.ctypes.data_as(C.POINTER(C.c_double)))
        return x.reshape(Cn, H, W), rep
    return round_solver


@pytest.mark.parametrize("split_channels,lbfgs", [(False, False), (True, False), (False, True)])
def test_irls_solve_mirror_reproduces_the_reference_solver(host_cg, oracle, ref, split_channels, lbfgs):
    """solver.solve_rounds + srb_cg.h's irls_solve + cg_minimize (host backend, oracle objective)
    against the reference's IRLSMapSolver::Solve (oracle/_ref: irls_map_solver.cpp, ALGLIB, the
    reference's TV regularizer; data term = oracle): BASELINE configuration 1's shape and defaults
    (28x28x3, 2x, 3x3 PSF, 4 frames, TV lambda 0.01, 20 IRLS x 50 CG), bit for bit."""
    from importlib import import_module
    solver = import_module("super-resolution_b200.solver")
    wl = import_module("super-resolution_b200.workloads")
    s, lam = 2, 0.01
    psf = oracle.gaussian_psf(3, 1.0)
    shifts = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], dtype=np.float64)
    m = oracle.Model(s, psf, shifts)
    truth = wl.ground_truth(28, 28, 3, 11)
    lr = np.stack([[oracle.forward(m, k, truth[c]) for c in range(3)] for k in range(4)])
    x0 = wl.bilinear_upsample(lr[0], s)
    opt_ref = ref.default_options()
    opt_ref.split_channels = 1 if split_channels else 0
    opt_ref.solver = 1 if lbfgs else 0                       # LBFGS_SOLVER, 5 correction pairs
    expect, _ = ref.solve(m, lr, x0, reg_kind=oracle.REG_TV, lam=lam, options=opt_ref)
    mine = solver.IrlsMapSolverOptions(split_channels=split_channels)
    obs_hr = oracle.upsample_observations(m, lr)
    got, reports = solver.solve_rounds(
        _host_irls_round_solver(host_cg, oracle, m, obs_hr, oracle.REG_TV, lam,
                                lbfgs_corrections=opt_ref.num_lbfgs_hessian_corrections if lbfgs else 0),
        x0, mine, regularization_parameter_sum=lam)
    assert len(reports) == (3 if split_channels else 1)
    assert all(r[0] >= 1 for r in reports)
    np.testing.assert_array_equal(got, expect)
