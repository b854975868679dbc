"""GPU tests of the device-resident solver on several devices (srb_multi_cg_minimize, srb_multi_lbfgs_minimize,
srb_multi_solve_irls; csrc/srb_multi_solver.cuh): RunCGSolverAnalyticalDiff / RunLBFGSSolverAnalyticalDiff
(alglib_objective.cpp:47-140) and IRLSMapSolver::RunIRLSLoop (irls_map_solver.cpp:45-157) with every solver vector
cut into the row bands of SRB_PARTITION_ROWS, one host thread.

The single-device solver is pinned against the reference's ALGLIB (tests/test_gpu_cg.py); the G-device solver runs
the same template over the same kernels and the same objective and differs only in the association of its
reductions (per-device sums added in device order), so its iterates must follow the single-device ones to rounding
and its iteration / evaluation counts and termination types must be identical.

`placement = shared`: all G contexts live on GPU 0 (SRB_MULTI_SHARE_DEVICES=1) -- every line of the multi-device
logic (ranges, halo pulls, events, per-device reductions, host-side sums) runs on a one-GPU box;
`placement = distinct`: G physical GPUs (skipped when the box has fewer)."""
from importlib import import_module

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
wl = import_module("super-resolution_b200.workloads")
FOLLOW_REL_L2 = 1e-8   # tests/test_gpu_cg.py's bar for "same algorithm, other reduction order"


@pytest.fixture(scope="module")
def srb():
    import srb200
    assert srb200.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return srb200


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def _problem(oracle, N, s, K, sigma, C, h, w, seed, frac=False):
    """A consistent super-resolution problem: observations of a smooth ground truth + noise, bilinear start."""
    rng = np.random.default_rng(seed)
    psf = wl.gaussian_psf(K, sigma)
    shifts = wl.default_shifts(N, s)
    if frac:
        shifts = shifts + rng.uniform(-0.4, 0.4, size=shifts.shape)
    m = oracle.Model(s, psf, shifts)
    truth = wl.ground_truth(h * s, w * s, C, seed)
    lr = np.stack([[oracle.forward(m, k, truth[c]) for c in range(C)] for k in range(N)])
    lr = lr + 0.005 * rng.standard_normal(lr.shape)
    x0 = wl.bilinear_upsample(lr[0], s)
    return psf, shifts, lr, x0


def _multi(srb, monkeypatch, lr_shape, s, psf, shifts, G, placement):
    if placement == "shared":
        monkeypatch.setenv("SRB_MULTI_SHARE_DEVICES", "1")
        devices = [0] * G
    else:
        monkeypatch.delenv("SRB_MULTI_SHARE_DEVICES", raising=False)
        if srb.device_count() < G:
            pytest.skip("needs %d GPUs" % G)
        devices = list(range(G))
    return srb.MultiEngine(lr_shape, s, psf, shifts, n_gpus=G, devices=devices, partition=srb.PARTITION_ROWS)


CASES = {
    # name: (N, s, K, sigma, C, h, w, regularizer, fractional shifts)
    "cfg3_tv": (16, 4, 7, 1.5, 2, 48, 80, "tv", False),          # 12 units of 32 rows, halo 7 rows
    "cfg2_btv_band": (9, 4, 5, 1.2, 1, 40, 64, "btv", False),    # empty phases, border band, BTV halo
    "cfg5_btv": (16, 4, 7, 1.5, 3, 72, 80, "btv", False),        # BTV by the tiled kernel, last tile row short
    "cfg4_merged": (32, 4, 5, 1.2, 2, 64, 64, "tv", False),      # two frames per phase, merged at upload
    "frac_none": (6, 2, 3, 0.8, 1, 64, 96, "none", True),        # fractional shifts (k_tile), no regularizer
    "two_units": (4, 2, 3, 0.8, 1, 24, 40, "tv", False),         # 2 units: devices beyond the second own nothing
}


def _configure(srb, e, lr, reg, lam=0.01):
    e.set_observations(lr)
    kind = {"tv": srb.REG_TV, "tv3d": srb.REG_TV3D, "btv": srb.REG_BTV, "none": srb.REG_NONE}[reg]
    e.set_regularizer(kind, lam if reg != "none" else 0.0)


@pytest.mark.parametrize("placement", ["shared", "distinct"])
@pytest.mark.parametrize("G", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("case", sorted(CASES))
def test_multi_cg_follows_the_single_device_solver(srb, oracle, monkeypatch, case, G, placement):
    if placement == "distinct" and G == 1:
        pytest.skip("G = 1 is covered by the shared placement")
    N, s, K, sigma, C, h, w, reg, frac = CASES[case]
    psf, shifts, lr, x0 = _problem(oracle, N, s, K, sigma, C, h, w, seed=11 + N + K, frac=frac)
    kw = dict(epsg=1e-9, epsf=0.0, epsx=0.0, maxits=8)
    with srb.Engine(lr.shape, s, psf, shifts) as e1:
        _configure(srb, e1, lr, reg)
        x1, r1 = e1.cg_minimize(x0, **kw)
        c1, _ = e1.eval(x1, want_grad=False)
    with _multi(srb, monkeypatch, lr.shape, s, psf, shifts, G, placement) as me:
        _configure(srb, me, lr, reg)
        xm, rm = me.cg_minimize(x0, **kw)
        xm2, rm2 = me.cg_minimize(x0, **kw)     # a second solve re-uses the workspace
        cm, _ = me.eval(xm, want_grad=False)
    print("%s G=%d %s: rel L2 vs one device %.3e; iterations %d / %d, evaluations %d / %d, cost %.12g / %.12g"
          % (case, G, placement, rel_l2(xm, x1), rm["iterations"], r1["iterations"], rm["num_evaluations"],
             r1["num_evaluations"], rm["final_cost"], r1["final_cost"]))
    assert rm["iterations"] == r1["iterations"] and rm["num_evaluations"] == r1["num_evaluations"]
    assert rm["termination_type"] == r1["termination_type"]
    assert rel_l2(xm, x1) <= FOLLOW_REL_L2
    assert abs(rm["final_cost"] - r1["final_cost"]) <= 1e-10 * abs(r1["final_cost"])
    assert abs(cm - c1) <= 1e-10 * abs(c1)
    assert r1["final_cost"] < e_cost(srb, lr, s, psf, shifts, reg, x0)         # the solve did minimise
    assert np.array_equal(xm2, xm) and rm2 == rm                              # deterministic


def e_cost(srb, lr, s, psf, shifts, reg, x):
    with srb.Engine(lr.shape, s, psf, shifts) as e:
        _configure(srb, e, lr, reg)
        return e.eval(x, want_grad=False)[0]


@pytest.mark.parametrize("placement", ["shared", "distinct"])
@pytest.mark.parametrize("G", [2, 3])
def test_multi_lbfgs_follows_the_single_device_solver(srb, oracle, monkeypatch, G, placement):
    N, s, K, sigma, C, h, w, reg, frac = CASES["cfg3_tv"]
    psf, shifts, lr, x0 = _problem(oracle, N, s, K, sigma, C, h, w, seed=5)
    kw = dict(corrections=3, epsg=1e-9, maxits=7)
    with srb.Engine(lr.shape, s, psf, shifts) as e1:
        _configure(srb, e1, lr, reg)
        x1, r1 = e1.lbfgs_minimize(x0, **kw)
    with _multi(srb, monkeypatch, lr.shape, s, psf, shifts, G, placement) as me:
        _configure(srb, me, lr, reg)
        xm, rm = me.lbfgs_minimize(x0, **kw)
    print("L-BFGS G=%d %s: rel L2 vs one device %.3e, iterations %d / %d" % (G, placement, rel_l2(xm, x1), rm["iterations"], r1["iterations"]))
    assert rm["iterations"] == r1["iterations"] and rm["termination_type"] == r1["termination_type"]
    assert rel_l2(xm, x1) <= FOLLOW_REL_L2


@pytest.mark.parametrize("placement", ["shared", "distinct"])
@pytest.mark.parametrize("G", [2, 4])
@pytest.mark.parametrize("case", ["cfg3_tv", "cfg2_btv_band"])
def test_multi_irls_follows_the_single_device_solver(srb, oracle, monkeypatch, case, G, placement):
    """Three IRLS rounds: every device re-weights from its own rows of the estimate (+ halo)."""
    N, s, K, sigma, C, h, w, reg, frac = CASES[case]
    psf, shifts, lr, x0 = _problem(oracle, N, s, K, sigma, C, h, w, seed=21)
    kw = dict(epsg=1e-9, maxits=6, max_irls_iterations=3, irls_cost_difference_threshold=1e-12)
    with srb.Engine(lr.shape, s, psf, shifts) as e1:
        _configure(srb, e1, lr, reg)
        x1, r1 = e1.solve_irls(x0, **kw)
    with _multi(srb, monkeypatch, lr.shape, s, psf, shifts, G, placement) as me:
        _configure(srb, me, lr, reg)
        xm, rm = me.solve_irls(x0, **kw)
    print("IRLS %s G=%d %s: rel L2 vs one device %.3e, reports %s / %s" % (case, G, placement, rel_l2(xm, x1), rm, r1))
    assert rm["num_irls_iterations"] == r1["num_irls_iterations"] == 3
    assert rm["num_solver_iterations"] == r1["num_solver_iterations"]
    assert rel_l2(xm, x1) <= 1e-7
    assert abs(rm["final_cost"] - r1["final_cost"]) <= 1e-9 * abs(r1["final_cost"])


def test_multi_solver_channel_range_and_fallbacks(srb, oracle, monkeypatch):
    """split_channels (a channel sub-range), 3-D TV (device 0 solves alone) and a frame-sharded context (refused)."""
    N, s, K, sigma, C, h, w = 8, 2, 5, 1.0, 3, 40, 64
    psf, shifts, lr, x0 = _problem(oracle, N, s, K, sigma, C, h, w, seed=8)
    kw = dict(epsg=1e-9, maxits=5)
    with srb.Engine(lr.shape, s, psf, shifts) as e1, _multi(srb, monkeypatch, lr.shape, s, psf, shifts, 3, "shared") as me:
        for e in (e1, me):
            _configure(srb, e, lr, "tv")
            e.set_channel_range(1, 3)
        x1, r1 = e1.cg_minimize(x0[1:], **kw)
        xm, rm = me.cg_minimize(x0[1:], **kw)
        assert rm["iterations"] == r1["iterations"] and rel_l2(xm, x1) <= FOLLOW_REL_L2
        for e in (e1, me):
            e.set_channel_range(0, 3)
            e.set_regularizer(srb.REG_TV3D, 0.01)
        x1, r1 = e1.cg_minimize(x0, **kw)
        xm, rm = me.cg_minimize(x0, **kw)
        assert rm == r1 and np.array_equal(xm, x1)     # the same single-device code path
    monkeypatch.setenv("SRB_MULTI_SHARE_DEVICES", "1")
    with srb.MultiEngine(lr.shape, s, psf, shifts, n_gpus=2, devices=[0, 0], partition=srb.PARTITION_FRAMES) as mf:
        _configure(srb, mf, lr, "tv")
        with pytest.raises(srb.SrbError):
            mf.cg_minimize(x0, **kw)
    with _multi(srb, monkeypatch, lr.shape, s, psf, shifts, 2, "shared") as me:
        with pytest.raises(srb.SrbError):               # no observations yet
            me.cg_minimize(x0, **kw)
        _configure(srb, me, lr, "tv")
        with pytest.raises(srb.SrbError):               # mincgsetcond asserts non-negative thresholds
            me.cg_minimize(x0, epsg=-1.0)


def test_irls_map_solver_solve_on_several_devices(srb, oracle, monkeypatch):
    """IRLSMapSolver::Solve (irls_map_solver.cpp:192-265; solver.solve: threshold scaling, one round per channel with
    split_channels) on a multi-device context: the same rounds as on one device, every round one srb_multi_solve_irls."""
    solver = import_module("super-resolution_b200.solver")
    N, s, K, sigma, C, h, w, reg, frac = CASES["cfg3_tv"]
    psf, shifts, lr, x0 = _problem(oracle, N, s, K, sigma, 3, h, w, seed=33)
    lam = 0.01
    for split in (False, True):
        opt = solver.IrlsMapSolverOptions(max_num_solver_iterations=5, max_num_irls_iterations=2, split_channels=split)
        with srb.Engine(lr.shape, s, psf, shifts) as e1:
            _configure(srb, e1, lr, reg, lam)
            x1, rep1 = solver.solve(e1, x0, opt, regularization_parameter_sum=lam)
        with _multi(srb, monkeypatch, lr.shape, s, psf, shifts, 3, "shared") as me:
            _configure(srb, me, lr, reg, lam)
            xm, repm = solver.solve(me, x0, opt, regularization_parameter_sum=lam)
        assert len(repm) == len(rep1) == (3 if split else 1)
        for a, b in zip(repm, rep1):
            assert a["num_irls_iterations"] == b["num_irls_iterations"]
            assert a["num_solver_iterations"] == b["num_solver_iterations"]
        print("IRLSMapSolver::Solve split_channels=%s on 3 contexts: rel L2 vs one device %.3e" % (split, rel_l2(xm, x1)))
        assert rel_l2(xm, x1) <= 1e-7
