"""GPU parity of the Z-layout tile kernel (k_tile_zt, the default for integer shifts with one distinct shift per
sub-pixel phase; SRB_ZLAYOUT=0 selects k_tile for the A/B comparisons below): observations gathered once onto the HR
grid -- frames with equal shifts averaged --, residual pass elementwise from a TMA box.  Same bar
as the default fused path: cost and gradient within 1e-12 (relative L2) of the oracle, and of the
default path on the same inputs."""
import os
from importlib import import_module

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_L2 = 1e-12


@pytest.fixture(scope="module")
def srb():
    import srb200
    assert srb200.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return srb200


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def _model(K, s, sigma, C, h, w, seed):
    wl = import_module("super-resolution_b200.workloads")
    rng = np.random.default_rng(seed)
    N = s * s
    psf = wl.gaussian_psf(K, sigma)
    shifts = wl.default_shifts(N, s)   # one frame per sub-pixel phase
    lr = rng.random((N, C, h, w))
    x = rng.random((C, h * s, w * s))
    return psf, shifts, lr, x


def _engine(srb, zlayout, lr, s, psf, shifts):
    old = os.environ.get("SRB_ZLAYOUT")
    os.environ["SRB_ZLAYOUT"] = "1" if zlayout else "0"
    try:
        e = srb.Engine(lr.shape, s, psf, shifts)
    finally:
        if old is None:
            del os.environ["SRB_ZLAYOUT"]
        else:
            os.environ["SRB_ZLAYOUT"] = old
    e.set_observations(lr)
    return e


@pytest.mark.parametrize("K,s,sigma,C,h,w", [
    (7, 4, 1.5, 2, 48, 80),     # cfg3's model: 192 x 320 HR, 6 x 5 tiles, interior and border tiles
    (3, 2, 0.8, 1, 80, 160),    # K = 3
    (5, 4, 1.2, 3, 40, 64),     # K = 5
    (9, 2, 2.0, 1, 96, 128),    # K = 9 (3 CTAs / SM instantiation)
    (7, 4, 1.5, 1, 45, 70),     # HR size not a multiple of the tile: partial tiles on the right / bottom
])
def test_zlayout_matches_oracle_and_default_path(srb, oracle, K, s, sigma, C, h, w):
    psf, shifts, lr, x = _model(K, s, sigma, C, h, w, seed=K * 100 + s)
    m = oracle.Model(s, psf, shifts)
    obs_hr = oracle.upsample_observations(m, lr)
    lam = 0.01
    rng = np.random.default_rng(5)
    wts = rng.uniform(0.5, 2.0, size=x.shape)
    cost_ref, g_ref = oracle.evaluate(m, x, obs_hr, reg_kind=oracle.REG_TV, lam=lam, weights=wts)
    with _engine(srb, True, lr, s, psf, shifts) as ez, _engine(srb, False, lr, s, psf, shifts) as ed:
        assert ez.zlayout_active and not ed.zlayout_active
        for e in (ez, ed):
            e.set_regularizer(srb.REG_TV, lam)
            e.set_irls_weights(wts)
        cz, gz = ez.eval(x)
        cd, gd = ed.eval(x)
        assert abs(cz - cost_ref) <= REL_L2 * abs(cost_ref)
        assert rel_l2(gz, g_ref) <= REL_L2
        assert abs(cz - cd) <= REL_L2 * abs(cd)
        assert rel_l2(gz, gd) <= REL_L2
        # cost only, and the data term alone
        cz2, _ = ez.eval(x, want_grad=False)
        assert abs(cz2 - cost_ref) <= REL_L2 * abs(cost_ref)
        ez.set_regularizer(srb.REG_NONE, 0.0)
        cost_d, g_d = oracle.data_term(m, x, obs_hr)
        cz3, gz3 = ez.eval(x)
        assert abs(cz3 - cost_d) <= REL_L2 * abs(cost_d)
        assert rel_l2(gz3, g_d) <= REL_L2


def test_zlayout_channel_range_and_new_observations(srb, oracle):
    """The Z-layout copy follows srb_set_observations, and a channel range reads its own planes."""
    K, s, sigma, C, h, w = 7, 4, 1.5, 3, 40, 64
    psf, shifts, lr, x = _model(K, s, sigma, C, h, w, seed=11)
    m = oracle.Model(s, psf, shifts)
    with _engine(srb, True, lr, s, psf, shifts) as e:
        assert e.zlayout_active
        lr2 = np.random.default_rng(12).random(lr.shape)
        e.set_observations(lr2)
        obs_hr = oracle.upsample_observations(m, lr2)
        e.set_channel_range(1, 3)
        cost_ref, g_ref = oracle.data_term(m, x[1:3], obs_hr, channel_start=1)
        c, g = e.eval(x[1:3])
        assert abs(c - cost_ref) <= REL_L2 * abs(cost_ref)
        assert rel_l2(g, g_ref) <= REL_L2


def test_zlayout_declines_models_it_does_not_cover(srb):
    """Fractional shifts, two DIFFERENT shifts on one sub-pixel phase, unequal frame counts per phase: the
    default kernels run."""
    wl = import_module("super-resolution_b200.workloads")
    rng = np.random.default_rng(3)
    psf = wl.gaussian_psf(5, 1.0)
    lr = rng.random((8, 1, 40, 64))
    sh = wl.default_shifts(8, 2)
    sh[4:, 0] += 2.0                                                         # same phases, one LR pixel further
    with _engine(srb, True, lr, 2, psf, sh) as e:
        assert not e.zlayout_active
    lr = rng.random((5, 1, 40, 64))
    with _engine(srb, True, lr, 2, psf, wl.default_shifts(5, 2)) as e:      # phase (0, 0) twice, the others once
        assert not e.zlayout_active
    lr = rng.random((4, 1, 40, 64))
    sh = np.array([[0.0, 0.0], [1.0, 0.0], [0.5, 1.0], [1.0, 1.0]])
    with _engine(srb, True, lr, 2, psf, sh) as e:                            # a fractional shift
        assert not e.zlayout_active


@pytest.mark.parametrize("frames", [[0, 1, 2, 3, 4, 5, 6, 7], [8, 9, 10, 11], [3, 12]])
def test_zlayout_with_empty_phases_matches_oracle(srb, oracle, frames):
    """A frame shard of cfg3's model (some sub-pixel phases carry no frame): SRB_ZLAYOUT=2 keeps the Z
    layout with NaN holes; cost and gradient of the shard against the oracle."""
    K, s, sigma, C, h, w = 7, 4, 1.5, 2, 48, 80
    psf, shifts, lr, x = _model(K, s, sigma, C, h, w, seed=21)
    shifts, lr = shifts[frames], np.ascontiguousarray(lr[frames])
    m = oracle.Model(s, psf, shifts)
    obs_hr = oracle.upsample_observations(m, lr)
    cost_ref, g_ref = oracle.evaluate(m, x, obs_hr, reg_kind=oracle.REG_TV, lam=0.01, weights=np.ones_like(x))
    old = os.environ.get("SRB_ZLAYOUT")
    os.environ["SRB_ZLAYOUT"] = "2"
    try:
        e = srb.Engine(lr.shape, s, psf, shifts)
    finally:
        if old is None:
            del os.environ["SRB_ZLAYOUT"]
        else:
            os.environ["SRB_ZLAYOUT"] = old
    with e:
        e.set_observations(lr)
        assert e.zlayout_active
        e.set_regularizer(srb.REG_TV, 0.01)
        c, g = e.eval(x)
    assert abs(c - cost_ref) <= REL_L2 * abs(cost_ref)
    assert rel_l2(g, g_ref) <= REL_L2


@pytest.mark.parametrize("K,s,sigma,C,h,w,N", [
    (5, 2, 1.5, 3, 80, 96, 8),      # cfg4's model: 2 frames on every sub-pixel phase
    (9, 4, 2.5, 2, 40, 48, 64),     # cfg5's model: 4 frames on every phase
    (7, 4, 2.0, 1, 48, 80, 24),     # 3 frames on 8 of the 16 phases, none on the others
    (3, 2, 0.8, 1, 45, 70, 12),     # 3 per phase, HR size not a multiple of the tile
])
def test_frames_with_equal_shifts_are_merged_at_upload(srb, oracle, K, s, sigma, C, h, w, N):
    """Several frames with the same shift: k_tile_zt runs on their mean, n ||A x - m||^2 + the constant
    sum_e ||y_e - m||^2 (srb_kernels_tilez.cuh).  Cost and gradient against the oracle, which evaluates every
    frame separately (objective_data_term.cpp:104-114), and against k_tile (SRB_ZLAYOUT=0), which does too."""
    wl = import_module("super-resolution_b200.workloads")
    rng = np.random.default_rng(K * 1000 + N)
    psf = wl.gaussian_psf(K, sigma)
    if N == 24:
        base = wl.default_shifts(16, s)[:8]
        shifts = np.concatenate([base, base, base])
    else:
        shifts = wl.default_shifts(N, s)
    x = rng.random((C, h * s, w * s))
    # observations the way a solve sees them: a common signal plus noise (the merged form must not lose digits
    # when the residual is small against the signal)
    common = rng.random((1, C, h, w))
    lr = common + 0.01 * rng.standard_normal((N, C, h, w))
    wts = rng.uniform(0.5, 2.0, size=x.shape)
    m = oracle.Model(s, psf, shifts)
    obs_hr = oracle.upsample_observations(m, lr)
    p = srb.plan(lr.shape, s, psf, shifts)
    assert p["zt_frames"] == (3 if N in (24, 12) else N // (s * s)), p
    cost_ref, g_ref = oracle.evaluate(m, x, obs_hr, reg_kind=oracle.REG_TV, lam=0.01, weights=wts)
    with _engine(srb, True, lr, s, psf, shifts) as ez, _engine(srb, False, lr, s, psf, shifts) as ed:
        assert ez.zlayout_active and not ed.zlayout_active
        for e in (ez, ed):
            e.set_regularizer(srb.REG_TV, 0.01)
            e.set_irls_weights(wts)
        cz, gz = ez.eval(x)
        cd, gd = ed.eval(x)
        assert abs(cz - cost_ref) <= REL_L2 * abs(cost_ref)
        assert rel_l2(gz, g_ref) <= REL_L2
        assert abs(cz - cd) <= REL_L2 * abs(cd) and rel_l2(gz, gd) <= REL_L2
        cz2, _ = ez.eval(x, want_grad=False)
        assert abs(cz2 - cost_ref) <= REL_L2 * abs(cost_ref)
        ez.set_regularizer(srb.REG_NONE, 0.0)
        if C > 1:   # the constant term is per channel
            ez.set_channel_range(1, C)
            cost_c, g_c = oracle.data_term(m, x[1:], obs_hr, channel_start=1)
            c1, g1 = ez.eval(x[1:])
            assert abs(c1 - cost_c) <= REL_L2 * abs(cost_c)
            assert rel_l2(g1, g_c) <= REL_L2
        # new observations: the mean and the constant follow
        ez.set_channel_range(0, C)
        lr2 = common + 0.02 * rng.standard_normal((N, C, h, w))
        ez.set_observations(lr2)
        cost2, g2 = oracle.data_term(m, x, oracle.upsample_observations(m, lr2))
        c2, gg2 = ez.eval(x)
        assert abs(c2 - cost2) <= REL_L2 * abs(cost2)
        assert rel_l2(gg2, g2) <= REL_L2


def test_merged_frames_near_the_solution_keep_the_cost_digits(srb, oracle):
    """Noise-free observations evaluated at the truth: the residual is ~0, the cost is dominated by rounding.
    The merged form n (b - m)^2 + sum (y_e - m)^2 is a sum of squares and must stay as small as the oracle's."""
    wl = import_module("super-resolution_b200.workloads")
    K, s, sigma, C, h, w, N = 5, 2, 1.5, 1, 64, 64, 8
    rng = np.random.default_rng(77)
    psf = wl.gaussian_psf(K, sigma)
    shifts = wl.default_shifts(N, s)
    x = wl.box_smooth(rng.random((C, h * s, w * s)))
    m = oracle.Model(s, psf, shifts)
    lr = np.stack([np.stack([oracle.forward(m, k, x[c]) for c in range(C)]) for k in range(N)])
    obs_hr = oracle.upsample_observations(m, lr)
    cost_ref, g_ref = oracle.data_term(m, x, obs_hr)
    with _engine(srb, True, lr, s, psf, shifts) as e:
        assert e.zlayout_active
        c, g = e.eval(x)
    scale = float(np.sum(obs_hr ** 2))
    assert cost_ref <= 1e-25 * scale and c <= 1e-25 * scale, (cost_ref, c, scale)
    assert np.abs(g).max() <= 1e-12 and np.abs(g_ref).max() <= 1e-12
