"""GPU parity of the opt-in "Z layout" variant of the fused tile kernel (k_tile_z, SRB_ZLAYOUT=1):
observations gathered once onto the HR grid, residual pass elementwise from a TMA box.  Same bar
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
    """Fractional shifts or several frames per phase: the knob is ignored, the default kernels run."""
    wl = import_module("super-resolution_b200.workloads")
    rng = np.random.default_rng(3)
    psf = wl.gaussian_psf(5, 1.0)
    lr = rng.random((8, 1, 40, 64))
    with _engine(srb, True, lr, 2, psf, wl.default_shifts(8, 2)) as e:      # two frames per phase
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
