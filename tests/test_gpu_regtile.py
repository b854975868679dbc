"""GPU parity of the tiled regularization kernels of the fused path (csrc/srb_kernels_regtile.cuh: k_btv_tile,
k_tv3d_tile), which replace the three reference-order launches behind the tile kernel: cost and gradient of the
whole objective against the oracle (oracle/sr_oracle.c follows btv_regularizer.cpp:19-170 and
tv_regularizer.cpp:21-227 loop for loop) within 1e-12 relative L2, on images with exact ties, for every BTV
range the kernel is instantiated for, tiles on every image border, the reference's three quirks (inclusive value
window vs exclusive gradient window, image pixel (0,0) skipped, no z part in the 3-D TV self term), HR row bands
(a rank's share of the regularization term) and channel ranges."""
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


def _problem(C, h, w, s, K, sigma, seed, ties=True):
    wl = import_module("super-resolution_b200.workloads")
    rng = np.random.default_rng(seed)
    N = s * s
    psf = wl.gaussian_psf(K, sigma)
    shifts = wl.default_shifts(N, s)
    lr = rng.random((N, C, h, w))
    x = wl.box_smooth(rng.random((C, h * s, w * s)), 3)
    if ties:   # exact ties: sgn(0) = 0 in every derivative
        x[0, 5, 7] = x[0, 5, 8]
        x[0, 6, 7] = x[0, 5, 7]
        x[0, 0, 0] = x[0, 0, 1]
        x[-1, 40, 3:9] = x[-1, 40, 3]
        if C > 1:
            x[1, 17, 20] = x[0, 17, 20]
    wts = rng.uniform(0.25, 4.0, size=x.shape)
    return psf, shifts, lr, x, wts


@pytest.mark.parametrize("R,decay", [(1, 0.25), (2, 0.5), (3, 0.5), (3, 1.0), (4, 0.7)])
def test_btv_tile_kernel_matches_oracle(srb, oracle, R, decay):
    C, h, w, s, K = 2, 40, 52, 4, 7          # 160 x 208 HR: 5 x 4 tiles, partial tiles right and below
    psf, shifts, lr, x, wts = _problem(C, h, w, s, K, 1.5, seed=10 + R)
    m = oracle.Model(s, psf, shifts)
    obs = oracle.upsample_observations(m, lr)
    lam = 0.02
    fo, go = oracle.evaluate(m, x, obs, oracle.REG_BTV, lam, wts, btv_range=R, btv_decay=decay)
    fd, gd = oracle.data_term(m, x, obs)
    with srb.Engine(lr.shape, s, psf, shifts) as e:
        e.set_observations(lr)
        e.set_regularizer(srb.REG_BTV, lam, R, decay)
        e.set_irls_weights(wts)
        assert e.active_path == srb.PATH_FUSED
        l0 = e.timing()["kernel_launches"]
        f, g = e.eval(x)
        # whole-image evaluation from device buffers: tile kernel + BTV tile kernel + finish
        assert abs(f - fo) <= REL_L2 * abs(fo)
        assert rel_l2(g, go) <= REL_L2
        # the regularization part on its own is held to the same bar (it is ~1e-2 of the gradient here)
        if R > 1:   # (R = 1: the exclusive gradient window is the pixel itself -- the derivative is identically 0)
            assert rel_l2(g - gd, go - gd) <= 1e-11
        f2, _ = e.eval(x, want_grad=False)
        assert abs(f2 - fo) <= REL_L2 * abs(fo)
        # against the reference-order kernels on the same device (bit-exact vs the reference's btv_regularizer.cpp)
        e.set_path(srb.PATH_REFERENCE_ORDER)
        fr, gr = e.eval(x)
        assert abs(f - fr) <= REL_L2 * abs(fr) and rel_l2(g, gr) <= REL_L2
        assert l0 is not None


def test_btv_pixel_zero_quirk_is_reproduced(srb, oracle):
    """btv_regularizer.cpp:143-146 skips image pixel (0,0) in the neighbour sum: the derivative at the pixels
    within R-1 of the origin lacks its contribution.  A large weight at (0,0) makes that visible."""
    C, h, w, s, K = 1, 16, 32, 2, 3
    psf, shifts, lr, x, wts = _problem(C, h, w, s, K, 0.8, seed=3, ties=False)
    wts[0, 0, 0] = 1.0e4
    m = oracle.Model(s, psf, shifts)
    obs = oracle.upsample_observations(m, lr)
    fo, go = oracle.evaluate(m, x, obs, oracle.REG_BTV, 0.05, wts, btv_range=3, btv_decay=0.5)
    with srb.Engine(lr.shape, s, psf, shifts) as e:
        e.set_observations(lr)
        e.set_regularizer(srb.REG_BTV, 0.05, 3, 0.5)
        e.set_irls_weights(wts)
        f, g = e.eval(x)
    assert abs(f - fo) <= REL_L2 * abs(fo)
    np.testing.assert_allclose(g[0, :4, :4], go[0, :4, :4], rtol=1e-11, atol=1e-13)
    assert rel_l2(g, go) <= REL_L2


@pytest.mark.parametrize("C", [1, 2, 5])
def test_tv3d_tile_kernel_matches_oracle(srb, oracle, C):
    h, w, s, K = 40, 52, 4, 5
    psf, shifts, lr, x, wts = _problem(C, h, w, s, K, 1.2, seed=20 + C)
    m = oracle.Model(s, psf, shifts)
    obs = oracle.upsample_observations(m, lr)
    lam = 0.03
    fo, go = oracle.evaluate(m, x, obs, oracle.REG_TV3D, lam, wts)
    fd, gd = oracle.data_term(m, x, obs)
    with srb.Engine(lr.shape, s, psf, shifts) as e:
        e.set_observations(lr)
        e.set_regularizer(srb.REG_TV3D, lam)
        e.set_irls_weights(wts)
        assert e.active_path == srb.PATH_FUSED
        f, g = e.eval(x)
        assert abs(f - fo) <= REL_L2 * abs(fo)
        assert rel_l2(g, go) <= REL_L2
        assert rel_l2(g - gd, go - gd) <= 1e-11
        f2, _ = e.eval(x, want_grad=False)
        assert abs(f2 - fo) <= REL_L2 * abs(fo)
        if C >= 3:   # a channel range couples only its own channels (the regularizer sees num_channels = range)
            e.set_channel_range(1, 4)
            e.set_irls_weights(wts[1:4])
            fo2, go2 = oracle.evaluate(m, x[1:4], obs[:, 1:4].copy(), oracle.REG_TV3D, lam, wts[1:4])
            f3, g3 = e.eval(x[1:4])
            assert abs(f3 - fo2) <= REL_L2 * abs(fo2)
            assert rel_l2(g3, go2) <= REL_L2


@pytest.mark.parametrize("kind", ["btv", "tv3d"])
def test_regularizer_row_bands_sum_to_the_whole(srb, oracle, kind):
    """Multi-GPU partition: every rank evaluates the regularization term for its HR row band only;
    bands that cut through tiles, summed, give the whole term."""
    C, h, w, s, K = 3, 40, 36, 4, 7
    psf, shifts, lr, x, wts = _problem(C, h, w, s, K, 1.5, seed=31)
    m = oracle.Model(s, psf, shifts)
    obs = oracle.upsample_observations(m, lr)
    okind = oracle.REG_BTV if kind == "btv" else oracle.REG_TV3D
    skind = srb.REG_BTV if kind == "btv" else srb.REG_TV3D
    fo, go = oracle.evaluate(m, x, obs, okind, 0.02, wts)
    fd, gd = oracle.data_term(m, x, obs)
    H = h * s
    bands = [(0, 37), (37, 96), (96, 96), (96, H)]
    with srb.Engine(lr.shape, s, psf, shifts) as e:
        e.set_observations(lr)
        fd_dev, gd_dev = e.eval(x)              # data term alone, this device's own bits
        gd_dev = gd_dev.copy()
        assert abs(fd_dev - fd) <= REL_L2 * abs(fd) and rel_l2(gd_dev, gd) <= REL_L2
        e.set_regularizer(skind, 0.02, 3, 0.5)
        e.set_irls_weights(wts)
        total_f, total_g = 0.0, np.zeros_like(x)
        for r0, r1 in bands:
            e.set_regularizer_rows(r0, r1)
            f, g = e.eval(x)
            total_f += f - fd_dev
            total_g += g - gd_dev
            # outside its band a rank adds nothing to the gradient
            assert np.abs(g - gd_dev)[:, :r0].max(initial=0.0) == 0.0
            assert np.abs(g - gd_dev)[:, r1:].max(initial=0.0) == 0.0
    assert abs(total_f - (fo - fd)) <= 1e-10 * abs(fo - fd)
    assert rel_l2(total_g, go - gd) <= 1e-10


def test_host_pipelined_eval_with_btv_matches_device_eval(srb, oracle):
    """srb_eval streams x in slices while the tile + BTV kernels run on the slices that have arrived; cfg2-sized."""
    wl = import_module("super-resolution_b200.workloads")
    w = wl.make(2, cheap=True)
    m = oracle.Model(w["s"], w["psf"], w["shifts"])
    obs = oracle.upsample_observations(m, w["lr"])
    fo, go = oracle.evaluate(m, w["x0"], obs, oracle.REG_BTV, w["lam"], None, threads=8)
    with srb.Engine(w["lr"].shape, w["s"], w["psf"], w["shifts"]) as e:
        e.set_observations(w["lr"])
        e.set_regularizer(srb.REG_BTV, w["lam"], 3, 0.5)
        f, g = e.eval(w["x0"])
    assert abs(f - fo) <= REL_L2 * abs(fo)
    assert rel_l2(g, go) <= REL_L2
