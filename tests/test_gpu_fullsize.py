"""Full-size GPU checks (run with `-m gpu`): BASELINE.json's configurations at their real sizes,
through size-independent properties -- the CPU oracle needs minutes per evaluation there -- plus
oracle parity on scaled-down versions of the configurations the bench does not run.

Properties (exact in real arithmetic, so they are held to fp64 round-off):
  * frame additivity: the data term is a sum over frames (objective_data_term.cpp:104-114), so two
    engines holding complementary frame blocks must add up to the engine holding all frames;
  * directional derivative: (f(x + e d) - f(x - e d)) / 2e == <g(x), d> for the quadratic data term
    (exact up to round-off for any e), and to O(e^2) with the TV term away from ties;
  * path agreement: fused tile kernel == reference-order kernels (which are bit-identical to the CPU
    reference on every case the oracle can reach).
"""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
wl = importlib.import_module("super-resolution_b200.workloads")


@pytest.fixture(scope="module")
def srb():
    import srb200
    assert srb200.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return srb200


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def _workload(cfg, **kw):
    cf = wl.CONFIGS[cfg]
    w = wl.make(cfg, cheap=True, **kw)          # decimated smooth truth + noise: shapes and statistics
    return cf, w


@pytest.mark.parametrize("cfg", [2, 3])
def test_full_size_paths_agree_and_frames_add_up(srb, cfg):
    cf, w = _workload(cfg)
    lr, x0 = w["lr"], w["x0"]
    N = lr.shape[0]
    rng = np.random.default_rng(cfg)
    wts = 0.5 + rng.random(x0.shape)
    with srb.Engine(lr.shape, w["s"], w["psf"], w["shifts"]) as e:
        e.set_observations(lr)
        e.set_regularizer(w["reg_kind"], w["lam"], w["btv_range"], w["btv_decay"])
        e.set_irls_weights(wts)
        assert e.active_path == srb.PATH_FUSED
        f_fused, g_fused = e.eval(x0)
        e.set_path(srb.PATH_REFERENCE_ORDER)
        f_ref, g_ref = e.eval(x0)
        e.set_path(srb.PATH_AUTO)
        assert abs(f_fused - f_ref) <= 1e-12 * abs(f_ref)
        assert rel(g_fused, g_ref) <= 1e-12
        # data term only: frame blocks add up
        e.set_regularizer(srb.REG_NONE, 0.0)
        f_all, g_all = e.eval(x0)
    parts = []
    for frames in (range(0, N // 2), range(N // 2, N)):
        frames = list(frames)
        with srb.Engine(lr[frames].shape, w["s"], w["psf"], w["shifts"][frames]) as e:
            e.set_observations(lr[frames])
            parts.append(e.eval(x0))
    assert abs(parts[0][0] + parts[1][0] - f_all) <= 1e-12 * abs(f_all)
    assert rel(parts[0][1] + parts[1][1], g_all) <= 1e-13


def test_full_size_directional_derivative_cfg3(srb):
    cf, w = _workload(3)
    lr, x0 = w["lr"], w["x0"]
    rng = np.random.default_rng(33)
    d = rng.standard_normal(x0.shape)
    with srb.Engine(lr.shape, w["s"], w["psf"], w["shifts"]) as e:
        e.set_observations(lr)
        f0, g = e.eval(x0)                      # data term only: exactly quadratic
        eps = 1e-3
        fp, _ = e.eval(x0 + eps * d, want_grad=False)
        fm, _ = e.eval(x0 - eps * d, want_grad=False)
        lhs, rhs = (fp - fm) / (2 * eps), float(np.vdot(g, d))
        assert abs(lhs - rhs) <= 1e-9 * abs(rhs), (lhs, rhs)
        # second difference = d^T H d > 0 and matches <g(x + eps d) - g(x - eps d), d> / (2 eps)
        _, gp = e.eval(x0 + eps * d)
        _, gm = e.eval(x0 - eps * d)
        curv = (fp - 2 * f0 + fm) / eps ** 2
        assert curv > 0
        assert abs(float(np.vdot(gp - gm, d)) / (2 * eps) - curv) <= 1e-6 * curv
        # with the TV term: O(eps^2) agreement away from ties
        e.set_regularizer(srb.REG_TV, 0.01)
        f0, g = e.eval(x0)
        eps = 1e-7
        fp, _ = e.eval(x0 + eps * d, want_grad=False)
        fm, _ = e.eval(x0 - eps * d, want_grad=False)
        lhs, rhs = (fp - fm) / (2 * eps), float(np.vdot(g, d))
        assert abs(lhs - rhs) <= 1e-5 * abs(rhs), (lhs, rhs)


@pytest.mark.parametrize("kind", [0, 1])
def test_cfg4_shaped_hyperspectral_vs_oracle(srb, oracle, kind):
    """Configuration 4 scaled down (many bands, 2x, 5x5 PSF, 8 frames; TV and 3-D TV)."""
    cf = wl.CONFIGS[4]
    rng = np.random.default_rng(44)
    C, h, w_, s, K, N = 12, 80, 104, cf["s"], cf["K"], cf["N"]   # 160 x 208 HR: has interior tiles
    psf = wl.gaussian_psf(K, cf["sigma"])
    shifts = wl.default_shifts(N, s)
    x = rng.random((C, h * s, w_ * s))
    lr = rng.random((N, C, h, w_))
    wts = 0.5 + rng.random(x.shape)
    m = oracle.Model(s, psf, shifts)
    fo, go = oracle.evaluate(m, x, oracle.upsample_observations(m, lr), kind, 0.01, wts, threads=8)
    with srb.Engine(lr.shape, s, psf, shifts) as e:
        e.set_observations(lr)
        e.set_regularizer(kind, 0.01)
        e.set_irls_weights(wts)
        f, g = e.eval(x)
        assert e.active_path == srb.PATH_FUSED
    assert abs(f - fo) <= 1e-12 * abs(fo)
    assert rel(g, go) <= 1e-12


def test_cfg5_shaped_many_frames_vs_oracle(srb, oracle):
    """Configuration 5 scaled down: 64 frames (four per sub-pixel phase), 4x, 9x9 PSF, BTV(3, 0.5)."""
    cf = wl.CONFIGS[5]
    rng = np.random.default_rng(55)
    C, h, w_, s, K, N = 3, 48, 52, cf["s"], cf["K"], cf["N"]      # 192 x 208 HR: has interior tiles
    psf = wl.gaussian_psf(K, cf["sigma"])
    shifts = wl.default_shifts(N, s)
    x = rng.random((C, h * s, w_ * s))
    lr = rng.random((N, C, h, w_))
    wts = 0.5 + rng.random(x.shape)
    m = oracle.Model(s, psf, shifts)
    fo, go = oracle.evaluate(m, x, oracle.upsample_observations(m, lr), oracle.REG_BTV, 0.01, wts, threads=8)
    with srb.Engine(lr.shape, s, psf, shifts) as e:
        e.set_observations(lr)
        e.set_regularizer(srb.REG_BTV, 0.01, 3, 0.5)
        e.set_irls_weights(wts)
        f, g = e.eval(x)
        assert e.active_path == srb.PATH_FUSED
    assert abs(f - fo) <= 1e-12 * abs(fo)
    assert rel(g, go) <= 1e-12


def _crop_check(oracle, w, g_gpu, x, wts, okind, lam, ch, r0, c0, size, margin):
    """Oracle on an HR crop [r0-margin, r0+size+margin) x [c0-margin, ...) of channel `ch`, all frames (clipped at
    the image: a crop that touches the image border keeps the real border there).  The data term and the 2-D
    regularizers are local, so away from the crop's ARTIFICIAL borders the oracle's gradient of the cropped problem
    is the gradient of the full problem: compared on [r0, r0+size) x [c0, c0+size)."""
    s = w["s"]
    H, W = x.shape[1:]
    R0, R1 = max(r0 - margin, 0), min(r0 + size + margin, H)
    C0, C1 = max(c0 - margin, 0), min(c0 + size + margin, W)
    assert R0 % s == 0 and C0 % s == 0 and R1 % s == 0 and C1 % s == 0
    xc = np.ascontiguousarray(x[ch:ch + 1, R0:R1, C0:C1])
    wc = np.ascontiguousarray(wts[ch:ch + 1, R0:R1, C0:C1])
    lrc = np.ascontiguousarray(w["lr"][:, ch:ch + 1, R0 // s:R1 // s, C0 // s:C1 // s])
    m = oracle.Model(s, w["psf"], w["shifts"])
    _, gc = oracle.evaluate(m, xc, oracle.upsample_observations(m, lrc), okind, lam, wc, btv_range=w["btv_range"],
                            btv_decay=w["btv_decay"], threads=8)
    a = g_gpu[ch, r0:r0 + size, c0:c0 + size]
    b = gc[0, r0 - R0:r0 - R0 + size, c0 - C0:c0 - C0 + size]
    return rel(a, b)


@pytest.mark.parametrize("cfg", [3, 4, 5])
def test_full_size_gradient_against_the_oracle_on_crops(srb, oracle, cfg):
    """BASELINE configurations 3, 4 and 5 at their real sizes (2048^2 x 3 x 16 frames; 1024^2 x 128 bands x 8 frames;
    4096^2 x 3 x 64 frames, BTV) evaluated on the GPU; the CPU oracle evaluates crops of the same problem -- an
    interior block, the top-left and the bottom-right image corners (real borders, incl. the border band of special
    samples) -- in first, middle and last channel.  Fused-path bar: 1e-12 relative L2 per block."""
    cf, w = _workload(cfg)
    x = w["x0"]
    rng = np.random.default_rng(100 + cfg)
    wts = 0.5 + rng.random(x.shape)
    okind = {0: oracle.REG_TV, 2: oracle.REG_BTV}[w["reg_kind"]]
    with srb.Engine(w["lr"].shape, w["s"], w["psf"], w["shifts"]) as e:
        e.set_observations(w["lr"])
        e.set_regularizer(w["reg_kind"], w["lam"], w["btv_range"], w["btv_decay"])
        e.set_irls_weights(wts)
        assert e.active_path == srb.PATH_FUSED and e.zlayout_active
        f, g = e.eval(x)
    assert np.isfinite(f) and f > 0
    H, W, Cn = cf["H"], cf["W"], cf["C"]
    size, margin = 96, 32
    blocks = [(H // 2 + 32, W // 2 - 64), (0, 0), (H - size, W - size), (0, W // 2), (H // 2, 0)]
    worst = 0.0
    for ch in sorted({0, Cn // 2, Cn - 1}):
        for (r0, c0) in blocks:
            worst = max(worst, _crop_check(oracle, w, g, x, wts, okind, w["lam"], ch, r0, c0, size, margin))
    print("cfg%d full size: worst block rel L2 vs oracle %.3e" % (cfg, worst))
    assert worst <= 1e-12
