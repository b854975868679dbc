"""GPU tests of the steps either side of the hot path (csrc/srb_frontend.cuh) through the C
ABI, against the oracle (oracle/frontend_oracle.py, pinned on the CPU in tests/test_frontend_oracle.py), the
reference's golden values and the committed cv2 fixtures."""
import math
import os
from importlib import import_module

import numpy as np
import pytest

from oracle import frontend_oracle as fo

pytestmark = pytest.mark.gpu
wl = import_module("super-resolution_b200.workloads")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frontend_fixtures.npz")


@pytest.fixture(scope="module")
def srb():
    import srb200
    assert srb200.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return srb200


@pytest.fixture(scope="module")
def eng(srb):
    s = 4
    with srb.Engine((16, 2, 24, 40), s, wl.gaussian_psf(7, 1.5), wl.default_shifts(16, s)) as e:
        yield e


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


# ---- N3 ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,H,W", [((3, 24, 40), 96, 160), ((1, 9, 11), 27, 33), ((2, 17, 5), 34, 20),
                                       ((1, 8, 8), 8, 8), ((1, 16, 12), 8, 6), ((1, 1, 7), 4, 28)])
def test_resize_linear_bit_exact_vs_oracle(eng, shape, H, W):
    x = np.random.default_rng(sum(shape)).random(shape)
    np.testing.assert_array_equal(eng.resize_linear(x, H, W), fo.resize_linear(x, H, W))


@pytest.mark.parametrize("name,tol", [("s2", 4e-16), ("s4_big", 4e-16), ("s3", 5e-7)])
def test_resize_linear_against_cv2_fixture(eng, gold, name, tol):
    src, dst = gold["resize_%s_src" % name], gold["resize_%s_dst" % name]
    assert np.abs(eng.resize_linear(src, dst.shape[1], dst.shape[2]) - dst).max() <= tol


def test_initial_estimate_is_the_upsampled_first_frame(srb, eng):
    """super_resolution.cpp:368-373: low_res_images[0].ResizeImage(scale, INTERPOLATE_LINEAR)."""
    lr = np.random.default_rng(4).random((eng.N, eng.C, eng.h, eng.w))
    eng.set_observations(lr)
    np.testing.assert_array_equal(eng.initial_estimate(), fo.resize_linear(lr[0], eng.H, eng.W))
    np.testing.assert_array_equal(eng.initial_estimate(frame=5), fo.resize_linear(lr[5], eng.H, eng.W))
    eng.set_channel_range(1, 2)
    np.testing.assert_array_equal(eng.initial_estimate(), fo.resize_linear(lr[0, 1:2], eng.H, eng.W))
    eng.set_channel_range(0, eng.C)
    with pytest.raises(srb.SrbError):
        eng.initial_estimate(frame=eng.N)


def test_scores_reference_golden_values(eng):
    """test/test_evaluation.cpp:12-140 through srb_score."""
    truth = np.array([[0.0, 0.1, 0.2, 0.3], [0.7, 0.6, 0.5, 0.4], [0.8, 0.9, 1.0, 0.5], [0.4, 0.6, 0.0, 1.0]])
    assert eng.score(truth, truth)[0] == math.inf
    img2 = truth.copy()
    img2.flat[6], img2.flat[15] = 0.25, 0.5
    assert abs(eng.score(img2, truth)[0] - 17.09269960975831) <= 4 * np.spacing(17.09269960975831)
    t = np.array([[0.5, 0.25], [0.75, 1.0]])
    i = np.array([[0.55, 0.25], [0.7, 1.0]])
    assert abs(eng.score(i, t)[1] - 0.991784423266513) <= 4 * np.spacing(1.0)
    assert abs(eng.score(np.stack([i, i]), np.stack([t, t]))[1] - 0.991784423266513) <= 4 * np.spacing(1.0)


def test_scores_match_oracle_on_images(eng):
    rng = np.random.default_rng(6)
    truth = wl.box_smooth(rng.random((3, 200, 333)))
    img = truth + 0.03 * rng.standard_normal(truth.shape)
    p, s = eng.score(img, truth)
    assert abs(p - fo.psnr(img, truth)) <= 1e-12 * abs(p)
    assert abs(s - fo.ssim(img, truth)) <= 1e-12
    p2, s2 = eng.score(img, truth, k1=0.05, k2=0.1, image_scale=255.0)
    assert p2 == p and abs(s2 - fo.ssim(img, truth, 0.05, 0.1, 255.0)) <= 1e-12


# ---- N2 ----------------------------------------------------------------------------------------------------------
def test_noise_matches_the_oracle_generator(srb, eng):
    """Same Philox counters, same Box-Muller: device libm vs numpy differ by rounding only."""
    x = np.random.default_rng(1).random((3, 37, 41))          # size not a multiple of 4
    y = eng.add_noise(x, 5.0, seed=0x123456789ABCDEF, stream_id=3)
    np.testing.assert_allclose(y, fo.add_noise(x, 5.0, 0x123456789ABCDEF, 3), rtol=0, atol=1e-15)
    np.testing.assert_array_equal(y, eng.add_noise(x, 5.0, seed=0x123456789ABCDEF, stream_id=3))   # deterministic
    assert not np.array_equal(y, eng.add_noise(x, 5.0, seed=2, stream_id=3))
    with pytest.raises(srb.SrbError):
        eng.add_noise(x, 0.0)                                    # additive_noise_module.cpp:15-17: CHECK_GT(sigma, 0)


def test_noise_statistics(eng):
    z = (eng.add_noise(np.zeros(1 << 20), 255.0, seed=99))      # sigma / 255 = 1
    assert abs(z.mean()) < 4e-3 and abs(z.std() - 1.0) < 4e-3
    assert abs(np.mean(z ** 4) - 3.0) < 5e-2 and abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 4e-3


def test_generate_observations_is_the_image_model_plus_noise(srb, oracle, eng):
    """image_model.cpp:76-84 per frame (the oracle's forward model) with the noise module last; keep=True makes the
    stack the context's observations without a host round trip."""
    rng = np.random.default_rng(12)
    hr = wl.box_smooth(rng.random((eng.C, eng.H, eng.W)))
    m = oracle.Model(eng.scale, wl.gaussian_psf(7, 1.5), wl.default_shifts(16, 4))
    clean = np.stack([np.stack([oracle.forward(m, k, hr[c]) for c in range(eng.C)]) for k in range(eng.N)])
    lr0 = eng.generate_observations(hr, noise_sigma=0.0, keep=False)
    np.testing.assert_array_equal(lr0, clean)
    lr = eng.generate_observations(hr, noise_sigma=2.0, seed=77, keep=True)
    np.testing.assert_allclose(lr, fo.add_noise(clean, 2.0, 77), rtol=0, atol=1e-15)
    x = rng.random(hr.shape)
    f1, g1 = eng.eval(x)
    eng.set_observations(lr)
    f2, g2 = eng.eval(x)
    assert f1 == f2 and np.array_equal(g1, g2)
    import torch
    hr_dev = torch.from_numpy(hr).cuda()
    lr_d = eng.generate_observations(hr_dev, noise_sigma=2.0, seed=77, keep=False)
    np.testing.assert_array_equal(lr_d, lr)


# ---- N4 ----------------------------------------------------------------------------------------------------------
def test_envi_reader_reference_golden(srb, eng, tmp_path):
    """test/test_hyperspectral_data_loader.cpp:52-86: bands 5-10, rows 2-8, columns 0-3 of the 10 x 9 x 5 cube whose
    values are band + row / 10 + col / 100; little- and big-endian files."""
    b, r, c = np.meshgrid(np.arange(10), np.arange(9), np.arange(5), indexing="ij")
    cube = (b + 0.1 * r + 0.01 * c).astype(np.float32)
    h = srb.EnviHeader(1, 4, 0, 0, 9, 5, 10)
    for big in (0, 1):
        path = str(tmp_path / ("cube%d" % big))
        cube.astype(">f4" if big else "<f4").tofile(path)
        h.big_endian = big
        img = eng.envi_read(path, h, rows=(2, 8), cols=(0, 3), bands=(5, 10))
        assert img.shape == (5, 6, 3)
        np.testing.assert_array_equal(img, fo.envi_read(path, 9, 5, 10, bool(big), (2, 8), (0, 3), (5, 10)))
        assert abs(img[0, 0, 0] - 5.20) <= 1e-6 and abs(img[4, 5, 2] - 9.72) <= 1e-6
        np.testing.assert_array_equal(eng.envi_read(path, h), cube.astype(np.float64))
    with pytest.raises(srb.SrbError):
        eng.envi_read(path, h, rows=(2, 12))
    h.num_data_bands = 11
    with pytest.raises(srb.SrbError):
        eng.envi_read(path, h)                                    # file shorter than the header says


def test_envi_write_then_read_on_device(srb, eng, tmp_path):
    img = np.random.default_rng(5).random((6, 33, 47))
    path = str(tmp_path / "w")
    srb.envi_write(path, img)
    h = srb.envi_read_header(path + ".hdr")
    np.testing.assert_array_equal(eng.envi_read(path, h), img.astype(np.float32).astype(np.float64))


def test_pca_projection_against_oracle_and_cv2(srb, eng, gold):
    """SpectralPCA::GetPCAImage / ReconstructImage (spectral_pca.cpp:98-153)."""
    data = gold["pca_data"]
    image = np.ascontiguousarray(data.T)
    pca = srb.SpectralPCA([image], num_pca_bands=5)
    probe = np.ascontiguousarray(gold["pca_probe"].T)            # [12][7]
    proj = eng.pca_project(pca, probe)
    np.testing.assert_allclose(proj, fo.pca_project(pca.mean, pca.eigenvectors, probe), rtol=0, atol=1e-13)
    sign = np.sign(np.sum(pca.eigenvectors * gold["pca_eigenvectors"], axis=1))
    np.testing.assert_allclose(proj * sign[:, None], gold["pca_projected"].T, rtol=0, atol=1e-8)
    back = eng.pca_reconstruct(pca, proj)
    np.testing.assert_allclose(back, gold["pca_backprojected"].T, rtol=0, atol=1e-8)       # sign-free
    # a hyperspectral-sized case: 40 bands -> 36 components (two output passes of the kernel), image-shaped
    rng = np.random.default_rng(3)
    cube = (rng.standard_normal((40, 50, 60)) * (1.5 ** -np.arange(40))[:, None, None]) + rng.random((40, 1, 1))
    p2 = srb.SpectralPCA([cube], num_pca_bands=36)
    pr = eng.pca_project(p2, cube)
    np.testing.assert_allclose(pr, fo.pca_project(p2.mean, p2.eigenvectors, cube), rtol=0, atol=1e-12)
    np.testing.assert_allclose(eng.pca_reconstruct(p2, pr), fo.pca_reconstruct(p2.mean, p2.eigenvectors, pr), rtol=0, atol=1e-12)
    p3 = srb.SpectralPCA([cube], num_pca_bands=40)              # test/test_spectral_pca.cpp: all components = lossless
    np.testing.assert_allclose(eng.pca_reconstruct(p3, eng.pca_project(p3, cube)), cube, rtol=0, atol=1e-10)


# ---- the whole chain on the device ---------------------------------------------------------------------------------
def test_device_pipeline_generate_estimate_solve_score(srb):
    """generate_data -> initial estimate -> IRLS + CG solve -> PSNR, observations and estimate never leaving the device
    except as the final image: the solve must beat the bilinear initial estimate (what test/test_map_solver.cpp's
    RegularizationTest asserts with PSNR orderings)."""
    s, K = 2, 3
    rng = np.random.default_rng(21)
    truth = wl.box_smooth(rng.random((1, 64, 96)), 5)
    with srb.Engine((4, 1, 32, 48), s, wl.gaussian_psf(K, 1.0), wl.default_shifts(4, s)) as e:
        e.generate_observations(truth, noise_sigma=1.0, seed=5, keep=True, want_lr=False)
        x0 = e.initial_estimate()
        e.set_regularizer(srb.REG_TV, 0.001)
        x, rep = e.solve_irls(x0, maxits=30, max_irls_iterations=3, irls_cost_difference_threshold=1e-9)
        p0, s0 = e.score(x0, truth)
        p1, s1 = e.score(x, truth)
    assert rep["num_irls_iterations"] >= 1 and p1 > p0 + 3.0 and s1 > s0
