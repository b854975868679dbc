"""CPU tests of the steps either side of the hot path: the oracle (oracle/frontend_oracle.py)
against the reference's golden values, Random123's known-answer vectors and committed cv2 fixtures
(tests/golden/make_frontend_golden.py), and the host-only C-ABI entry points (ENVI header / writer, SpectralPCA
training) against the oracle.  Nothing here needs a GPU."""
import importlib
import math
import os

import numpy as np
import pytest

from oracle import frontend_oracle as fo

srb = importlib.import_module("super-resolution_b200")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frontend_fixtures.npz")
REFERENCE = "/root/reference"   # present in the build container only


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


# ---- N3: bilinear resize ------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,tol", [("s2", 4e-16), ("s4", 4e-16), ("s4_big", 4e-16), ("s3", 5e-7)])
def test_resize_linear_oracle_against_cv2(gold, name, tol):
    """cv2.resize(INTER_LINEAR) on doubles.  Power-of-two factors: every interpolation weight is exact in float and
    in double and the oracle agrees with cv2 to 2 ulp (cv2's own SIMD build rounds a*(1-w) + b*w differently in
    places).  Factor 3: OpenCV evaluates part of its coefficient tables in float (~6e-8 relative weight error), so
    cv2 itself is only reproducible to ~3e-7; the oracle uses the exact weights."""
    src, dst = gold["resize_%s_src" % name], gold["resize_%s_dst" % name]
    out = fo.resize_linear(src, dst.shape[1], dst.shape[2])
    assert out.shape == dst.shape
    assert np.abs(out - dst).max() <= tol


def test_resize_linear_geometry():
    """Half-pixel centres with clamping: constant images stay constant, the first / last s/2 outputs replicate the
    border sample, and a linear ramp is reproduced in the interior (image_data.cpp:353-364: new size = int(w * s))."""
    x = np.full((1, 5, 7), 0.375)
    assert np.array_equal(fo.resize_linear(x, 20, 28), np.full((1, 20, 28), 0.375))
    ramp = np.arange(8, dtype=np.float64)[None, None, :] * np.ones((1, 3, 1))
    up = fo.resize_linear(ramp, 12, 32)
    assert np.array_equal(up[0, :, :2], np.zeros((12, 2))) and np.array_equal(up[0, :, -2:], np.full((12, 2), 7.0))
    np.testing.assert_allclose(up[0, 0, 2:-2], (np.arange(2, 30) + 0.5) / 4 - 0.5, rtol=0, atol=1e-15)


# ---- N3: scores (test/test_evaluation.cpp) ---------------------------------------------------------------------
TRUTH = np.array([[0.0, 0.1, 0.2, 0.3], [0.7, 0.6, 0.5, 0.4], [0.8, 0.9, 1.0, 0.5], [0.4, 0.6, 0.0, 1.0]])
IMAGE_3 = np.array([[0.2, 0.9, 1.0, 0.0], [0.7, 0.0, 0.8, 0.3], [0.1, 0.0, 0.2, 1.0], [0.0, 0.5, 0.5, 0.3]])


def test_psnr_reference_golden_values():
    """test/test_evaluation.cpp:12-97 (EXPECT_DOUBLE_EQ = 4 ulp)."""
    assert fo.psnr(TRUTH, TRUTH) == math.inf
    img2 = TRUTH.copy()
    img2.flat[6], img2.flat[15] = 0.25, 0.5
    assert abs(fo.psnr(img2, TRUTH) - 17.09269960975831) <= 4 * np.spacing(17.09269960975831)
    ssd = float(np.sum((TRUTH - IMAGE_3) ** 2))
    assert abs(fo.psnr(IMAGE_3, TRUTH) - 10.0 * math.log10(1.0 / (ssd / 16.0))) <= 1e-14
    multi_t = np.stack([TRUTH] * 3)
    multi_i = np.stack([TRUTH, img2, IMAGE_3])
    assert abs(fo.psnr(multi_i, multi_t) - 10.0 * math.log10(1.0 / ((0.3125 + ssd) / 48.0))) <= 1e-14


def test_ssim_reference_golden_values():
    """test/test_evaluation.cpp:99-140."""
    t = np.array([[0.5, 0.25], [0.75, 1.0]])
    i = np.array([[0.55, 0.25], [0.7, 1.0]])
    assert abs(fo.ssim(i, t) - 0.991784423266513) <= 4 * np.spacing(1.0)
    assert abs(fo.ssim(np.stack([i, i]), np.stack([t, t])) - 0.991784423266513) <= 4 * np.spacing(1.0)
    assert abs(fo.ssim(t, t) - 1.0) <= 1e-15


# ---- N2: noise generator ------------------------------------------------------------------------------------------
def test_philox4x32_10_known_answers():
    """Random123 kat_vectors (philox4x32, 10 rounds)."""
    assert fo.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert fo.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert fo.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # the vectorised form used for whole images == the scalar one
    w = fo.philox_words(5, seed=0x299f31d0a4093822, stream=7)
    for g in range(5):
        assert [int(v) for v in w[g]] == fo.philox4x32_10([g, 0, 7, 0], [0xa4093822, 0x299f31d0])


def test_noise_statistics_and_scale():
    """additive_noise_module.cpp:26-27: sigma is on the 0..255 scale; samples are N(0, 1), independent."""
    z = fo.standard_normals(400000, seed=11)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    assert abs(np.mean(z ** 3)) < 2e-2 and abs(np.mean(z ** 4) - 3.0) < 5e-2
    assert abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 5e-3
    assert not np.array_equal(z[:100], fo.standard_normals(100, seed=12))
    assert not np.array_equal(z[:100], fo.standard_normals(100, seed=11, stream=1))
    x = np.zeros((2, 50, 60))
    y = fo.add_noise(x, 5.0, seed=3)
    assert abs(y.std() - 5.0 / 255.0) < 5e-4


# ---- N4: ENVI ----------------------------------------------------------------------------------------------------
def _example_cube():
    """test_data/example_envi_data's pattern (test/test_hyperspectral_data_loader.cpp:59-66): 10 bands of 9 x 5,
    value = band + row / 10 + col / 100."""
    b, r, c = np.meshgrid(np.arange(10), np.arange(9), np.arange(5), indexing="ij")
    return (b + 0.1 * r + 0.01 * c).astype(np.float32)


def test_envi_reader_reference_golden(tmp_path):
    """test/test_hyperspectral_data_loader.cpp:52-86 with test_data/test_hs_config.txt's range."""
    cube = _example_cube()
    path = tmp_path / "example_envi_data"
    cube.astype("<f4").tofile(path)
    if os.path.exists(os.path.join(REFERENCE, "test_data", "example_envi_data")):
        ref_bytes = open(os.path.join(REFERENCE, "test_data", "example_envi_data"), "rb").read()
        mine = np.frombuffer(ref_bytes, dtype="<f4").reshape(10, 9, 5)
        assert np.abs(mine - cube).max() <= 1e-6      # the reference's own test tolerance (float text -> float)
        cube = mine.copy()
        cube.tofile(path)
    img = fo.envi_read(str(path), 9, 5, 10, False, (2, 8), (0, 3), (5, 10))
    assert img.shape == (5, 6, 3)
    exp0 = np.array([[5.20, 5.21, 5.22], [5.30, 5.31, 5.32], [5.40, 5.41, 5.42], [5.50, 5.51, 5.52], [5.60, 5.61, 5.62],
                     [5.70, 5.71, 5.72]])
    assert np.abs(img[0] - exp0).max() <= 1e-6 and np.abs(img[4] - (exp0 + 4.0)).max() <= 1e-6
    # big-endian file, same values
    bpath = tmp_path / "big"
    cube.astype(">f4").tofile(bpath)
    assert np.array_equal(fo.envi_read(str(bpath), 9, 5, 10, True, (2, 8), (0, 3), (5, 10)), img)


HEADER_TEXT = """ENVI
description = {
  Example ENVI header file}
samples = 11620
lines   = 11620
bands   = 1506
header offset = 0
file type = ENVI Standard
data type = 4
interleave = bsq
sensor type = Unknown
byte order = 0
wavelength units = Unknown
band names = {
 Band 1, Band 2}
"""


def test_envi_header_parser(tmp_path):
    """test/test_hyperspectral_data_loader.cpp:35-50; the C-ABI parser (host only) == the oracle."""
    p = tmp_path / "h.hdr"
    p.write_text(HEADER_TEXT)
    paths = [str(p)]
    ref_hdr = os.path.join(REFERENCE, "test_data", "example_envi_header.hdr")
    if os.path.exists(ref_hdr):
        paths.append(ref_hdr)
    for path in paths:
        ho = fo.envi_read_header(path)
        assert (ho["interleave_bsq"], ho["data_type"], ho["big_endian"], ho["header_offset"]) == (1, 4, 0, 0)
        assert (ho["num_data_rows"], ho["num_data_cols"], ho["num_data_bands"]) == (11620, 11620, 1506)
        h = srb.envi_read_header(path)
        for k, v in ho.items():
            assert getattr(h, k) == v, k
    p.write_text(HEADER_TEXT.replace("byte order = 0", "byte order = 1").replace("bsq", "bil").replace("= 4", "= 12"))
    h = srb.envi_read_header(str(p))
    assert (h.big_endian, h.interleave_bsq, h.data_type) == (1, 0, 12)
    with pytest.raises(srb.SrbError):
        srb.envi_read_header(str(tmp_path / "missing.hdr"))


def test_envi_writer_round_trip(tmp_path):
    """WriteBinaryFileBSQ (hyperspectral_data_loader.cpp:120-196), test :88-118: what is written reads back within the
    float conversion; the .hdr / .config carry the reference's keys (samples = rows, lines = columns as it writes them)."""
    img = np.random.default_rng(5).random((4, 6, 7))
    path = str(tmp_path / "out_envi")
    srb.envi_write(path, img)
    h = srb.envi_read_header(path + ".hdr")
    assert (h.num_data_rows, h.num_data_cols, h.num_data_bands, h.big_endian, h.header_offset) == (6, 7, 4, 0, 0)
    back = fo.envi_read(path, 6, 7, 4, False, (0, 6), (0, 7), (0, 4))
    assert np.array_equal(back, img.astype(np.float32).astype(np.float64))
    cfg = dict(line.split(None, 1) for line in open(path + ".config").read().splitlines() if line and not line.startswith("#"))
    assert cfg["file"] == path and cfg["interleave"] == "bsq" and cfg["data_type"] == "float" and cfg["big_endian"] == "false"
    assert (cfg["num_data_rows"], cfg["num_data_cols"], cfg["num_data_bands"]) == ("6", "7", "4")
    assert (cfg["end_row"], cfg["end_col"], cfg["end_band"]) == ("6", "7", "4")


# ---- N4: SpectralPCA -------------------------------------------------------------------------------------------
def _aligned(evec, ref):
    sign = np.sign(np.sum(evec * ref, axis=1))
    return evec * sign[:, None]


def test_pca_oracle_and_library_against_cv2(gold):
    """cv2.PCACompute2 / PCAProject / PCABackProject on the committed data set; eigenvectors up to sign."""
    data = gold["pca_data"]                       # 120 samples x 12 bands
    image = np.ascontiguousarray(data.T)          # one [C][P] image whose sub-sampling picks every pixel
    assert np.array_equal(fo.pca_input_data([image]), data)
    mean, evec, ev = fo.pca_train([image], num_pca_bands=5)
    np.testing.assert_allclose(mean, gold["pca_mean"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(ev, gold["pca_eigenvalues"], rtol=1e-10)
    np.testing.assert_allclose(_aligned(evec, gold["pca_eigenvectors"]), gold["pca_eigenvectors"], rtol=0, atol=1e-8)
    pca = srb.SpectralPCA([image], num_pca_bands=5)          # host-only training in the library
    assert (pca.num_bands, pca.num_components) == (12, 5)
    np.testing.assert_allclose(pca.mean, gold["pca_mean"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(pca.eigenvalues, gold["pca_eigenvalues"], rtol=1e-10)
    np.testing.assert_allclose(_aligned(pca.eigenvectors, gold["pca_eigenvectors"]), gold["pca_eigenvectors"], rtol=0, atol=1e-8)
    # projection / back-projection of the oracle with cv2's own basis == cv2
    probe = gold["pca_probe"]
    proj = fo.pca_project(gold["pca_mean"], gold["pca_eigenvectors"], probe.T)
    np.testing.assert_allclose(proj.T, gold["pca_projected"], rtol=0, atol=1e-13)
    back = fo.pca_reconstruct(gold["pca_mean"], gold["pca_eigenvectors"], proj)
    np.testing.assert_allclose(back.T, gold["pca_backprojected"], rtol=0, atol=1e-13)
    # retained variance
    assert fo.pca_train([image], retained_variance=0.97)[1].shape[0] == int(gold["pca_rv_components"][0])
    assert srb.SpectralPCA([image], retained_variance=0.97).num_components == int(gold["pca_rv_components"][0])


def test_pca_subsampling_rule():
    """GetPCAInputData (spectral_pca.cpp:44-93): 10 * C samples in all, split over the images, every
    (P // per_image)-th pixel; test/test_spectral_pca.cpp:19-75 reconstructs a 3-band image exactly with 3 components."""
    rng = np.random.default_rng(8)
    imgs = [rng.random((4, 9, 11)) for _ in range(3)]
    d = fo.pca_input_data(imgs)
    assert d.shape == (39, 4)                          # 40 // 3 = 13 per image, stride 99 // 13 = 7
    assert np.array_equal(d[13:26, 2], imgs[1][2].reshape(-1)[np.arange(13) * 7])
    pca = srb.SpectralPCA(imgs, num_pca_bands=4)
    mean, evec, ev = fo.pca_train(imgs, num_pca_bands=4)
    np.testing.assert_allclose(pca.eigenvalues, ev, rtol=1e-10)
    x = imgs[0]
    np.testing.assert_allclose(fo.pca_reconstruct(pca.mean, pca.eigenvectors, fo.pca_project(pca.mean, pca.eigenvectors, x)),
                               x, rtol=0, atol=1e-12)
    with pytest.raises(srb.SrbError):
        srb.SpectralPCA(imgs, num_pca_bands=5)
