"""Pins the CPU oracle (oracle/sr_oracle.c) against

  (1) every golden vector the reference's own tests hold for the hot path,
      transcribed below with the reference test file:line each comes from;
  (2) fixtures produced by the OpenCV entry points the reference calls (tests/golden/make_golden.py);
  (3) the reference's own regularizer / IRLS-term / solver sources compiled unmodified (oracle/_ref).
"""
import numpy as np
import pytest

# test/test_image_model.cpp:22-27
SMALL_TEST_IMAGE = np.array([[1, 2, 3, 4, 5, 6],
                             [7, 8, 9, 0, 1, 2],
                             [9, 7, 5, 4, 2, 1],
                             [2, 4, 6, 8, 0, 1]], dtype=np.float64)


# ---------------------------------------------------------------- (1) reference golden vectors
def test_nearest_and_additive_resize_golden(oracle):
    """test/test_image_data.cpp:311-401 (exact)."""
    img = np.array([[0.1, 0.2, 0.3, 0.4], [0.5, 0.6, 0.7, 0.8],
                    [0.9, 1.0, 0.0, 0.2], [0.4, 0.6, 0.8, 1.0]])
    np.testing.assert_array_equal(oracle.resize_nearest(img, 2, 2), [[0.1, 0.3], [0.9, 0.0]])
    h, w = oracle.lr_size(2, 4, 4)          # ResizeImage(0.5, NEAREST)
    assert (h, w) == (2, 2)
    big = oracle.resize_nearest(img, 8, 8)
    np.testing.assert_array_equal(big, np.repeat(np.repeat(img, 2, axis=0), 2, axis=1))
    up = oracle.resize_additive(img, 8, 8)
    exp = np.zeros((8, 8))
    exp[::2, ::2] = img
    np.testing.assert_array_equal(up, exp)
    down = oracle.resize_additive(img, 2, 2)
    exp_down = np.array([[0.1 + 0.2 + 0.5 + 0.6, 0.3 + 0.4 + 0.7 + 0.8],
                         [0.9 + 1.0 + 0.4 + 0.6, 0.0 + 0.2 + 0.8 + 1.0]])
    np.testing.assert_array_equal(down, exp_down)


def test_downsampling_module_golden(oracle):
    """test/test_image_model.cpp:87-226: nearest decimation vector and zero-insert transpose."""
    m = oracle.Model(2, None, None, num_frames=1)
    np.testing.assert_array_equal(oracle.forward(m, 0, SMALL_TEST_IMAGE), [[1, 3, 5], [9, 5, 2]])
    exp = np.zeros((8, 12))
    exp[::2, ::2] = SMALL_TEST_IMAGE
    np.testing.assert_array_equal(oracle.transpose(m, 0, SMALL_TEST_IMAGE), exp)


def test_motion_module_golden(oracle):
    """test/test_image_model.cpp:229-348: operator matrices for shifts (0,0), (1,1), (-1,0) on 3x3.
    The matrices say out(r,c) = in(r-dy, c-dx); warpAffine must agree for integer shifts."""
    img = np.arange(1, 10, dtype=np.float64).reshape(3, 3)
    for dx, dy in [(0, 0), (1, 1), (-1, 0)]:
        exp = np.zeros((3, 3))
        for r in range(3):
            for c in range(3):
                sr, sc = r - dy, c - dx
                if 0 <= sr < 3 and 0 <= sc < 3:
                    exp[r, c] = img[sr, sc]
        np.testing.assert_array_equal(oracle.warp_shift(img, dx, dy), exp)


def test_blur_module_golden(oracle):
    """test/test_image_model.cpp:350-408: 3x3 sigma=0.849321 blur of the 4x6 image, tol 1e-3;
    transpose equals forward (symmetric kernel)."""
    exp = np.array([[1.875, 3.0, 3.125, 2.625, 2.75, 2.4375],
                    [4.5625, 6.25, 5.3125, 3.1875, 2.3125, 1.9375],
                    [5.0, 6.5, 5.75, 3.875, 1.9375, 0.9375],
                    [2.5625, 3.75, 4.3125, 3.6875, 1.6875, 0.5]])
    psf = oracle.gaussian_psf(3, 0.849321)
    assert np.abs(oracle.filter2d(SMALL_TEST_IMAGE, psf) - exp).max() <= 1e-3
    assert np.abs(oracle.filter2d(SMALL_TEST_IMAGE, psf.T.copy()) - exp).max() <= 1e-3


def test_kernel_operator_matrix_golden(oracle):
    """test/test_image_model.cpp:49-78: correlation with zero border."""
    kernel = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=np.float64)
    img = np.array([[1, 3, 5], [9, 5, 2]], dtype=np.float64)
    np.testing.assert_array_equal(oracle.filter2d(img, kernel), [[11, 1, -11], [13, -10, -13]])


TV_IMAGE = np.array([[0, 0, 1], [0, 1, 3], [-3, -1, 0]], dtype=np.float64)
TV_EXPECTED = np.array([[0, 2, 2], [4, 4, 3], [2, 1, 0]], dtype=np.float64)


def test_tv_values_golden(oracle):
    """test/test_tv_regularizer.cpp:62-73 (2-D, 3 channels, exact)."""
    x = np.stack([TV_IMAGE] * 3)
    np.testing.assert_array_equal(oracle.reg_apply(oracle.REG_TV, x), np.stack([TV_EXPECTED] * 3))


def test_tv3d_values_golden(oracle):
    """test/test_tv_regularizer.cpp:76-145 (exact)."""
    x = np.array([TV_IMAGE, np.zeros((3, 3)), [[0, -1, 2], [-3, 4, 5], [6, 7, -8]]])
    exp = np.array([[[0, 2, 3], [4, 5, 6], [5, 2, 0]],
                    [[0, 1, 2], [3, 4, 5], [6, 7, 8]],
                    [[4, 8, 3], [16, 4, 13], [1, 15, 0]]], dtype=np.float64)
    np.testing.assert_array_equal(oracle.reg_apply(oracle.REG_TV3D, x), exp)


def test_tv_gradient_vs_finite_differences(oracle):
    """test/test_tv_regularizer.cpp:150-198: analytic gradient vs central differences, tol 1e-4."""
    x = TV_IMAGE[None]
    values, partials = oracle.reg_apply_diff(oracle.REG_TV, x, np.ones_like(x))
    np.testing.assert_array_equal(values[0], TV_EXPECTED)
    d = 1e-6
    for i in range(9):
        xp, xm = x.copy().ravel(), x.copy().ravel()
        xp[i] += d
        xm[i] -= d
        fp = np.sum(oracle.reg_apply(oracle.REG_TV, xp.reshape(1, 3, 3)) ** 2)
        fm = np.sum(oracle.reg_apply(oracle.REG_TV, xm.reshape(1, 3, 3)) ** 2)
        assert abs((fp - fm) / (2 * d) - partials.ravel()[i]) <= 1e-4


BTV_IMAGE = np.array([[0, 0, 1, 2, 1], [0, 1, 3, 2, 3], [5, 4, 3, -2, 1],
                      [4, 6, 9, 3, 0], [-3, -1, 0, 6, 0]], dtype=np.float64)


def test_btv_values_golden(oracle):
    """test/test_btv_regularizer.cpp:21-95 (EXPECT_DOUBLE_EQ)."""
    v = oracle.reg_apply(oracle.REG_BTV, BTV_IMAGE[None], btv_range=2, btv_decay=0.5)
    assert v[0, 0, 0] == 2.8125 and v[0, 4, 4] == 0.0
    v2 = oracle.reg_apply(oracle.REG_BTV, np.stack([BTV_IMAGE] * 2), btv_range=1, btv_decay=0.25)
    assert v2.ravel()[7] == 0.5625 and v2.ravel()[25 + 7] == 0.5625
    assert v2.ravel()[24] == 0.0 and v2.ravel()[49] == 0.0
    vd, _ = oracle.reg_apply_diff(oracle.REG_BTV, BTV_IMAGE[None], np.full((1, 5, 5), 0.5),
                                  btv_range=2, btv_decay=0.5)
    assert vd[0, 0, 0] == 2.8125 and vd[0, 4, 4] == 0.0


# ---------------------------------------------------------------- (2) OpenCV fixtures
def test_gaussian_kernel_vs_opencv(oracle, cv2_fixtures):
    g = cv2_fixtures
    for i, (n, sigma) in enumerate(g["gauss_params"]):
        np.testing.assert_allclose(oracle.gaussian_kernel(int(n), sigma), g[f"gauss_{i}"],
                                   rtol=0, atol=2.3e-16)
        np.testing.assert_allclose(oracle.gaussian_psf(int(n), sigma), g[f"gauss_psf_{i}"],
                                   rtol=0, atol=2.3e-16)


def test_warp_quantisation_vs_opencv(oracle, cv2_fixtures):
    """The 1/32-px fixed-point shift rule (bit-exact integer map)."""
    g = cv2_fixtures
    n = np.array([oracle.warp_quantize(d) for d in g["warp_sweep_d"]])
    np.testing.assert_array_equal(n, g["warp_sweep_n"])
    ny = np.array([oracle.warp_quantize(d) for d in g["warp_sweep_dy"]])
    np.testing.assert_array_equal(ny, g["warp_sweep_ny"])


def test_warp_values_vs_opencv(oracle, cv2_fixtures):
    g = cv2_fixtures
    for i, (dx, dy) in enumerate(g["warp_shifts"]):
        np.testing.assert_array_equal(oracle.warp_shift(g["warp_img"], dx, dy), g[f"warp_{i}"])


def test_filter2d_vs_opencv(oracle, cv2_fixtures):
    g = cv2_fixtures
    np.testing.assert_array_equal(oracle.filter2d(g["filt_img"], g["filt_asym_kernel"]),
                                  g["filt_asym"])
    np.testing.assert_array_equal(
        oracle.filter2d(g["filt_img"], g["filt_asym_kernel"].T.copy()), g["filt_asym_t"])
    for i, (n, sigma) in enumerate([(3, 0.849321), (5, 1.5), (7, 2.0), (9, 2.5)]):
        np.testing.assert_allclose(oracle.filter2d(g["filt_img"], oracle.gaussian_psf(n, sigma)),
                                   g[f"filt_gauss_{i}"], rtol=0, atol=1e-15)


def test_nearest_index_map_vs_opencv_bit_exact(oracle, cv2_fixtures):
    g = cv2_fixtures
    maps = g["nn_maps"]
    for n, n2, off in g["nn_pairs"]:
        m = [oracle.nearest_index(q, int(n), int(n2)) for q in range(int(n2))]
        np.testing.assert_array_equal(m, maps[off:off + n2])


def _case(g, oracle, name):
    C, H, W, s, K, N = (int(v) for v in g[f"{name}_meta"])
    psf = g[f"{name}_psf"] if K > 0 else None
    sh = g[f"{name}_shifts"] if len(g[f"{name}_shifts"]) else None
    return oracle.Model(s, psf, sh, num_frames=N), C, N


CASES = ["int_s2", "int_s4", "int_s3_neg", "frac_s2", "frac_s4", "noblur", "nomotion", "k9_s4"]


@pytest.mark.parametrize("name", CASES)
def test_forward_transpose_data_term_vs_opencv(oracle, cv2_fixtures, name):
    """The whole data term (objective_data_term.cpp:15-116) restated with cv2 calls vs the oracle.
    K <= 7 is bit-exact (direct filter2D path); K = 9 goes through OpenCV's DFT path (~1e-15)."""
    g = cv2_fixtures
    m, C, N = _case(g, oracle, name)
    x, lr = g[f"{name}_x"], g[f"{name}_lr"]
    tol = dict(rtol=0, atol=0) if name != "k9_s4" else dict(rtol=0, atol=2e-15)
    for k in range(N):
        for c in range(C):
            np.testing.assert_allclose(oracle.forward(m, k, x[c]), g[f"{name}_forward"][k, c], **tol)
            np.testing.assert_allclose(oracle.transpose(m, k, lr[k, c]),
                                       g[f"{name}_transpose"][k, c], **tol)
    obs = oracle.upsample_observations(m, lr)
    cost, grad = oracle.data_term(m, x, obs)
    np.testing.assert_allclose(cost, float(g[f"{name}_cost"]), rtol=1e-14)
    np.testing.assert_allclose(grad, g[f"{name}_grad"], rtol=0,
                               atol=0 if name != "k9_s4" else 1e-13)
    cost_only, none = oracle.data_term(m, x, obs, want_grad=False)
    assert none is None and cost_only == cost
    cost_t, grad_t = oracle.data_term(m, x, obs, threads=4)
    assert abs(cost_t - cost) <= 1e-14 * abs(cost)   # per-(frame, channel) sums re-associated
    np.testing.assert_array_equal(grad_t, grad)


@pytest.mark.parametrize("name", ["int_s2", "int_s3_neg", "frac_s2", "nomotion"])
def test_data_term_matches_dense_operator_matrices(oracle, cv2_fixtures, name):
    """cost = s^2 sum ||A_k x - y_k||^2, grad = 2 s^2 sum A_k^T (A_k x - y_k) with A_k read off
    the oracle's forward / transpose applied to unit vectors (checks forward/transpose adjointness
    the way degradation_operator.cpp:21-81 / GetModelMatrix would)."""
    g = cv2_fixtures
    m, C, N = _case(g, oracle, name)
    x, lr = g[f"{name}_x"], g[f"{name}_lr"]
    _, H, W = x.shape
    s = m.scale
    h, w = oracle.lr_size(s, H, W)
    cost_ref, grad_ref = 0.0, np.zeros_like(x)
    for k in range(N):
        A = np.zeros((h * w, H * W))
        At = np.zeros((H * W, h * w))
        for p in range(H * W):
            e = np.zeros(H * W)
            e[p] = 1
            A[:, p] = oracle.forward(m, k, e.reshape(H, W)).ravel()
        for q in range(h * w):
            e = np.zeros(h * w)
            e[q] = 1
            At[:, q] = oracle.transpose(m, k, e.reshape(h, w)).ravel()
        for c in range(C):
            r = A @ x[c].ravel() - lr[k, c].ravel()
            cost_ref += s * s * float(r @ r)
            grad_ref[c] += (2 * s * s * (At @ r)).reshape(H, W)
    obs = oracle.upsample_observations(m, lr)
    cost, grad = oracle.data_term(m, x, obs)
    np.testing.assert_allclose(cost, cost_ref, rtol=1e-12)
    np.testing.assert_allclose(grad, grad_ref, rtol=0, atol=1e-12)


# ---------------------------------------------------------------- (3) reference sources (oracle/_ref)
@pytest.mark.parametrize("kind,R,decay", [(0, 3, 0.5), (1, 3, 0.5), (2, 1, 0.25), (2, 2, 0.5),
                                          (2, 3, 0.5), (2, 3, 1.0), (2, 4, 0.7)])
def test_regularizers_bit_exact_vs_reference_sources(oracle, ref, kind, R, decay):
    """tv_regularizer.cpp / btv_regularizer.cpp compiled unmodified vs the restatement, on an image
    with exact ties (sign(0) = 0 paths) -- values and gradients bit-for-bit."""
    rng = np.random.default_rng(7)
    x = rng.random((3, 9, 11))
    x[0, 2, 3] = x[0, 2, 4]
    x[1, 1, 1] = x[1, 2, 1]
    x[2, 5, 5] = x[1, 5, 5]
    c = rng.random(x.shape)
    np.testing.assert_array_equal(oracle.reg_apply(kind, x, R, decay), ref.reg_apply(kind, x, R, decay))
    v, p = oracle.reg_apply_diff(kind, x, c, R, decay)
    v2, p2 = ref.reg_apply_diff(kind, x, c, R, decay)
    np.testing.assert_array_equal(v, v2)
    np.testing.assert_array_equal(p, p2)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_irls_term_and_compute_all_terms_vs_reference_sources(oracle, ref, cv2_fixtures, kind):
    """objective_irls_regularization_term.cpp and objective_function.cpp compiled unmodified."""
    g = cv2_fixtures
    m, C, N = _case(g, oracle, "int_s3_neg")
    x, lr = g["int_s3_neg_x"], g["int_s3_neg_lr"]
    rng = np.random.default_rng(3)
    w = 0.5 + rng.random(x.shape)
    g1, g2 = np.full_like(x, 0.25), np.full_like(x, 0.25)
    f1 = oracle.irls_term(kind, 0.01, w, x, g1)
    f2 = ref.irls_term(kind, 0.01, w, x, g2)
    assert f1 == f2
    np.testing.assert_array_equal(g1, g2)
    assert oracle.irls_term(kind, 0.0, w, x, g1) == 0.0       # lambda <= 0 early-out (:15-18)
    obs = oracle.upsample_observations(m, lr)
    fa, ga = oracle.evaluate(m, x, obs, kind, 0.01, w)
    fb, gb = ref.compute_all_terms(m, x, obs, kind, 0.01, w)
    assert fa == fb
    np.testing.assert_array_equal(ga, gb)
    fc, none = oracle.evaluate(m, x, obs, kind, 0.01, w, want_grad=False)
    assert none is None and fc == fa


def test_reweight(oracle):
    """irls_map_solver.cpp:128-143: w = 1 / max(1e-5, r)."""
    x = np.stack([TV_IMAGE])
    w = oracle.reweight(oracle.REG_TV, x)
    exp = 1.0 / np.maximum(1e-5, TV_EXPECTED)
    np.testing.assert_array_equal(w[0], exp)
    assert w[0, 0, 0] == 1.0 / 0.00001


def test_reference_solver_small_data_test(oracle, ref):
    """test/test_map_solver.cpp:79-199 (SmallDataTest) through the reference's own IRLSMapSolver +
    ALGLIB with the oracle data term: recovers the 4x4 ground truth within 1e-3, also with 10
    channels and with split_channels."""
    lr1 = np.array([np.full((2, 2), v) for v in (0.4, 0.2, 0.0, 1.0)])[:, None]
    truth = np.array([[0.4, 0.2, 0.4, 0.2], [0.0, 1.0, 0.0, 1.0]] * 2)
    m = oracle.Model(2, None, [(0, 0), (-1, 0), (0, -1), (-1, -1)])
    res, _ = ref.solve(m, lr1, np.zeros((1, 4, 4)))
    assert np.abs(res[0] - truth).max() <= 1e-3
    lr10 = np.repeat(lr1, 10, axis=1)
    res10, _ = ref.solve(m, lr10, np.zeros((10, 4, 4)))
    assert np.abs(res10 - truth[None]).max() <= 1e-3
    opt = ref.default_options()
    opt.split_channels = 1
    res10s, _ = ref.solve(m, lr10, np.zeros((10, 4, 4)), options=opt)
    assert np.abs(res10s - truth[None]).max() <= 1e-3


def test_reference_solver_real_icon_data_test(oracle, ref, cv2_fixtures):
    """test/test_map_solver.cpp:205-308 (RealIconDataTest): fb.png gray, 2x, four integer shifts, no
    blur; solver == ground truth == closed form (sum A^T A)^-1 sum A^T y on the interior 26x26."""
    g = cv2_fixtures
    truth = g["fb_gray_u8"].astype(np.float64) / 255.0
    m = oracle.Model(2, None, g["icon_shifts"])
    lr = g["icon_lr"]
    for k in range(4):  # the fixture LR frames (cv2) equal the oracle forward model
        np.testing.assert_array_equal(oracle.forward(m, k, truth), lr[k, 0])
    res, st = ref.solve(m, lr, g["icon_x0"])
    assert np.abs(res[0] - truth)[1:27, 1:27].max() <= 1e-3
    assert st.num_data_term_evals > 0
    # closed form through dense operator matrices
    Z = np.zeros((784, 784))
    b = np.zeros(784)
    for k in range(4):
        A = np.zeros((196, 784))
        for p in range(784):
            e = np.zeros(784)
            e[p] = 1
            A[:, p] = oracle.forward(m, k, e.reshape(28, 28)).ravel()
        Z += A.T @ A
        b += A.T @ lr[k, 0].ravel()
    closed = (np.linalg.pinv(Z) @ b).reshape(28, 28)
    assert np.abs(closed - truth)[1:27, 1:27].max() <= 1e-3
