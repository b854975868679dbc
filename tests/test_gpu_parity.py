"""GPU parity tests (run with `-m gpu` on the B200 box): the CUDA path, called through the C ABI
(libsrb200.so via super-resolution_b200/engine.py), against

  * the committed golden fixtures (tests/golden/cv2_fixtures.npz: outputs of the OpenCV entry
    points the reference calls, and the reference's own golden vectors),
  * the CPU oracle (oracle/sr_oracle.c) on the same seeded inputs,
  * the reference's own regularizer / objective sources (oracle/_ref, prebuilt).

Bars: bit-exact for index maps, for the regularizers and for the reference-order kernels'
LR prediction and gradient; the fused kernel reassociates the sums (blur commuted with the
shifts) and is held to 1e-12 relative L2 / 1e-11 max-abs relative to max|g|; costs (tree
reductions) to 1e-12 relative.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = ["int_s2", "int_s4", "int_s3_neg", "frac_s2", "frac_s4", "noblur", "nomotion", "k9_s4"]
SMALL_TEST_IMAGE = np.array([[1, 2, 3, 4, 5, 6], [7, 8, 9, 0, 1, 2], [9, 7, 5, 4, 2, 1],
                             [2, 4, 6, 8, 0, 1]], dtype=np.float64)
FUSED_REL_L2 = 1e-12
COST_RTOL = 1e-12


@pytest.fixture(scope="module")
def srb():
    import srb200
    assert srb200.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return srb200


def _case(g, name):
    C, H, W, s, K, N = (int(v) for v in g[f"{name}_meta"])
    psf = g[f"{name}_psf"] if K > 0 else None
    sh = g[f"{name}_shifts"] if len(g[f"{name}_shifts"]) else None
    return C, H, W, s, psf, sh, N


def _engine(srb, g, name):
    C, H, W, s, psf, sh, N = _case(g, name)
    lr = g[f"{name}_lr"]
    e = srb.Engine(lr.shape, s, psf, sh)
    e.set_observations(lr)
    return e


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


# ------------------------------------------------------------------ reference golden vectors
def test_downsampling_and_motion_golden(srb):
    """test/test_image_model.cpp:87-226 (nearest decimation, zero-insert transpose) and :229-348
    (shift semantics out(r,c) = in(r-dy, c-dx)), exact."""
    with srb.Engine((1, 1, 2, 3), 2) as e:
        np.testing.assert_array_equal(e.forward(0, SMALL_TEST_IMAGE), [[1, 3, 5], [9, 5, 2]])
        exp = np.zeros((8, 12))
        exp[::2, ::2] = SMALL_TEST_IMAGE
        np.testing.assert_array_equal(e.transpose(0, SMALL_TEST_IMAGE), exp)
    img = np.arange(1, 10, dtype=np.float64).reshape(3, 3)
    shifts = [(0, 0), (1, 1), (-1, 0)]
    with srb.Engine((3, 1, 3, 3), 1, None, shifts) as e:
        for k, (dx, dy) in enumerate(shifts):
            exp = np.zeros((3, 3))
            for r in range(3):
                for c in range(3):
                    if 0 <= r - dy < 3 and 0 <= c - dx < 3:
                        exp[r, c] = img[r - dy, c - dx]
            np.testing.assert_array_equal(e.forward(k, img), exp)


def test_blur_module_golden(srb):
    """test/test_image_model.cpp:350-408: 3x3 sigma=0.849321 blur of the 4x6 image (tol 1e-3),
    forward == transpose for the symmetric kernel."""
    from importlib import import_module
    wl = import_module("super-resolution_b200.workloads")
    exp = np.array([[1.875, 3.0, 3.125, 2.625, 2.75, 2.4375],
                    [4.5625, 6.25, 5.3125, 3.1875, 2.3125, 1.9375],
                    [5.0, 6.5, 5.75, 3.875, 1.9375, 0.9375],
                    [2.5625, 3.75, 4.3125, 3.6875, 1.6875, 0.5]])
    with srb.Engine((1, 1, 4, 6), 1, wl.gaussian_psf(3, 0.849321)) as e:
        assert np.abs(e.forward(0, SMALL_TEST_IMAGE) - exp).max() <= 1e-3
        assert np.abs(e.transpose(0, SMALL_TEST_IMAGE) - exp).max() <= 1e-3
    kernel = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=np.float64)  # :49-78
    with srb.Engine((1, 1, 2, 3), 1, kernel) as e:
        np.testing.assert_array_equal(e.forward(0, np.array([[1, 3, 5], [9, 5, 2.0]])),
                                      [[11, 1, -11], [13, -10, -13]])


def test_decimation_index_map_bit_exact(srb, oracle, cv2_fixtures):
    """The cv::resize INTER_NEAREST index map, for sizes that are NOT multiples of the scale too:
    srb_forward of an index ramp returns the source index itself."""
    g = cv2_fixtures
    maps = g["nn_maps"]
    engines = {s: srb.Engine((1, 1, 1, 1), s) for s in (2, 3, 4, 5, 7)}
    checked, seen = 0, set()
    for n, n2, off in g["nn_pairs"]:
        n, n2, off = int(n), int(n2), int(off)
        if n2 > n or n2 <= 0 or (n, n2) in seen:
            continue
        for s, e in engines.items():
            if int(n * (1.0 / s)) != n2:
                continue
            seen.add((n, n2))
            rows = s + 1
            ramp = np.tile(np.arange(n, dtype=np.float64), (rows, 1))   # value = column index
            got = e.forward(0, ramp)
            np.testing.assert_array_equal(got[0], maps[off:off + n2])      # cv2's own index map
            got = e.forward(0, ramp.T.copy())                            # value = row index
            np.testing.assert_array_equal(got[:, 0], maps[off:off + n2])
            checked += 1
            break
    assert checked >= 50
    for (W, s) in [(7, 2), (10, 3), (29, 4), (100, 7), (2047, 4), (4099, 5)]:
        w2 = int(W * (1.0 / s))
        ramp = np.tile(np.arange(W, dtype=np.float64), (s + 1, 1))
        got = engines[s].forward(0, ramp)
        exp = [oracle.nearest_index(q, W, w2) for q in range(w2)]
        np.testing.assert_array_equal(got[0], exp)
    for e in engines.values():
        e.close()


# ------------------------------------------------------------------ OpenCV fixtures + oracle
@pytest.mark.parametrize("name", CASES)
def test_forward_transpose_vs_opencv_fixtures(srb, cv2_fixtures, name):
    g = cv2_fixtures
    C, H, W, s, psf, sh, N = _case(g, name)
    x, lr = g[f"{name}_x"], g[f"{name}_lr"]
    tol = dict(rtol=0, atol=0) if name != "k9_s4" else dict(rtol=0, atol=2e-15)
    with _engine(srb, g, name) as e:
        for k in range(N):
            for c in range(C):
                np.testing.assert_allclose(e.forward(k, x[c]), g[f"{name}_forward"][k, c], **tol)
                np.testing.assert_allclose(e.transpose(k, lr[k, c]), g[f"{name}_transpose"][k, c], **tol)


@pytest.mark.parametrize("name", CASES)
def test_data_term_reference_order_bit_exact(srb, oracle, cv2_fixtures, name):
    """ObjectiveDataTerm::Compute through the reference-order kernels: gradient bit-identical to
    the fixture (cv2) / oracle, cost to 1e-12; accumulation into a non-zero gradient; cost-only."""
    g = cv2_fixtures
    x = g[f"{name}_x"]
    with _engine(srb, g, name) as e:
        e.set_path(srb.PATH_REFERENCE_ORDER)
        grad = np.zeros_like(x)
        cost = e.data_term(x, grad)
        np.testing.assert_allclose(cost, float(g[f"{name}_cost"]), rtol=COST_RTOL)
        np.testing.assert_allclose(grad, g[f"{name}_grad"], rtol=0, atol=0 if name != "k9_s4" else 1e-13)
        # accumulate semantics (objective_data_term.cpp:63-71) against the oracle
        C, H, W, s, psf, sh, N = _case(g, name)
        m = oracle.Model(s, psf, sh, num_frames=N)
        obs = oracle.upsample_observations(m, g[f"{name}_lr"])
        g0 = np.random.default_rng(5).random(x.shape)
        ga, gb = g0.copy(), g0.copy()
        e.data_term(x, ga)
        oracle.data_term(m, x, obs, grad=gb)
        np.testing.assert_array_equal(ga, gb)
        assert e.data_term(x, None) == cost
        # full evaluation without regularizer == data term
        f, ge = e.eval(x)
        assert f == cost
        np.testing.assert_array_equal(ge, grad)
        f2, none = e.eval(x, want_grad=False)
        assert none is None and f2 == cost


@pytest.mark.parametrize("kind,R,decay", [(0, 3, 0.5), (1, 3, 0.5), (2, 1, 0.25), (2, 2, 0.5),
                                          (2, 3, 0.5), (2, 3, 1.0), (2, 4, 0.7)])
def test_regularizers_bit_exact(srb, oracle, ref, kind, R, decay):
    """Regularizer::ApplyToImage / ApplyToImageWithDifferentiation vs the reference's own
    tv_regularizer.cpp / btv_regularizer.cpp (oracle/_ref) and the oracle: bit-for-bit, on an image
    with exact ties."""
    rng = np.random.default_rng(7)
    x = rng.random((3, 9, 11))
    x[0, 2, 3] = x[0, 2, 4]
    x[1, 1, 1] = x[1, 2, 1]
    x[2, 5, 5] = x[1, 5, 5]
    cst = rng.random(x.shape)
    with srb.Engine((1, 3, 9, 11), 1) as e:
        e.set_regularizer(kind, 1.0, R, decay)
        np.testing.assert_array_equal(e.reg_apply(x), ref.reg_apply(kind, x, R, decay))
        v, p = e.reg_apply_diff(x, cst)
        v2, p2 = ref.reg_apply_diff(kind, x, cst, R, decay)
        np.testing.assert_array_equal(v, v2)
        np.testing.assert_array_equal(p, p2)
        vo, po = oracle.reg_apply_diff(kind, x, cst, R, decay)
        np.testing.assert_array_equal(p, po)
        # IRLS re-weighting (irls_map_solver.cpp:128-143)
        np.testing.assert_array_equal(e.reweight(x), oracle.reweight(kind, x, R, decay))


def test_regularizer_reference_golden_values(srb):
    """test/test_tv_regularizer.cpp:62-145 and test/test_btv_regularizer.cpp:21-73, exact."""
    tv_img = np.array([[0, 0, 1], [0, 1, 3], [-3, -1, 0]], dtype=np.float64)
    tv_exp = np.array([[0, 2, 2], [4, 4, 3], [2, 1, 0]], dtype=np.float64)
    with srb.Engine((1, 3, 3, 3), 1) as e:
        e.set_regularizer(srb.REG_TV, 1.0)
        np.testing.assert_array_equal(e.reg_apply(np.stack([tv_img] * 3)), np.stack([tv_exp] * 3))
        e.set_regularizer(srb.REG_TV3D, 1.0)
        x = np.array([tv_img, np.zeros((3, 3)), [[0, -1, 2], [-3, 4, 5], [6, 7, -8]]])
        exp = np.array([[[0, 2, 3], [4, 5, 6], [5, 2, 0]], [[0, 1, 2], [3, 4, 5], [6, 7, 8]],
                        [[4, 8, 3], [16, 4, 13], [1, 15, 0]]], dtype=np.float64)
        np.testing.assert_array_equal(e.reg_apply(x), exp)
    btv_img = np.array([[0, 0, 1, 2, 1], [0, 1, 3, 2, 3], [5, 4, 3, -2, 1], [4, 6, 9, 3, 0],
                        [-3, -1, 0, 6, 0]], dtype=np.float64)
    with srb.Engine((1, 1, 5, 5), 1) as e:
        e.set_regularizer(srb.REG_BTV, 1.0, 2, 0.5)
        v = e.reg_apply(btv_img[None])
        assert v[0, 0, 0] == 2.8125 and v[0, 4, 4] == 0.0


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("name", ["int_s3_neg", "frac_s2"])
def test_compute_all_terms_reference_order_bit_exact(srb, oracle, ref, cv2_fixtures, kind, name):
    """ObjectiveFunction::ComputeAllTerms and ObjectiveIRLSRegularizationTerm::Compute against the
    reference's objective_function.cpp / objective_irls_regularization_term.cpp (oracle/_ref)."""
    g = cv2_fixtures
    C, H, W, s, psf, sh, N = _case(g, name)
    x, lr = g[f"{name}_x"], g[f"{name}_lr"]
    w = 0.5 + np.random.default_rng(3).random(x.shape)
    m = oracle.Model(s, psf, sh, num_frames=N)
    obs = oracle.upsample_observations(m, lr)
    with _engine(srb, g, name) as e:
        e.set_path(srb.PATH_REFERENCE_ORDER)
        e.set_regularizer(kind, 0.01)
        e.set_irls_weights(w)
        g1, g2 = np.full_like(x, 0.25), np.full_like(x, 0.25)
        f1 = e.irls_term(x, g1)
        f2 = ref.irls_term(kind, 0.01, w, x, g2)
        np.testing.assert_allclose(f1, f2, rtol=COST_RTOL)
        np.testing.assert_array_equal(g1, g2)
        fa, ga = e.eval(x)
        fb, gb = ref.compute_all_terms(m, x, obs, kind, 0.01, w)
        np.testing.assert_allclose(fa, fb, rtol=COST_RTOL)
        np.testing.assert_array_equal(ga, gb)
        fc, none = e.eval(x, want_grad=False)
        assert none is None and fc == fa
        # lambda <= 0 removes the term (objective_irls_regularization_term.cpp:15-18)
        e.set_regularizer(kind, 0.0)
        f0, _ = e.eval(x)
        np.testing.assert_allclose(f0, oracle.data_term(m, x, obs)[0], rtol=COST_RTOL)


def test_channel_range_split_channels(srb, oracle, cv2_fixtures):
    """split_channels (irls_map_solver.cpp:200-206): evaluations restricted to [c0, c1)."""
    g = cv2_fixtures
    name = "int_s2"
    C, H, W, s, psf, sh, N = _case(g, name)
    x, lr = g[f"{name}_x"], g[f"{name}_lr"]
    m = oracle.Model(s, psf, sh, num_frames=N)
    obs = oracle.upsample_observations(m, lr)
    with _engine(srb, g, name) as e:
        e.set_path(srb.PATH_REFERENCE_ORDER)
        for c in range(C):
            e.set_channel_range(c, c + 1)
            f, gr = e.eval(x[c:c + 1])
            fo, go = oracle.data_term(m, x[c:c + 1], obs, channel_start=c)
            np.testing.assert_allclose(f, fo, rtol=COST_RTOL)
            np.testing.assert_array_equal(gr, go)
        with pytest.raises(srb.SrbError):
            e.set_channel_range(0, C + 1)


def test_error_statuses(srb):
    """The reference CHECK-fails on these; the C ABI returns a status instead."""
    with pytest.raises(srb.SrbError):
        srb.Engine((0, 1, 4, 4), 2)                      # 0 observations
    with pytest.raises(srb.SrbError):
        srb.Engine((1, 1, 4, 4), 0)                      # scale < 1
    with pytest.raises(srb.SrbError):
        srb.Engine((1, 1, 4, 4), 2, np.ones((4, 4)))     # even blur kernel
    with srb.Engine((2, 1, 4, 4), 2) as e:
        with pytest.raises(srb.SrbError):
            e.eval(np.zeros((1, 8, 8)))                  # observations not set
        with pytest.raises(srb.SrbError):
            e.set_regularizer(srb.REG_BTV, 0.1, 0, 0.5)  # btv_regularizer.cpp:58-61
        with pytest.raises(srb.SrbError):
            e.set_regularizer(srb.REG_BTV, 0.1, 3, 1.5)
        with pytest.raises(srb.SrbError):
            e.forward(5, np.zeros((8, 8)))


# ------------------------------------------------------------------ seeded random cases vs oracle
def _random_case(seed, C, h, w, s, K, N, frac, sigma=1.2):
    from importlib import import_module
    wl = import_module("super-resolution_b200.workloads")
    rng = np.random.default_rng(seed)
    psf = wl.gaussian_psf(K, sigma) if K else None
    if frac:
        shifts = rng.uniform(-2.5, 2.5, size=(N, 2))
    else:
        shifts = rng.integers(-3, 4, size=(N, 2)).astype(np.float64)
    x = rng.random((C, h * s, w * s))
    lr = rng.random((N, C, h, w))
    return psf, shifts, x, lr


@pytest.mark.parametrize("seed,C,h,w,s,K,N,frac", [
    (11, 2, 17, 23, 2, 3, 5, False), (12, 1, 16, 16, 4, 5, 9, False), (13, 3, 9, 14, 3, 7, 4, False),
    (14, 2, 16, 12, 4, 7, 16, False), (15, 1, 13, 11, 2, 5, 3, True), (16, 2, 12, 20, 4, 9, 6, True),
    (17, 1, 40, 40, 4, 7, 16, False), (18, 1, 8, 8, 1, 3, 2, True), (19, 2, 33, 17, 2, 0, 4, True),
])
@pytest.mark.parametrize("kind", [-1, 0, 2])
def test_eval_vs_oracle_all_paths(srb, oracle, seed, C, h, w, s, K, N, frac, kind):
    """Full objective on seeded inputs, every available kernel path, against the oracle."""
    psf, shifts, x, lr = _random_case(seed, C, h, w, s, K, N, frac)
    m = oracle.Model(s, psf, shifts)
    obs = oracle.upsample_observations(m, lr)
    wts = 0.5 + np.random.default_rng(seed + 100).random(x.shape)
    lam = 0.02 if kind >= 0 else 0.0
    fo, go = oracle.evaluate(m, x, obs, max(kind, 0), lam, wts if kind >= 0 else None)
    with srb.Engine(lr.shape, s, psf, shifts) as e:
        e.set_observations(lr)
        if kind >= 0:
            e.set_regularizer(kind, lam)
            e.set_irls_weights(wts)
        e.set_path(srb.PATH_REFERENCE_ORDER)
        f, gr = e.eval(x)
        np.testing.assert_allclose(f, fo, rtol=COST_RTOL)
        np.testing.assert_array_equal(gr, go)
        e.set_path(srb.PATH_AUTO)
        if e.active_path == srb.PATH_FUSED:
            f, gr = e.eval(x)
            np.testing.assert_allclose(f, fo, rtol=COST_RTOL)
            assert rel_l2(gr, go) <= FUSED_REL_L2
            assert np.abs(gr - go).max() <= 1e-11 * np.abs(go).max()
            f2, none = e.eval(x, want_grad=False)
            np.testing.assert_allclose(f2, fo, rtol=COST_RTOL)


def test_forward_all_equals_per_frame_forward(srb, oracle):
    """srb_forward_all (the whole LR stack in one launch) == srb_forward frame by frame == oracle,
    bit for bit (same kernel, same operation order as the reference)."""
    rng = np.random.default_rng(77)
    C, h, w, s, K, N = 3, 21, 17, 3, 5, 6
    psf = oracle.gaussian_psf(K, 1.3)
    shifts = np.vstack([rng.integers(-3, 4, size=(3, 2)).astype(np.float64), rng.uniform(-2.5, 2.5, size=(3, 2))])
    x = rng.random((C, h * s, w * s))
    m = oracle.Model(s, psf, shifts)
    with srb.Engine((N, C, h, w), s, psf, shifts) as e:
        e.set_channel_range(1, 2)          # forward_all ignores the active channel range
        stack = e.forward_all(x)
        for k in range(N):
            for c in range(C):
                np.testing.assert_array_equal(stack[k, c], e.forward(k, x[c]))
                np.testing.assert_array_equal(stack[k, c], oracle.forward(m, k, x[c]))
