"""Host logic of the fused path, on CPU (no GPU needed): srb_plan / srb_quantize_shift /
srb_sample_is_special are pure host code in libsrb200.so.

The fused tile kernel commutes the PSF with the shift, which is exact only for
"regular" LR samples; the planner must flag every sample for which it is not.  That rule is checked
here against brute force with the CPU oracle: a sample is TRULY special when moving the blur across
the warp changes its LR prediction (forward) or its back-projected HR image (transpose)."""
import importlib

import numpy as np
import pytest

srb = importlib.import_module("super-resolution_b200")
wl = importlib.import_module("super-resolution_b200.workloads")


def test_warp_quantisation_matches_the_oracle(oracle):
    """cv::warpAffine's fixed-point translation: library host code == oracle (pinned against cv2)."""
    rng = np.random.default_rng(0)
    shifts = np.concatenate([np.linspace(-3, 3, 1201), rng.uniform(-40, 40, 500), [0.015625, -0.015625, 1 / 64 + 1e-12]])
    for d in shifts:
        # the library quantises the FORWARD warp of shift d: source offset of -d
        assert srb.quantize_shift(d) == oracle.warp_quantize(d), d


def test_plans_of_the_baseline_configurations():
    # z: row-major Z layout (k_tile_z, round 1) -- 1 one frame on every phase, 2 at most one, 0 does not qualify;
    # zt_frames (below): the transposed Z layout k_tile_zt, the default kernel of all five configurations
    expect = {1: dict(frac=0, kh=1, per_phase=(1, 1), table=1, z=1),
              2: dict(frac=0, kh=2, per_phase=(0, 1), table=0, z=2),     # 9 frames over 16 phases
              3: dict(frac=0, kh=3, per_phase=(1, 1), table=1, z=1),
              4: dict(frac=0, kh=2, per_phase=(2, 2), table=2, z=0),     # 8 frames over 4 phases
              5: dict(frac=0, kh=4, per_phase=(4, 4), table=4, z=0)}     # 64 frames over 16 phases
    for cfg, ex in expect.items():
        cf = wl.CONFIGS[cfg]
        s = cf["s"]
        shifts = np.array(cf["shifts"], float) if "shifts" in cf else wl.default_shifts(cf["N"], s)
        p = srb.plan((cf["N"], cf["C"], cf["H"] // s, cf["W"] // s), s, wl.gaussian_psf(cf["K"], cf["sigma"]), shifts)
        assert p["fused"] == 1, (cfg, p["why"])
        assert (p["hr_height"], p["hr_width"]) == (cf["H"], cf["W"])
        assert p["fractional"] == ex["frac"] and p["psf_half"] == ex["kh"]
        assert (p["min_entries_per_phase"], p["max_entries_per_phase"]) == ex["per_phase"], (cfg, p)
        assert p["table_driven"] == ex["table"]      # frames per phase of the table-driven residual pass (0: generic)
        assert p["zlayout"] == ex["z"], (cfg, p)
        # k_tile_zt: frames per non-empty phase, all with the same shift, averaged at upload
        assert p["zt_frames"] == max(1, cf["N"] // (s * s)), (cfg, p)
        assert p["num_entries"] == cf["N"]
        # default shifts are >= 0 and at most s - 1: at most a thin band at the top / left
        assert p["band_hi_r"] >= cf["H"] // s - 2 and p["band_hi_c"] >= cf["W"] // s - 2
        assert p["band_lo_r"] <= 1 and p["band_lo_c"] <= 1


def test_models_the_tile_kernel_does_not_cover_are_reported_not_failed():
    psf = wl.gaussian_psf(5, 1.5)
    rough = psf.copy()
    rough[0, 1] += 0.01                                   # not rank 1
    p = srb.plan((2, 1, 16, 16), 2, rough, [(0, 0), (1, 0)])
    assert p["fused"] == 0 and "separable" in p["why"]
    p = srb.plan((2, 1, 16, 16), 2, wl.gaussian_psf(11, 3.0), [(0, 0), (1, 0)])
    assert p["fused"] == 0 and "9x9" in p["why"]
    p = srb.plan((2, 1, 4, 4), 2, psf, [(0, 0), (7, 0)])  # shift larger than the image allows
    assert p["fused"] == 0 and "small" in p["why"]
    with pytest.raises(srb.SrbError):                      # the reference CHECK-fails on these
        srb.plan((0, 1, 4, 4), 2, psf, None)
    with pytest.raises(srb.SrbError):
        srb.plan((1, 1, 4, 4), 2, wl.gaussian_psf(4, 1.0), None)


@pytest.mark.parametrize("s,K", [(2, 5), (4, 7), (3, 3), (2, 9)])
def test_special_sample_rule_covers_brute_force(oracle, s, K):
    """Vertical shifts only: for every LR row q, (truly special by brute force) => (planner flags it)."""
    o = oracle
    h, w = 14, 6
    H, W = h * s, w * s
    hk = K // 2
    psf = wl.gaussian_psf(K, 0.4 * K)
    rng = np.random.default_rng(s * 10 + K)
    x = rng.random((H, W)) + 0.5
    flagged_total = truly_total = 0
    P = s * (hk + 6)                                  # zero margin, a multiple of s
    xp = np.pad(x, P)                                 # the kernel blurs the zero-EXTENDED estimate ...
    bx_ext = o.filter2d(xp, psf)                      # ... so Bx also exists just outside the image
    for dy in [0.0, 1.0, -1.0, 2.0, -3.0, float(hk), float(-hk), hk + 1.0, -(hk + 2.0), 0.3, -0.3, 1.7, -2.4, 0.5]:
        m = o.Model(s, psf, [(0.0, dy)])
        # forward: D B M x (reference) vs D M (B x) evaluated on the extended domain (tile kernel)
        exact = o.forward(m, 0, x)
        commuted = o.warp_shift(bx_ext, 0.0, dy)[P:P + H:s, P:P + W:s]
        fwd_special = np.abs(exact - commuted).max(axis=1) > 1e-12
        # transpose: M^T B^T D^T e_q (reference) vs B^T (M^T D^T e_q) on the extended domain, cropped
        tr_special = np.zeros(h, bool)
        for q in range(h):
            e = np.zeros((h, w))
            e[q, 2] = 1.0                             # column 2: away from the left / right border
            exact_t = o.transpose(m, 0, e)
            up = np.zeros((H + 2 * P, W + 2 * P))
            up[P + s * q, P + s * 2] = 1.0            # zero insertion
            commuted_t = o.filter2d(o.warp_shift(up, 0.0, -dy), psf.T.copy())[P:P + H, P:P + W]
            tr_special[q] = np.abs(exact_t - commuted_t).max() > 1e-12
        for q in range(h):
            flagged = srb.sample_is_special(q, H, hk, s, dy)
            truly = bool(fwd_special[q] or tr_special[q])
            assert flagged or not truly, (s, K, dy, q)
            flagged_total += flagged
            truly_total += truly
    assert truly_total > 0                      # the cases do exercise the rule ...
    assert flagged_total <= 4 * truly_total + 40  # ... and the rule stays a thin band, not everything


def test_solver_options_mirror_the_reference_defaults_and_scaling():
    """IrlsMapSolverOptions: defaults of map_solver.h:54-62 / irls_map_solver.h:27-35 and
    AdjustThresholdsAdaptively (map_solver.cpp:16-26, irls_map_solver.cpp:161-171): thresholds scale
    by num_parameters * sum(lambda), up only."""
    from importlib import import_module
    solver = import_module("super-resolution_b200.solver")
    o = solver.IrlsMapSolverOptions()
    assert (o.max_num_solver_iterations, o.max_num_irls_iterations) == (50, 20)
    assert (o.gradient_norm_threshold, o.cost_decrease_threshold, o.parameter_variation_threshold,
            o.irls_cost_difference_threshold) == (1e-6, 1e-6, 1e-6, 1e-5)
    same = o.adjusted(10, 0.01)             # scale 0.1 < 1: unchanged
    assert same == o and same is not o
    up = o.adjusted(2352, 0.01)             # cfg1: 28*28*3 parameters, lambda 0.01
    assert up.gradient_norm_threshold == 1e-6 * (2352 * 0.01)
    assert up.cost_decrease_threshold == 1e-6 * (2352 * 0.01)
    assert up.parameter_variation_threshold == 1e-6 * (2352 * 0.01)
    assert up.irls_cost_difference_threshold == 1e-5 * (2352 * 0.01)
    assert (up.max_num_solver_iterations, up.max_num_irls_iterations) == (50, 20)


def test_zlayout_qualification_of_frame_shards_and_other_models():
    """cfg3's model split over 2 / 4 / 8 ranks (sharding.frame_shard): every shard has at most one frame
    per phase -> Z layout with holes; fractional shifts, no PSF and PSFs wider than 9x9 do not qualify."""
    from importlib import import_module
    sharding = import_module("super-resolution_b200.sharding")
    cf = wl.CONFIGS[3]
    s = cf["s"]
    shifts = wl.default_shifts(cf["N"], s)
    psf = wl.gaussian_psf(cf["K"], cf["sigma"])
    for world in (2, 4, 8):
        for rank in range(world):
            fr = sharding.frame_shard(cf["N"], rank, world)
            p = srb.plan((len(fr), cf["C"], cf["H"] // s, cf["W"] // s), s, psf, shifts[fr])
            assert p["fused"] == 1 and p["zlayout"] == 2 and p["table_driven"] == 0, (world, rank, p)
    frac = shifts.copy()
    frac[3, 0] += 0.5
    assert srb.plan((16, 1, 64, 64), s, psf, frac)["zlayout"] == 0
    assert srb.plan((16, 1, 64, 64), s, psf, frac)["zt_frames"] == 0
    far = np.concatenate([shifts, shifts + np.array([s, 0.0])])            # same phases, different shifts
    assert srb.plan((32, 1, 64, 64), s, psf, far)["zt_frames"] == 0
    assert srb.plan((32, 1, 64, 64), s, psf, np.concatenate([shifts, shifts]))["zt_frames"] == 2
    assert srb.plan((20, 1, 64, 64), s, psf, np.concatenate([shifts, shifts[:4]]))["zt_frames"] == 0   # unequal counts
    assert srb.plan((16, 1, 64, 64), s, None, shifts)["zlayout"] == 0          # no PSF: nothing to gain
    assert srb.plan((4, 1, 64, 64), 2, wl.gaussian_psf(3, 0.8), wl.default_shifts(4, 2))["zlayout"] == 1
