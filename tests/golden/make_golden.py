#!/usr/bin/env python
"""Generates tests/golden/cv2_fixtures.npz -- run in the BUILD container only (needs cv2).

The reference delegates all data-term arithmetic to four OpenCV entry points
(motion_module.cpp:23 warpAffine, matrix_util.cpp:20-27 filter2D, image_data.cpp:341-347 resize,
blur_module.cpp:20-22 getGaussianKernel).  OpenCV is not vendored by the reference and its C++
headers are absent here, but the same entry points are callable through the cv2 4.13 wheel.  This
script calls them exactly the way the reference does and stores small input/output pairs; the
oracle (oracle/sr_oracle.c) and the CUDA path are both tested against them
(tests/test_oracle_golden.py, tests/test_gpu_parity.py).

It also restates ObjectiveDataTerm::Compute (objective_data_term.cpp:15-116) with those cv2
calls -- same order of operations as the reference -- and stores cost / gradient for a few small
problems, including fractional shifts, so the fused formulas are pinned end to end.

Usage:  python tests/golden/make_golden.py        (writes cv2_fixtures.npz next to this file)
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
cv2.setNumThreads(1)


def warp(img, dx, dy):
    """motion_module.cpp:18-24: warpAffine(channel, channel, [1 0 dx; 0 1 dy], size)."""
    m = np.array([[1.0, 0.0, dx], [0.0, 1.0, dy]], dtype=np.float64)
    return cv2.warpAffine(img, m, (img.shape[1], img.shape[0]))


def blur(img, kernel):
    """matrix_util.cpp:20-27: filter2D(..., anchor (-1,-1), delta 0, BORDER_CONSTANT)."""
    return cv2.filter2D(img, -1, kernel, anchor=(-1, -1), delta=0, borderType=cv2.BORDER_CONSTANT)


def resize_nn(img, w2, h2):
    """image_data.cpp:341-347."""
    return cv2.resize(img, (w2, h2), fx=0, fy=0, interpolation=cv2.INTER_NEAREST)


def gaussian_psf(n, sigma):
    """blur_module.cpp:20-22."""
    kx = cv2.getGaussianKernel(n, sigma)
    ky = cv2.getGaussianKernel(n, sigma)
    return kx @ ky.T


def additive_down(img, s):
    """image_data.cpp:116-133 (sum pool)."""
    H, W = img.shape
    out = np.zeros((H // s, W // s))
    for r in range(H):
        for c in range(W):
            out[r // s, c // s] += img[r, c]
    return out


def additive_up(img, s):
    """image_data.cpp:99-115 (zero insertion)."""
    h, w = img.shape
    out = np.zeros((h * s, w * s))
    out[::s, ::s] = img
    return out


def forward(x, s, psf, shift):
    y = x
    if shift is not None:
        y = warp(y, shift[0], shift[1])
    if psf is not None:
        y = blur(y, psf)
    H, W = y.shape
    f = 1.0 / float(s)
    return resize_nn(y, int(W * f), int(H * f))


def transpose(r, s, psf, shift):
    y = additive_up(r, s)
    if psf is not None:
        y = blur(y, np.ascontiguousarray(psf.T))
    if shift is not None:
        y = warp(y, -shift[0], -shift[1])
    return y


def data_term(x, obs_hr, s, psf, shifts):
    """objective_data_term.cpp:15-116 with cv2 calls.  x [C][H][W]; obs_hr [N][C][H][W]."""
    C, H, W = x.shape
    cost = 0.0
    grad = np.zeros_like(x)
    for k in range(obs_hr.shape[0]):
        shift = None if shifts is None else shifts[k]
        for c in range(C):
            d = forward(x[c], s, psf, shift)
            d = resize_nn(d, W, H)
            res = d - obs_hr[k, c]
            cost += float(np.sum(res * res))
            rl = additive_down(res, s)
            grad[c] += 2 * transpose(rl, s, psf, shift)
    return cost, grad


def main():
    rng = np.random.default_rng(20260101)
    out = {}

    # 1. Gaussian kernels
    gk = [(3, 1.0), (3, 0.849321), (5, 1.5), (7, 2.0), (9, 2.5), (3, 3.0), (5, 0.7)]
    out["gauss_params"] = np.array(gk, dtype=np.float64)
    for i, (n, sg) in enumerate(gk):
        out[f"gauss_{i}"] = cv2.getGaussianKernel(n, sg).ravel()
        out[f"gauss_psf_{i}"] = gaussian_psf(n, sg)

    # 2. warpAffine: shift quantisation sweep on a ramp + values on a random image
    ramp = np.tile(np.arange(64, dtype=np.float64), (4, 1))
    sweep = np.linspace(-2.5, 2.5, 4001)
    nq = []
    for d in sweep:
        w = warp(ramp, d, 0.0)
        nq.append(int(round((w[1, 32] - 32.0) * 32)))
    out["warp_sweep_d"] = sweep
    out["warp_sweep_n"] = np.array(nq, dtype=np.int32)
    rampy = np.ascontiguousarray(ramp.T)
    sweep_y = np.linspace(-2.5, 2.5, 1001)
    nqy = []
    for d in sweep_y:
        w = warp(rampy, 0.0, d)
        nqy.append(int(round((w[32, 1] - 32.0) * 32)))
    out["warp_sweep_dy"] = sweep_y
    out["warp_sweep_ny"] = np.array(nqy, dtype=np.int32)

    img = rng.random((13, 17))
    shifts = [(0, 0), (1, 1), (-1, 0), (3, -2), (20, 0), (0, -13), (0.3, 0.0), (0.5, 0.5),
              (-0.7, 1.25), (0.016, 0.49), (2.5, -1.5), (-3.96875, 0.984375), (0.484375, -0.015625)]
    out["warp_img"] = img
    out["warp_shifts"] = np.array(shifts, dtype=np.float64)
    for i, (dx, dy) in enumerate(shifts):
        out[f"warp_{i}"] = warp(img, dx, dy)

    # 3. filter2D (zero border), symmetric Gaussian and an asymmetric kernel + its transpose
    out["filt_img"] = img
    for i, (n, sg) in enumerate([(3, 0.849321), (5, 1.5), (7, 2.0), (9, 2.5)]):
        out[f"filt_gauss_{i}"] = blur(img, gaussian_psf(n, sg))
    asym = rng.random((5, 5))
    out["filt_asym_kernel"] = asym
    out["filt_asym"] = blur(img, asym)
    out["filt_asym_t"] = blur(img, np.ascontiguousarray(asym.T))

    # 4. nearest resize index maps, incl. non-divisible sizes (src n, dst n2)
    pairs = []
    maps = []
    for n in list(range(1, 40)) + [100, 511, 512, 2048]:
        for s in (1, 2, 3, 4, 5):
            n2 = int(n * (1.0 / s))
            if n2 < 1:
                continue
            src = np.arange(n, dtype=np.float64).reshape(1, n)
            m = resize_nn(src, n2, 1).ravel().astype(np.int32)
            pairs.append((n, n2, len(maps)))
            maps.extend(m.tolist())
            # upsample back n2 -> n2*s
            src2 = np.arange(n2, dtype=np.float64).reshape(1, n2)
            m2 = resize_nn(src2, n2 * s, 1).ravel().astype(np.int32)
            pairs.append((n2, n2 * s, len(maps)))
            maps.extend(m2.tolist())
    out["nn_pairs"] = np.array(pairs, dtype=np.int64)
    out["nn_maps"] = np.array(maps, dtype=np.int32)

    # 5. forward / transpose / data term on small problems
    cases = [
        # name, C, H, W, s, K, sigma, shifts
        ("int_s2", 2, 12, 16, 2, 3, 1.0, [(0, 0), (1, 1), (0, 1), (1, 0)]),
        ("int_s4", 1, 24, 20, 4, 5, 1.5, [(k % 4, (k // 4) % 4) for k in range(9)]),
        ("int_s3_neg", 3, 15, 18, 3, 7, 2.0, [(0, 0), (-1, 2), (2, -2), (-4, 5), (1, 0)]),
        ("frac_s2", 2, 14, 10, 2, 3, 0.8, [(0, 0), (0.3, -0.6), (1.5, 0.25), (-0.75, 2.031)]),
        ("frac_s4", 1, 16, 24, 4, 7, 2.0, [(0.5, 0.5), (-1.2, 3.4), (2.9, -0.1)]),
        ("noblur", 1, 8, 8, 2, 0, 0.0, [(0, 0), (-1, 0), (0, -1), (-1, -1)]),
        ("nomotion", 2, 9, 12, 3, 3, 3.0, None),
        ("k9_s4", 1, 32, 28, 4, 9, 2.5, [(0, 0), (3, 1), (2, 2)]),
    ]
    out["case_names"] = np.array([c[0] for c in cases])
    for name, C, H, W, s, K, sg, sh in cases:
        psf = gaussian_psf(K, sg) if K > 0 else None
        x_true = rng.random((C, H, W))
        x = rng.random((C, H, W))
        n_frames = 3 if sh is None else len(sh)
        lr = np.stack([np.stack([forward(x_true[c], s, psf, None if sh is None else sh[k])
                                 for c in range(C)]) for k in range(n_frames)])
        lr += 0.01 * rng.standard_normal(lr.shape)
        obs_hr = np.stack([np.stack([resize_nn(lr[k, c], W, H) for c in range(C)])
                           for k in range(n_frames)])
        cost, grad = data_term(x, obs_hr, s, psf, sh)
        fw = np.stack([np.stack([forward(x[c], s, psf, None if sh is None else sh[k])
                                 for c in range(C)]) for k in range(n_frames)])
        tr = np.stack([np.stack([transpose(lr[k, c], s, psf, None if sh is None else sh[k])
                                 for c in range(C)]) for k in range(n_frames)])
        out[f"{name}_meta"] = np.array([C, H, W, s, K, n_frames], dtype=np.int64)
        out[f"{name}_psf"] = psf if psf is not None else np.zeros((0, 0))
        out[f"{name}_shifts"] = (np.array(sh, dtype=np.float64) if sh is not None
                                 else np.zeros((0, 2)))
        out[f"{name}_x"] = x
        out[f"{name}_lr"] = lr
        out[f"{name}_forward"] = fw
        out[f"{name}_transpose"] = tr
        out[f"{name}_cost"] = np.array(cost)
        out[f"{name}_grad"] = grad


    # 6. The reference's own test image (test_data/fb.png, 28x28x3) as decoded arrays, the
    #    RealIconDataTest inputs (test/test_map_solver.cpp:205-308) and the config-1 inputs
    #    (BASELINE.json configs[0]: 4-frame shift sequence, 2x, 3x3 Gaussian PSF sigma 1, TV).
    ref_data = "/root/reference/test_data"
    gray = cv2.imread(os.path.join(ref_data, "fb.png"), cv2.IMREAD_GRAYSCALE)
    bgr = cv2.imread(os.path.join(ref_data, "fb.png"), cv2.IMREAD_UNCHANGED)
    out["fb_gray_u8"] = gray
    out["fb_bgr_u8"] = bgr
    truth = gray.astype(np.float64) / 255.0          # ImageData(cv::Mat) normalises by 1/255
    icon_shifts = [(0, 0), (1, 0), (0, 1), (1, 1)]
    icon_lr = np.stack([forward(truth, 2, None, sh)[None] for sh in icon_shifts])
    out["icon_shifts"] = np.array(icon_shifts, dtype=np.float64)
    out["icon_lr"] = icon_lr
    out["icon_x0"] = cv2.resize(icon_lr[0, 0], (28, 28), fx=0, fy=0,
                                interpolation=cv2.INTER_LINEAR)[None]
    truth3 = np.ascontiguousarray(bgr.transpose(2, 0, 1)).astype(np.float64) / 255.0
    with open(os.path.join(ref_data, "test_motion_sequence_4.txt")) as f:
        seq = [tuple(float(v) for v in line.split()) for line in f if line.strip()]
    psf1 = gaussian_psf(3, 1.0)
    cfg1_lr = np.stack([np.stack([forward(truth3[c], 2, psf1, sh) for c in range(3)])
                        for sh in seq])
    out["cfg1_shifts"] = np.array(seq, dtype=np.float64)
    out["cfg1_psf"] = psf1
    out["cfg1_lr"] = cfg1_lr
    out["cfg1_x0"] = np.stack([cv2.resize(cfg1_lr[0, c], (28, 28), fx=0, fy=0,
                                          interpolation=cv2.INTER_LINEAR) for c in range(3)])

    path = os.path.join(HERE, "cv2_fixtures.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; cv2", cv2.__version__)


if __name__ == "__main__":
    main()
