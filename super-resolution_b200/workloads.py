"""Synthetic workloads of BASELINE.json's five configurations, numpy only.

Everything is deterministic (`numpy.random.default_rng(seed)`, seed = configuration number).
LR stacks are produced by a caller-supplied forward model `forward(k, hr_plane) -> lr_plane`
(the engine's srb_forward on the GPU, or the oracle in CPU tests) so that this module itself
contains no image-model arithmetic.
"""
import numpy as np

# cfg -> (H, W, C, N, s, K, sigma, regularizer, seed)
CONFIGS = {
    1: dict(H=28, W=28, C=3, N=4, s=2, K=3, sigma=1.0, reg="tv", seed=1,
            shifts=[(0, 0), (1, 1), (0, 1), (1, 0)], noise=0.0,
            name="cfg1 28x28x3 N=4 s=2 K=3 TV (fb.png-shaped, test_motion_sequence_4)"),
    2: dict(H=512, W=512, C=1, N=9, s=4, K=5, sigma=1.5, reg="btv", seed=2,
            name="cfg2 512x512x1 N=9 s=4 K=5 BTV"),
    3: dict(H=2048, W=2048, C=3, N=16, s=4, K=7, sigma=2.0, reg="tv", seed=3,
            name="cfg3 2048x2048x3 N=16 s=4 K=7 TV"),
    4: dict(H=1024, W=1024, C=128, N=8, s=2, K=5, sigma=1.5, reg="tv", seed=4,
            name="cfg4 1024x1024x128 N=8 s=2 K=5 TV (hyperspectral)"),
    5: dict(H=4096, W=4096, C=3, N=64, s=4, K=9, sigma=2.5, reg="btv", seed=5,
            name="cfg5 4096x4096x3 N=64 s=4 K=9 BTV"),
}
REG_KIND = {"none": -1, "tv": 0, "tv3d": 1, "btv": 2}
LAMBDA = 0.01
BTV_RANGE, BTV_DECAY = 3, 0.5
NOISE_SIGMA = 2.0 / 255.0


def gaussian_kernel(n, sigma):
    """cv::getGaussianKernel(n, sigma, CV_64F) for sigma > 0 (blur_module.cpp:20-21)."""
    x = np.arange(n, dtype=np.float64) - (n - 1) * 0.5
    t = np.exp((-0.5 / (sigma * sigma)) * x * x)
    return t * (1.0 / t.sum())


def gaussian_psf(n, sigma):
    """blur_kernel_ = kernel_x * kernel_y.t() (blur_module.cpp:22)."""
    g = gaussian_kernel(n, sigma)
    return np.outer(g, g)


def default_shifts(N, s):
    """Integer HR shifts (k mod s, floor(k/s) mod s) -- every sub-pixel phase of the LR grid."""
    return np.array([(k % s, (k // s) % s) for k in range(N)], dtype=np.float64)


def box_smooth(img, k=5):
    """k x k box filter with edge replication, per channel (makes TV/BTV meaningful)."""
    pad = k // 2
    out = np.empty_like(img)
    for c in range(img.shape[0]):
        a = np.pad(img[c], pad, mode="edge")
        cs = np.cumsum(np.cumsum(a, axis=0), axis=1)
        cs = np.pad(cs, ((1, 0), (1, 0)))
        H, W = img.shape[1:]
        out[c] = (cs[k:k + H, k:k + W] - cs[0:H, k:k + W] - cs[k:k + H, 0:W] + cs[0:H, 0:W]) / (k * k)
    return out


def ground_truth(H, W, C, seed):
    rng = np.random.default_rng(seed)
    return box_smooth(rng.random((C, H, W)))


def bilinear_upsample(lr, s):
    """Bilinear interpolation with half-pixel centres and edge replication (the initial estimate
    of super_resolution.cpp:371-373 is LR frame 0 upsampled bilinearly)."""
    Cn, h, w = lr.shape
    H, W = h * s, w * s
    ry = np.clip((np.arange(H) + 0.5) / s - 0.5, 0, h - 1)
    rx = np.clip((np.arange(W) + 0.5) / s - 0.5, 0, w - 1)
    y0 = np.floor(ry).astype(int)
    x0 = np.floor(rx).astype(int)
    y1 = np.minimum(y0 + 1, h - 1)
    x1 = np.minimum(x0 + 1, w - 1)
    fy = (ry - y0)[None, :, None]
    fx = (rx - x0)[None, None, :]
    a = lr[:, y0][:, :, x0] * (1 - fx) + lr[:, y0][:, :, x1] * fx
    b = lr[:, y1][:, :, x0] * (1 - fx) + lr[:, y1][:, :, x1] * fx
    return a * (1 - fy) + b * fy


def make(cfg, forward=None, H=None, W=None, C=None, N=None, frames=None, cheap=False):
    """Builds one workload.  `forward(k, hr_plane)` degrades one HR channel plane to LR for frame k;
    when None (or cheap=True) the LR stack is decimated smoothed truth plus noise -- same shapes
    and statistics, used where only throughput matters.  H/W/C/N override the configuration's
    size (scaled-down parity cases); `frames` selects a subset of frame indices (a rank's shard).
    Returns dict(x_true, x0, lr, psf, shifts, s, K, reg_kind, lam, btv_range, btv_decay, name)."""
    cf = dict(CONFIGS[cfg])
    H = H or cf["H"]
    W = W or cf["W"]
    Cn = C or cf["C"]
    Nn = N or cf["N"]
    s, K = cf["s"], cf["K"]
    psf = gaussian_psf(K, cf["sigma"])
    shifts = np.array(cf["shifts"], dtype=np.float64) if "shifts" in cf and Nn == cf["N"] \
        else default_shifts(Nn, s)
    x_true = ground_truth(H, W, Cn, cf["seed"])
    h, w = H // s, W // s
    ks = list(range(Nn)) if frames is None else list(frames)
    rng = np.random.default_rng(1000 + cf["seed"])
    noise_sigma = cf.get("noise", NOISE_SIGMA)
    lr = np.empty((len(ks), Cn, h, w))
    for i, k in enumerate(ks):
        for c in range(Cn):
            if forward is None or cheap:
                dx, dy = int(shifts[k][0]), int(shifts[k][1])
                sm = np.roll(x_true[c], (dy, dx), axis=(0, 1))
                lr[i, c] = sm[::s, ::s][:h, :w]
            else:
                lr[i, c] = forward(k, x_true[c])
        # noise drawn per frame so that a frame shard sees the same values as the full stack
        frame_rng = np.random.default_rng([1000 + cf["seed"], k])
        if noise_sigma > 0:
            lr[i] += frame_rng.normal(0.0, noise_sigma, size=lr[i].shape)
    del rng
    x0 = bilinear_upsample(lr[0], s) if 0 in ks else bilinear_upsample(lr[0], s)
    return dict(x_true=x_true, x0=x0, lr=lr, psf=psf, shifts=shifts[ks], all_shifts=shifts, s=s, K=K,
                reg_kind=REG_KIND[cf["reg"]], lam=LAMBDA, btv_range=BTV_RANGE, btv_decay=BTV_DECAY,
                name=cf["name"], H=H, W=W, C=Cn, N=len(ks), frames=ks)


def algorithmic_bytes(H, W, C, N, s, has_reg=True, elem=8):
    """read x, read IRLS weights, write g, read every LR observation once."""
    return int(elem * C * H * W * ((3 if has_reg else 2) + N / float(s * s)))


def work_units(H, W, C, N):
    """HR px * frames * channels processed by one evaluation (the metric's unit)."""
    return H * W * C * N
