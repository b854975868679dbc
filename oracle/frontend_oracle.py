"""CPU restatement (numpy / pure Python) of the steps either side of the hot path -- SURVEY.md section 8f, rows
N2 - N4.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this
module; the product (libsrb200.so and the package) never does.

Pinned against: the reference's own golden values (test/test_evaluation.cpp:12-140, test/
test_hyperspectral_data_loader.cpp:35-86, test_data/example_envi_data's value pattern), Random123's known-answer
vectors for Philox4x32-10, and fixtures produced by the same OpenCV entry points in Python cv2 4.13
(tests/golden/make_frontend_golden.py -> frontend_fixtures.npz): cv2.resize(INTER_LINEAR), cv2.PCACompute2 /
PCAProject / PCABackProject.
"""
import math
import struct

import numpy as np


# ---- N3: cv::resize(INTER_LINEAR) as ImageData::ResizeImage calls it (image_data.cpp:310-350) ----------------
def linear_coefficients(n_dst, n_src):
    """OpenCV resize.cpp (resizeGeneric_ set-up): f = (d + 0.5) * (n_src / n_dst) - 0.5, i = floor(f), w = f - i;
    i < 0 -> (0, 0); i >= n_src - 1 -> (n_src - 1, 0)."""
    scale = float(n_src) / float(n_dst)
    i0 = np.empty(n_dst, dtype=np.int64)
    w = np.empty(n_dst, dtype=np.float64)
    for d in range(n_dst):
        f = (d + 0.5) * scale - 0.5
        i = math.floor(f)
        f -= i
        if i < 0:
            i, f = 0, 0.0
        if i >= n_src - 1:
            i, f = n_src - 1, 0.0
        i0[d], w[d] = i, f
    return i0, np.minimum(i0 + 1, n_src - 1), w


def resize_linear(src, H, W):
    """[C][h][w] -> [C][H][W]: horizontal pass a*(1-w) + b*w per source row, then the vertical pass."""
    src = np.asarray(src, dtype=np.float64)
    Cn, h, w = src.shape
    x0, x1, wx = linear_coefficients(W, w)
    y0, y1, wy = linear_coefficients(H, h)
    rows = src[:, :, x0] * (1.0 - wx) + src[:, :, x1] * wx
    return rows[:, y0, :] * (1.0 - wy)[None, :, None] + rows[:, y1, :] * wy[None, :, None]


# ---- N3: src/evaluation ---------------------------------------------------------------------------------------
def psnr(image, truth):
    """PeakSignalToNoiseRatioEvaluator::Evaluate (peak_signal_to_noise_ratio.cpp:29-52), sums in its order."""
    a = np.asarray(image, dtype=np.float64).reshape(-1)
    b = np.asarray(truth, dtype=np.float64).reshape(-1)
    d = b - a
    ssd = float(np.cumsum(d * d)[-1])          # sequential, like the reference's loop
    mse = ssd / float(a.size)
    if mse == 0.0:
        return math.inf
    return 20.0 * math.log10(1.0) - 10.0 * math.log10(mse)


def ssim(image, truth, k1=0.01, k2=0.03, image_scale=1.0):
    """StructuralSimilarityEvaluator (structural_similarity.cpp:9-103): global statistics, not windowed."""
    a = np.asarray(image, dtype=np.float64).reshape(-1)
    b = np.asarray(truth, dtype=np.float64).reshape(-1)
    n = float(a.size)
    mean_t = float(np.cumsum(b)[-1]) / n
    mean_i = float(np.cumsum(a)[-1]) / n
    var_t = float(np.cumsum((b - mean_t) * (b - mean_t))[-1]) / n
    var_i = float(np.cumsum((a - mean_i) * (a - mean_i))[-1]) / n
    cov = float(np.cumsum((a - mean_i) * (b - mean_t))[-1]) / n
    c1 = (k1 * image_scale) ** 2
    c2 = (k2 * image_scale) ** 2
    return ((2 * mean_t * mean_i + c1) * (2 * cov + c2)) / ((mean_t * mean_t + mean_i * mean_i + c1) * (var_t + var_i + c2))


# ---- N2: the noise generator of srb_add_noise (Philox4x32-10 + Box-Muller) -------------------------------------
M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter, key):
    """Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3" (SC'11); Random123 philox4x32."""
    c = [int(v) & 0xFFFFFFFF for v in counter]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c


def philox_words(groups, seed, stream=0):
    """Vectorised philox4x32_10 for counters (g, stream), g = 0 .. groups-1, key = seed.  Returns [groups][4] uint64."""
    g = np.arange(groups, dtype=np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    c = [g & mask, g >> np.uint64(32), np.full(groups, stream & 0xFFFFFFFF, np.uint64), np.full(groups, (stream >> 32) & 0xFFFFFFFF, np.uint64)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & mask, p1 & mask, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & mask, p0 & mask]
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return np.stack(c, axis=1)


def standard_normals(n, seed, stream=0):
    """z_0 .. z_{n-1} exactly as k_add_noise forms them: samples 4g .. 4g+3 from counter (g, stream), two Box-Muller
    pairs with u = (r + 0.5) * 2^-32."""
    groups = (n + 3) // 4
    wds = philox_words(groups, seed, stream).astype(np.float64)
    u = (wds + 0.5) * 2.0 ** -32
    z = np.empty((groups, 4))
    for p in range(2):
        r = np.sqrt(-2.0 * np.log(u[:, 2 * p]))
        z[:, 2 * p] = r * np.cos(2.0 * np.pi * u[:, 2 * p + 1])
        z[:, 2 * p + 1] = r * np.sin(2.0 * np.pi * u[:, 2 * p + 1])
    return z.reshape(-1)[:n]


def add_noise(data, sigma, seed, stream=0):
    """AdditiveNoiseModule::ApplyToImage (additive_noise_module.cpp:19-36): sigma is on the 0..255 scale."""
    data = np.asarray(data, dtype=np.float64)
    return data + (sigma / 255.0) * standard_normals(data.size, seed, stream).reshape(data.shape)


# ---- N4: ENVI float32 BSQ (hyperspectral_data_loader.cpp) -------------------------------------------------------
def envi_read_header(path):
    """HSIBinaryDataParameters::ReadHeaderFromFile (:226-270) over ConfigurationFileReader with '=' (config_reader.cpp:16-36)."""
    h = dict(interleave_bsq=1, data_type=4, big_endian=0, header_offset=0, num_data_rows=0, num_data_cols=0, num_data_bands=0)
    for line in open(path, "r", errors="replace").read().split("\n"):
        if line.startswith("#") or "=" not in line:
            continue
        key, value = [t.strip() for t in line.split("=", 1)]
        if key == "interleave":
            h["interleave_bsq"] = 1 if value == "bsq" else 0
        elif key == "data type":
            h["data_type"] = _atoi(value)
        elif key == "byte order":
            h["big_endian"] = 1 if value == "1" else 0
        elif key == "header offset":
            h["header_offset"] = _atoi(value)
        elif key == "samples":
            h["num_data_rows"] = _atoi(value)
        elif key == "lines":
            h["num_data_cols"] = _atoi(value)
        elif key == "bands":
            h["num_data_bands"] = _atoi(value)
    return h


def _atoi(s):
    s = s.strip()
    n = 0
    while n < len(s) and (s[n].isdigit() or (n == 0 and s[n] in "+-")):
        n += 1
    try:
        return int(s[:n])
    except ValueError:
        return 0


def envi_read(path, rows, cols, bands, big_endian, r, c, b):
    """ReadBinaryFileBSQ<float> (:68-118) for header_offset 0: element (band, row, col) is float number
    band * rows * cols + row * cols + col of the file, byte-reversed when the file's endianness differs."""
    raw = np.fromfile(path, dtype=">f4" if big_endian else "<f4", count=rows * cols * bands).reshape(bands, rows, cols)
    return raw[b[0]:b[1], r[0]:r[1], c[0]:c[1]].astype(np.float64)


# ---- N4: SpectralPCA (spectral_pca.cpp) -----------------------------------------------------------------------
def pca_input_data(images):
    """GetPCAInputData (:27-96): 10 * C samples in all, every (P // per_image)-th pixel of each image."""
    images = [np.asarray(im, dtype=np.float64).reshape(im.shape[0], -1) for im in images]
    Cn, P = images[0].shape
    per_image = min((Cn * 10) // len(images), P)
    skip = P // per_image
    rows = [im[:, np.arange(per_image) * skip].T for im in images]
    return np.concatenate(rows, axis=0)


def pca_train(images, num_pca_bands=0, retained_variance=0.0):
    """cv::PCA(data, noArray(), DATA_AS_ROW, k | retainedVariance): (mean [C], eigenvectors [k][C], eigenvalues [k])."""
    data = pca_input_data(images)
    mean = data.mean(axis=0)
    d = data - mean
    cov = d.T @ d / float(data.shape[0])
    w, v = np.linalg.eigh(cov)
    order = np.argsort(-w, kind="stable")
    w, v = w[order], v[:, order].T
    k = num_pca_bands
    if k == 0:
        # OpenCV's computeCumulativeEnergy (pca.cpp): L = the first index whose cumulative energy EXCEEDS the
        # fraction (the component that crosses it is not kept), then max(2, L)
        cum = np.cumsum(w) / np.sum(w)
        above = np.nonzero(cum > retained_variance)[0]
        k = int(above[0]) if above.size else len(w)
        k = min(max(k, 2), len(w))
    return mean, v[:k], w[:k]


def pca_project(mean, eigenvectors, image):
    x = np.asarray(image, dtype=np.float64)
    flat = x.reshape(x.shape[0], -1)
    return (eigenvectors @ (flat - mean[:, None])).reshape((eigenvectors.shape[0],) + x.shape[1:])


def pca_reconstruct(mean, eigenvectors, pca_image):
    y = np.asarray(pca_image, dtype=np.float64)
    flat = y.reshape(y.shape[0], -1)
    return (eigenvectors.T @ flat + mean[:, None]).reshape((eigenvectors.shape[1],) + y.shape[1:])


def float_bits(v):
    return struct.unpack("<I", struct.pack("<f", v))[0]
