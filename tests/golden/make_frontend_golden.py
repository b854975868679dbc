#!/usr/bin/env python
"""Generates tests/golden/frontend_fixtures.npz: outputs of the OpenCV entry points the reference calls for the
steps either side of the hot path, produced by the same functions in Python cv2 (run in
the build container, cv2 4.13.0; the fixtures are committed because cv2's behaviour is the only ground truth for
the un-vendored OpenCV dependency).

    python tests/golden/make_frontend_golden.py
"""
import os

import cv2
import numpy as np

out = {}
rng = np.random.default_rng(2024)
# cv::resize(INTER_LINEAR) as ImageData::ResizeImage(scale, INTERPOLATE_LINEAR) calls it (image_data.cpp:310-364)
for name, (h, w, s) in {"s2": (13, 17, 2), "s3": (9, 11, 3), "s4": (12, 10, 4), "s4_big": (40, 56, 4)}.items():
    src = rng.random((2, h, w))
    dst = np.stack([cv2.resize(src[c], (int(w * s), int(h * s)), interpolation=cv2.INTER_LINEAR) for c in range(2)])
    out["resize_%s_src" % name] = src
    out["resize_%s_dst" % name] = dst
# cv::PCA (DATA_AS_ROW) as SpectralPCA builds it (spectral_pca.cpp:155-173), on pixel vectors with a decaying spectrum
Cn, n = 12, 120
basis = np.linalg.qr(rng.standard_normal((Cn, Cn)))[0]
data = (rng.standard_normal((n, Cn)) * (2.0 ** -np.arange(Cn))) @ basis + rng.random(Cn)
mean, evec, eval_ = cv2.PCACompute2(data, mean=None, maxComponents=5)
out["pca_data"] = data
out["pca_mean"] = mean.reshape(-1)
out["pca_eigenvectors"] = evec
out["pca_eigenvalues"] = eval_.reshape(-1)
probe = rng.random((7, Cn))
proj = cv2.PCAProject(probe, mean, evec)
out["pca_probe"] = probe
out["pca_projected"] = proj
out["pca_backprojected"] = cv2.PCABackProject(proj, mean, evec)
mean_rv, evec_rv, eval_rv = cv2.PCACompute2(data, mean=None, retainedVariance=0.97)
out["pca_rv_components"] = np.array([evec_rv.shape[0]])
out["cv2_version"] = np.array(cv2.__version__)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "frontend_fixtures.npz"), **out)
print("wrote frontend_fixtures.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})
