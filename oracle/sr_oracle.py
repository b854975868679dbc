"""ctypes binding of the CPU ORACLE (oracle/sr_oracle.c) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (super-resolution_b200/) never does.

Parity status: pinned (tests/test_oracle_golden.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsr_oracle.so")
_dp = C.POINTER(C.c_double)

REG_TV, REG_TV3D, REG_BTV = 0, 1, 2


class _Model(C.Structure):
    _fields_ = [("scale", C.c_int), ("psf_size", C.c_int), ("psf", _dp),
                ("num_frames", C.c_int), ("shifts", _dp)]


def build(force=False):
    """Compile oracle/_build/libsr_oracle.so (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "_build/libsr_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.sro_gaussian_kernel.argtypes = [C.c_int, C.c_double, _dp]
        L.sro_gaussian_psf.argtypes = [C.c_int, C.c_double, _dp]
        L.sro_warp_shift.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_double, _dp]
        L.sro_warp_quantize.argtypes = [C.c_double]
        L.sro_warp_quantize.restype = C.c_int
        L.sro_filter2d.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int, _dp]
        L.sro_resize_nearest.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int]
        L.sro_nearest_index.argtypes = [C.c_int, C.c_int, C.c_int]
        L.sro_nearest_index.restype = C.c_int
        L.sro_resize_additive.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int]
        L.sro_forward.argtypes = [C.POINTER(_Model), C.c_int, _dp, C.c_int, C.c_int, _dp]
        L.sro_transpose.argtypes = [C.POINTER(_Model), C.c_int, _dp, C.c_int, C.c_int, _dp]
        L.sro_data_term.argtypes = [C.POINTER(_Model), _dp, C.c_int, C.c_int, C.c_int, _dp,
                                    C.c_int, C.c_int, _dp, C.c_int]
        L.sro_data_term.restype = C.c_double
        L.sro_reg_apply.argtypes = [C.c_int, C.c_int, C.c_double, _dp, C.c_int, C.c_int, C.c_int,
                                    _dp]
        L.sro_reg_apply_diff.argtypes = [C.c_int, C.c_int, C.c_double, _dp, _dp, C.c_int, C.c_int,
                                         C.c_int, _dp, _dp]
        L.sro_irls_term.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, C.c_int,
                                    C.c_int, C.c_int, _dp]
        L.sro_irls_term.restype = C.c_double
        L.sro_reweight.argtypes = [C.c_int, C.c_int, C.c_double, _dp, C.c_int, C.c_int, C.c_int,
                                   _dp]
        L.sro_eval.argtypes = [C.POINTER(_Model), _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int,
                               C.c_int, C.c_double, C.c_double, _dp, _dp, C.c_int]
        L.sro_eval.restype = C.c_double
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Model:
    """Oracle-side description of A_k = D B M_k (image_model.cpp:17-61)."""

    def __init__(self, scale, psf=None, shifts=None, num_frames=None):
        self.scale = int(scale)
        self.psf = None if psf is None else _f64(psf)
        self.shifts = None if shifts is None else _f64(shifts).reshape(-1, 2)
        if num_frames is None:
            num_frames = 0 if self.shifts is None else len(self.shifts)
        self.num_frames = int(num_frames)
        self._c = _Model(self.scale, 0 if self.psf is None else self.psf.shape[0], _p(self.psf),
                         self.num_frames, _p(self.shifts))

    @property
    def c(self):
        return C.byref(self._c)


def gaussian_kernel(n, sigma):
    out = np.empty(n)
    lib().sro_gaussian_kernel(n, sigma, _p(out))
    return out


def gaussian_psf(n, sigma):
    out = np.empty((n, n))
    lib().sro_gaussian_psf(n, sigma, _p(out))
    return out


def warp_quantize(d):
    return lib().sro_warp_quantize(float(d))


def warp_shift(img, dx, dy):
    img = _f64(img)
    out = np.empty_like(img)
    lib().sro_warp_shift(_p(img), img.shape[0], img.shape[1], dx, dy, _p(out))
    return out


def filter2d(img, kernel):
    img, kernel = _f64(img), _f64(kernel)
    out = np.empty_like(img)
    lib().sro_filter2d(_p(img), img.shape[0], img.shape[1], _p(kernel), kernel.shape[0],
                       kernel.shape[1], _p(out))
    return out


def nearest_index(q, n_src, n_dst):
    return lib().sro_nearest_index(q, n_src, n_dst)


def resize_nearest(img, h2, w2):
    img = _f64(img)
    out = np.empty((h2, w2))
    lib().sro_resize_nearest(_p(img), img.shape[0], img.shape[1], _p(out), h2, w2)
    return out


def resize_additive(img, h2, w2):
    img = _f64(img)
    out = np.empty((h2, w2))
    lib().sro_resize_additive(_p(img), img.shape[0], img.shape[1], _p(out), h2, w2)
    return out


def lr_size(scale, H, W):
    f = 1.0 / float(scale)
    return int(H * f), int(W * f)


def forward(model, k, hr):
    """hr: [H][W] -> [h][w] (one channel)."""
    hr = _f64(hr)
    H, W = hr.shape
    h, w = lr_size(model.scale, H, W)
    out = np.empty((h, w))
    lib().sro_forward(model.c, k, _p(hr), H, W, _p(out))
    return out


def transpose(model, k, lr):
    lr = _f64(lr)
    h, w = lr.shape
    out = np.empty((h * model.scale, w * model.scale))
    lib().sro_transpose(model.c, k, _p(lr), h, w, _p(out))
    return out


def upsample_observations(model, lr_stack):
    """MapSolver constructor (map_solver.cpp:81-85): [N][C][h][w] -> [N][C][H][W] nearest."""
    lr_stack = _f64(lr_stack)
    N, Cn, h, w = lr_stack.shape
    H, W = h * model.scale, w * model.scale
    out = np.empty((N, Cn, H, W))
    for k in range(N):
        for c in range(Cn):
            out[k, c] = resize_nearest(lr_stack[k, c], H, W)
    return out


def data_term(model, x, obs_hr, want_grad=True, grad=None, channel_start=0, threads=1):
    """x: [C][H][W]; obs_hr: [N][C_total][H][W].  Returns (cost, grad) with grad accumulated."""
    x, obs_hr = _f64(x), _f64(obs_hr)
    Cn, H, W = x.shape
    if want_grad and grad is None:
        grad = np.zeros_like(x)
    cost = lib().sro_data_term(model.c, _p(x), H, W, Cn, _p(obs_hr), obs_hr.shape[1],
                               channel_start, _p(grad) if want_grad else None, threads)
    return cost, grad


def reg_apply(kind, x, btv_range=3, btv_decay=0.5):
    x = _f64(x)
    Cn, H, W = x.shape
    out = np.empty_like(x)
    lib().sro_reg_apply(kind, btv_range, btv_decay, _p(x), H, W, Cn, _p(out))
    return out


def reg_apply_diff(kind, x, constants, btv_range=3, btv_decay=0.5):
    x, constants = _f64(x), _f64(constants)
    Cn, H, W = x.shape
    values, partials = np.empty_like(x), np.empty_like(x)
    lib().sro_reg_apply_diff(kind, btv_range, btv_decay, _p(x), _p(constants), H, W, Cn,
                             _p(values), _p(partials))
    return values, partials


def irls_term(kind, lam, weights, x, grad=None, btv_range=3, btv_decay=0.5):
    x, weights = _f64(x), _f64(weights)
    Cn, H, W = x.shape
    return lib().sro_irls_term(kind, btv_range, btv_decay, lam, _p(weights), _p(x), H, W, Cn,
                               _p(grad))


def reweight(kind, x, btv_range=3, btv_decay=0.5):
    x = _f64(x)
    Cn, H, W = x.shape
    out = np.empty_like(x)
    lib().sro_reweight(kind, btv_range, btv_decay, _p(x), H, W, Cn, _p(out))
    return out


def evaluate(model, x, obs_hr, reg_kind=REG_TV, lam=0.0, weights=None, want_grad=True,
             btv_range=3, btv_decay=0.5, threads=1):
    """ObjectiveFunction::ComputeAllTerms.  Returns (cost, grad or None)."""
    x, obs_hr = _f64(x), _f64(obs_hr)
    Cn, H, W = x.shape
    grad = np.empty_like(x) if want_grad else None
    # IRLS weights start at 1 (irls_map_solver.cpp:66-74): a regularizer without explicit weights
    # means the first IRLS round, not "no regularizer"
    w = (np.ones_like(x) if lam > 0.0 else None) if weights is None else _f64(weights)
    cost = lib().sro_eval(model.c, _p(x), H, W, Cn, _p(obs_hr), reg_kind, btv_range, btv_decay,
                          lam, _p(w), _p(grad), threads)
    return cost, grad
