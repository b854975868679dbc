"""ctypes binding of oracle/_ref/libsr_ref.so -- the REFERENCE'S OWN sources (ALGLIB 3.10.0, TV/BTV
regularizers, ObjectiveFunction, IRLS term, IRLSMapSolver) compiled unmodified from /root/reference,
with the OpenCV-dependent data term supplied by the oracle restatement or by a plugged-in callback.
TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/ref_shim.cpp).

`oracle/_ref/` is built in the build container (`make -C oracle ref`), is git-ignored and travels
to the GPU box prebuilt; nothing here reads /root/reference at run time.
"""
import ctypes as C
import os

import numpy as np

from . import sr_oracle as _o  # noqa: F401  (also makes sure libsr_oracle.so is built)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libsr_ref.so")
_FUSED_LIB_PATH = os.path.join(_HERE, "_ref", "libsr_ref_fused.so")
_dp = C.POINTER(C.c_double)

DATA_TERM_CB = C.CFUNCTYPE(C.c_double, _dp, _dp, C.c_int, C.c_int, C.c_void_p)
REG_APPLY_CB = C.CFUNCTYPE(None, _dp, C.c_int, _dp, C.c_void_p)
REG_APPLY_DIFF_CB = C.CFUNCTYPE(None, _dp, _dp, C.c_int, _dp, _dp, C.c_void_p)


class Callbacks(C.Structure):
    _fields_ = [("data_term", DATA_TERM_CB), ("reg_apply", REG_APPLY_CB),
                ("reg_apply_diff", REG_APPLY_DIFF_CB), ("user", C.c_void_p)]


class Options(C.Structure):
    _fields_ = [("solver", C.c_int), ("max_num_solver_iterations", C.c_int),
                ("max_num_irls_iterations", C.c_int), ("gradient_norm_threshold", C.c_double),
                ("cost_decrease_threshold", C.c_double),
                ("parameter_variation_threshold", C.c_double),
                ("irls_cost_difference_threshold", C.c_double), ("split_channels", C.c_int),
                ("num_lbfgs_hessian_corrections", C.c_int),
                ("use_numerical_differentiation", C.c_int),
                ("numerical_differentiation_step", C.c_double), ("num_threads", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("num_data_term_evals", C.c_long), ("seconds_in_data_term", C.c_double),
                ("seconds_total", C.c_double)]


def available():
    return os.path.exists(_LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _o.lib()
        L = C.CDLL(_LIB_PATH)
        L.ref_reg_apply.argtypes = [C.c_int, C.c_int, C.c_double, _dp, C.c_int, C.c_int, C.c_int,
                                    _dp]
        L.ref_reg_apply_diff.argtypes = [C.c_int, C.c_int, C.c_double, _dp, _dp, C.c_int, C.c_int,
                                         C.c_int, _dp, _dp]
        L.ref_irls_term.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, C.c_int,
                                    C.c_int, C.c_int, _dp]
        L.ref_irls_term.restype = C.c_double
        L.ref_compute_all_terms.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_int, _dp,
                                            C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp,
                                            C.c_int]
        L.ref_compute_all_terms.restype = C.c_double
        L.ref_default_options.argtypes = [C.POINTER(Options)]
        L.ref_solve.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int,
                                C.c_int, C.c_double, C.c_double, C.POINTER(Options),
                                C.POINTER(Callbacks), _dp, C.POINTER(Stats)]
        L.ref_solve.restype = C.c_int
        _lib = L
    return _lib


_fused = None


def fused_lib():
    """oracle/_ref/libsr_ref_fused.so: the reference's solver with the product's C++ adapters plugged in.
    The only oracle library that links libsrb200.so; the CPU reference arm never loads it."""
    global _fused
    if _fused is None:
        lib()
        L = C.CDLL(_FUSED_LIB_PATH)
        L.ref_solve_fused.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, C.c_int, C.c_double,
                                      C.POINTER(Options), _dp, C.POINTER(Stats)]
        L.ref_solve_fused.restype = C.c_int
        L.ref_solve_fused_multi.argtypes = L.ref_solve_fused.argtypes
        L.ref_solve_fused_multi.restype = C.c_int
        L.ref_solve_adapters.argtypes = [C.c_void_p, C.c_void_p, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp,
                                         C.c_int, C.c_double, C.POINTER(Options), _dp, C.POINTER(Stats)]
        L.ref_solve_adapters.restype = C.c_int
        _fused = L
    return _fused


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def reg_apply(kind, x, btv_range=3, btv_decay=0.5):
    x = _f64(x)
    Cn, H, W = x.shape
    out = np.empty_like(x)
    lib().ref_reg_apply(kind, btv_range, btv_decay, _p(x), H, W, Cn, _p(out))
    return out


def reg_apply_diff(kind, x, constants, btv_range=3, btv_decay=0.5):
    x, constants = _f64(x), _f64(constants)
    Cn, H, W = x.shape
    v, p = np.empty_like(x), np.empty_like(x)
    lib().ref_reg_apply_diff(kind, btv_range, btv_decay, _p(x), _p(constants), H, W, Cn, _p(v),
                             _p(p))
    return v, p


def irls_term(kind, lam, weights, x, grad=None, btv_range=3, btv_decay=0.5):
    x, weights = _f64(x), _f64(weights)
    Cn, H, W = x.shape
    return lib().ref_irls_term(kind, btv_range, btv_decay, lam, _p(weights), _p(x), H, W, Cn,
                               _p(grad))


def compute_all_terms(model, x, obs_hr, reg_kind=0, lam=0.0, weights=None, want_grad=True,
                      btv_range=3, btv_decay=0.5, threads=1):
    x, obs_hr = _f64(x), _f64(obs_hr)
    Cn, H, W = x.shape
    g = np.empty_like(x) if want_grad else None
    w = None if weights is None else _f64(weights)
    f = lib().ref_compute_all_terms(C.cast(model.c, C.c_void_p), _p(x), H, W, Cn, _p(obs_hr),
                                    reg_kind, btv_range, btv_decay, lam, _p(w), _p(g), threads)
    return f, g


def default_options():
    o = Options()
    lib().ref_default_options(C.byref(o))
    return o


def solve(model, lr, x0, reg_kind=-1, lam=0.0, btv_range=3, btv_decay=0.5, options=None,
          callbacks=None):
    """Reference IRLSMapSolver::Solve.  lr [N][C][h][w], x0 [C][H][W] -> (result, Stats)."""
    lr, x0 = _f64(lr), _f64(x0)
    N, Cn, h, w = lr.shape
    out = np.empty_like(x0)
    st = Stats()
    opt = options if options is not None else default_options()
    rc = lib().ref_solve(C.cast(model.c, C.c_void_p), _p(lr), N, Cn, h, w, _p(x0), reg_kind,
                         btv_range, btv_decay, lam, C.byref(opt),
                         C.byref(callbacks) if callbacks is not None else None, _p(out),
                         C.byref(st))
    assert rc == 0
    return out, st


def solve_fused(engine, x0, has_regularizer, lambda_sum, options=None):
    """The reference's IRLS + ALGLIB loop with the B200 engine plugged in through the C++ adapters
    of include/srb200_adapters.hpp (ref_shim.cpp: ref_solve_fused).  `engine` is a configured
    super-resolution_b200 Engine (model, observations, regularizer).  Returns (result, Stats)."""
    x0 = _f64(x0)
    Cn, H, W = x0.shape
    out = np.empty_like(x0)
    st = Stats()
    opt = options if options is not None else default_options()
    rc = fused_lib().ref_solve_fused(engine._ctx, Cn, H, W, _p(x0), 1 if has_regularizer else 0,
                               float(lambda_sum), C.byref(opt), _p(out), C.byref(st))
    assert rc == 0
    return out, st


def solve_fused_multi(multi_engine, x0, has_regularizer, lambda_sum, options=None):
    """solve_fused with ONE host thread driving several GPUs (CudaMultiObjectiveTerm -> srb_multi_eval).
    `multi_engine` is a configured super-resolution_b200 MultiEngine."""
    x0 = _f64(x0)
    Cn, H, W = x0.shape
    out = np.empty_like(x0)
    st = Stats()
    opt = options if options is not None else default_options()
    rc = fused_lib().ref_solve_fused_multi(multi_engine._ctx, Cn, H, W, _p(x0), 1 if has_regularizer else 0,
                                           float(lambda_sum), C.byref(opt), _p(out), C.byref(st))
    assert rc == 0
    return out, st


def solve_adapters(engine, model, lr, x0, has_regularizer, lam, options=None):
    """The reference's UNMODIFIED IRLSMapSolver::Solve with CudaObjectiveDataTerm and CudaRegularizer
    (include/srb200_adapters.hpp) instantiated in C++ (ref_fused.cpp: ref_solve_adapters).  `engine` holds the
    model, the observations and the regularizer kind.  Returns (result, Stats)."""
    lr, x0 = _f64(lr), _f64(x0)
    N, Cn, h, w = lr.shape
    out = np.empty_like(x0)
    st = Stats()
    opt = options if options is not None else default_options()
    rc = fused_lib().ref_solve_adapters(engine._ctx, C.cast(model.c, C.c_void_p), _p(lr), N, Cn, h, w, _p(x0),
                                        1 if has_regularizer else 0, float(lam), C.byref(opt), _p(out), C.byref(st))
    assert rc == 0
    return out, st
