"""ctypes binding of libsrb200.so (include/srb200.h).

`Engine` is a thin, typed wrapper over one srb_ctx; every method maps 1:1 to a C-ABI entry point
and raises `SrbError` on a non-zero status (the reference aborts through glog CHECK on the same
conditions).  There is no CPU fallback: constructing an Engine without a CUDA device fails.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_dp = C.POINTER(C.c_double)
_ctx_p = C.c_void_p

REG_NONE, REG_TV, REG_TV3D, REG_BTV = -1, 0, 1, 2
PATH_AUTO, PATH_REFERENCE_ORDER, PATH_FUSED = 0, 1, 2
PARTITION_FRAMES, PARTITION_ROWS = 0, 1
_STATUS = {0: "SRB_OK", 1: "SRB_ERR_INVALID", 2: "SRB_ERR_CUDA", 3: "SRB_ERR_GEOMETRY",
           4: "SRB_ERR_STATE", 5: "SRB_ERR_NOMEM"}

# name -> (restype, argtypes); also the export list tests/test_cabi.py checks against srb200.h
SIGNATURES = {
    "srb_version": (C.c_char_p, []),
    "srb_device_count": (C.c_int, []),
    "srb_plan": (C.c_int, [C.c_void_p, C.c_void_p]),
    "srb_quantize_shift": (C.c_int, [C.c_double]),
    "srb_sample_is_special": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "srb_create": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(_ctx_p)]),
    "srb_create_shard": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(_ctx_p)]),
    "srb_destroy": (None, [_ctx_p]),
    "srb_last_error": (C.c_char_p, [_ctx_p]),
    "srb_set_observations": (C.c_int, [_ctx_p, C.c_void_p]),
    "srb_set_observations_dev": (C.c_int, [_ctx_p, C.c_void_p]),
    "srb_set_channel_range": (C.c_int, [_ctx_p, C.c_int, C.c_int]),
    "srb_set_regularizer": (C.c_int, [_ctx_p, C.c_int, C.c_double, C.c_int, C.c_double]),
    "srb_set_irls_weights": (C.c_int, [_ctx_p, C.c_void_p]),
    "srb_reweight": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p]),
    "srb_reweight_dev": (C.c_int, [_ctx_p, C.c_void_p]),
    "srb_cg_minimize": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "srb_cg_minimize_dev": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "srb_lbfgs_minimize": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "srb_lbfgs_minimize_dev": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "srb_solve_irls": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p]),
    "srb_set_path": (C.c_int, [_ctx_p, C.c_int]),
    "srb_active_path": (C.c_int, [_ctx_p]),
    "srb_set_strict_cost": (C.c_int, [_ctx_p, C.c_int]),
    "srb_zlayout_active": (C.c_int, [_ctx_p]),
    "srb_set_regularizer_rows": (C.c_int, [_ctx_p, C.c_int, C.c_int]),
    "srb_eval": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, _dp]),
    "srb_eval_dev": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, _dp]),
    "srb_eval_partial_dev": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p]),
    "srb_num_units": (C.c_int, [_ctx_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "srb_unit_range": (C.c_int, [_ctx_p, C.c_int, C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "srb_eval_units_dev": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "srb_eval_finish_dev": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p]),
    "srb_eval_unit_range_dev": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "srb_halo_rows": (C.c_int, [_ctx_p]),
    "srb_set_profiling": (C.c_int, [_ctx_p, C.c_int]),
    "srb_peer_sizes": (C.c_int, [_ctx_p, C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "srb_dev_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_ulonglong]),
    "srb_dev_free": (C.c_int, [C.c_void_p]),
    "srb_ipc_export": (C.c_int, [C.c_void_p, C.c_char_p]),
    "srb_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "srb_ipc_close": (C.c_int, [C.c_void_p]),
    "srb_peer_setup": (C.c_int, [_ctx_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "srb_peer_scatter_dev": (C.c_int, [_ctx_p, C.c_void_p]),
    "srb_peer_gather_dev": (C.c_int, [_ctx_p]),
    "srb_peer_status": (C.c_int, [_ctx_p]),
    "srb_memcpy_d2h": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_ulonglong]),
    "srb_data_term": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, _dp]),
    "srb_irls_term": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, _dp]),
    "srb_reg_apply": (C.c_int, [_ctx_p, C.c_void_p, C.c_int, C.c_void_p]),
    "srb_reg_apply_diff": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "srb_forward": (C.c_int, [_ctx_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "srb_forward_all": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p]),
    "srb_transpose": (C.c_int, [_ctx_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "srb_pin_host": (C.c_int, [C.c_void_p, C.c_ulonglong]),
    "srb_unpin_host": (C.c_int, [C.c_void_p]),
    "srb_stream": (C.c_void_p, [_ctx_p]),
    "srb_dev_x": (C.c_void_p, [_ctx_p]),
    "srb_dev_gradient": (C.c_void_p, [_ctx_p]),
    "srb_synchronize": (C.c_int, [_ctx_p]),
    "srb_get_timing": (C.c_int, [_ctx_p, C.c_void_p]),
    # steps either side of the hot path
    "srb_resize_linear": (C.c_int, [_ctx_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "srb_initial_estimate": (C.c_int, [_ctx_p, C.c_int, C.c_void_p]),
    "srb_initial_estimate_dev": (C.c_int, [_ctx_p, C.c_int, C.c_void_p]),
    "srb_score": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_double, C.c_double, C.c_double, _dp, _dp]),
    "srb_score_dev": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_double, C.c_double, C.c_double, _dp, _dp]),
    "srb_add_noise": (C.c_int, [_ctx_p, C.c_void_p, C.c_ulonglong, C.c_double, C.c_ulonglong, C.c_ulonglong]),
    "srb_add_noise_dev": (C.c_int, [_ctx_p, C.c_void_p, C.c_ulonglong, C.c_double, C.c_ulonglong, C.c_ulonglong]),
    "srb_generate_observations": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_double, C.c_ulonglong, C.c_void_p, C.c_int]),
    "srb_envi_read_header": (C.c_int, [C.c_char_p, C.c_void_p]),
    "srb_envi_read": (C.c_int, [_ctx_p, C.c_char_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "srb_envi_read_dev": (C.c_int, [_ctx_p, C.c_char_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "srb_envi_write": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "srb_pca_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_ulonglong, C.c_int, C.c_double, C.POINTER(_ctx_p)]),
    "srb_pca_destroy": (None, [_ctx_p]),
    "srb_pca_num_components": (C.c_int, [_ctx_p]),
    "srb_pca_num_bands": (C.c_int, [_ctx_p]),
    "srb_pca_get": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "srb_pca_project": (C.c_int, [_ctx_p, _ctx_p, C.c_void_p, C.c_ulonglong, C.c_void_p]),
    "srb_pca_reconstruct": (C.c_int, [_ctx_p, _ctx_p, C.c_void_p, C.c_ulonglong, C.c_void_p]),
    # single-process multi-GPU form
    "srb_multi_create": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(_ctx_p)]),
    "srb_multi_create_partitioned": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(_ctx_p)]),
    "srb_multi_destroy": (None, [_ctx_p]),
    "srb_multi_last_error": (C.c_char_p, [_ctx_p]),
    "srb_multi_num_gpus": (C.c_int, [_ctx_p]),
    "srb_multi_rank_ctx": (_ctx_p, [_ctx_p, C.c_int]),
    "srb_multi_set_observations": (C.c_int, [_ctx_p, C.c_void_p]),
    "srb_multi_set_channel_range": (C.c_int, [_ctx_p, C.c_int, C.c_int]),
    "srb_multi_set_regularizer": (C.c_int, [_ctx_p, C.c_int, C.c_double, C.c_int, C.c_double]),
    "srb_multi_set_irls_weights": (C.c_int, [_ctx_p, C.c_void_p]),
    "srb_multi_reweight": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p]),
    "srb_multi_set_path": (C.c_int, [_ctx_p, C.c_int]),
    "srb_multi_eval": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, _dp]),
    "srb_multi_cg_minimize": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "srb_multi_lbfgs_minimize": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "srb_multi_solve_irls": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p]),
    "srb_multi_get_timing": (C.c_int, [_ctx_p, C.c_void_p]),
}


class ModelDesc(C.Structure):
    _fields_ = [("lr_height", C.c_int), ("lr_width", C.c_int), ("num_channels", C.c_int),
                ("num_frames", C.c_int), ("scale", C.c_int), ("psf_size", C.c_int),
                ("psf", _dp), ("shifts", _dp)]


class EnviHeader(C.Structure):
    """srb_envi_header (HSIBinaryDataParameters, hyperspectral_data_loader.h)."""
    _fields_ = [("interleave_bsq", C.c_int), ("data_type", C.c_int), ("big_endian", C.c_int),
                ("header_offset", C.c_int), ("num_data_rows", C.c_int), ("num_data_cols", C.c_int),
                ("num_data_bands", C.c_int)]


class Timing(C.Structure):
    _fields_ = [("last_eval_kernel_ms", C.c_double), ("last_eval_h2d_ms", C.c_double),
                ("last_eval_d2h_ms", C.c_double), ("num_evals", C.c_ulonglong),
                ("kernel_launches", C.c_ulonglong), ("algorithmic_bytes_per_eval", C.c_ulonglong),
                ("last_main_kernel_ms", C.c_double)]


class CgOptions(C.Structure):
    """srb_cg_options: the thresholds of mincgsetcond (alglib_objective.cpp:57-62)."""
    _fields_ = [("gradient_norm_threshold", C.c_double), ("cost_decrease_threshold", C.c_double),
                ("parameter_variation_threshold", C.c_double), ("max_num_solver_iterations", C.c_int),
                ("num_lbfgs_hessian_corrections", C.c_int)]


class CgReport(C.Structure):
    _fields_ = [("iterations", C.c_int), ("num_evaluations", C.c_int), ("termination_type", C.c_int),
                ("num_restarts", C.c_int), ("final_cost", C.c_double)]


class IrlsReport(C.Structure):
    _fields_ = [("num_irls_iterations", C.c_int), ("num_solver_iterations", C.c_int),
                ("num_evaluations", C.c_int), ("last_termination_type", C.c_int), ("final_cost", C.c_double)]


def _as_dict(st):
    return {name: getattr(st, name) for name, _ in st._fields_}


class PlanInfo(C.Structure):
    _fields_ = [("hr_height", C.c_int), ("hr_width", C.c_int), ("warps_uniform", C.c_int),
                ("warps_integer", C.c_int), ("fused", C.c_int), ("fractional", C.c_int),
                ("psf_half", C.c_int), ("num_entries", C.c_int), ("min_entries_per_phase", C.c_int),
                ("max_entries_per_phase", C.c_int), ("band_lo_r", C.c_int), ("band_hi_r", C.c_int),
                ("band_lo_c", C.c_int), ("band_hi_c", C.c_int), ("table_driven", C.c_int), ("zlayout", C.c_int),
                ("zt_frames", C.c_int), ("why", C.c_char * 160)]


class SrbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (_STATUS.get(status, status), message))
        self.status = status


_lib = None


def library_path():
    return _build.LIB


def load_library(rebuild=False):
    """Loads (building first if needed) libsrb200.so.  Raises if it cannot be built or loaded --
    the product never substitutes a CPU implementation."""
    global _lib
    if _lib is None or rebuild:
        path = _build.build(force=rebuild) if (rebuild or not os.path.exists(_build.LIB)) else _build.LIB
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def device_count():
    return load_library().srb_device_count()


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _host_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dev_ptr(t):
    """Device pointer of a torch CUDA tensor (float64, contiguous) or a raw int address."""
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    assert t.is_cuda and t.is_contiguous() and t.element_size() == 8
    return C.c_void_p(t.data_ptr())


class Engine:
    """One srb_ctx: the image formation model A_k = D B M_k plus the observations of this rank's
    frame shard, on one CUDA device."""

    def __init__(self, lr_shape, scale, psf=None, shifts=None, device=0):
        """lr_shape = (N, C, h, w); psf: K x K array or None; shifts: (N, 2) (dx, dy) or None."""
        self._lib = load_library()
        self._ctx = _ctx_p()
        N, Cn, h, w = (int(v) for v in lr_shape)
        self.N, self.C, self.h, self.w, self.scale = N, Cn, h, w, int(scale)
        self.H, self.W = h * self.scale, w * self.scale
        self.c0, self.c1 = 0, Cn
        self._psf = None if psf is None else _f64(psf)
        self._shifts = None if shifts is None else _f64(shifts).reshape(-1, 2)
        if self._shifts is not None and len(self._shifts) != N:
            raise SrbError(1, "need one (dx, dy) shift per frame")
        desc = ModelDesc(h, w, Cn, N, self.scale, 0 if self._psf is None else self._psf.shape[0],
                         None if self._psf is None else self._psf.ctypes.data_as(_dp),
                         None if self._shifts is None else self._shifts.ctypes.data_as(_dp))
        st = self._lib.srb_create(C.byref(desc), int(device), C.byref(self._ctx))
        if st != 0:
            msg = self._lib.srb_last_error(self._ctx).decode() if self._ctx else "allocation failed"
            if self._ctx:
                self._lib.srb_destroy(self._ctx)
            self._ctx = None
            raise SrbError(st, msg)

    # -- life cycle
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.srb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, st):
        if st != 0:
            raise SrbError(st, self._lib.srb_last_error(self._ctx).decode())

    @property
    def num_active(self):
        return (self.c1 - self.c0) * self.H * self.W

    # -- configuration
    def set_observations(self, lr):
        if hasattr(lr, "is_cuda"):
            assert tuple(lr.shape) == (self.N, self.C, self.h, self.w)
            self._check(self._lib.srb_set_observations_dev(self._ctx, _dev_ptr(lr)))
            return
        lr = _f64(lr)
        assert lr.shape == (self.N, self.C, self.h, self.w), lr.shape
        self._check(self._lib.srb_set_observations(self._ctx, _host_ptr(lr)))

    def set_channel_range(self, c0, c1):
        self._check(self._lib.srb_set_channel_range(self._ctx, int(c0), int(c1)))
        self.c0, self.c1 = int(c0), int(c1)

    def set_regularizer(self, kind, lam, btv_range=3, btv_decay=0.5):
        self._check(self._lib.srb_set_regularizer(self._ctx, int(kind), float(lam), int(btv_range),
                                                  float(btv_decay)))

    def set_irls_weights(self, weights):
        w = None if weights is None else _f64(weights).reshape(-1)
        if w is not None:
            assert w.size == self.num_active
        self._check(self._lib.srb_set_irls_weights(self._ctx, _host_ptr(w)))

    def reweight(self, x=None, want_weights=True):
        xa = None if x is None else _f64(x).reshape(-1)
        out = np.empty(self.num_active) if want_weights else None
        self._check(self._lib.srb_reweight(self._ctx, _host_ptr(xa), _host_ptr(out)))
        return None if out is None else out.reshape(self.c1 - self.c0, self.H, self.W)

    def reweight_dev(self, x_dev):
        self._check(self._lib.srb_reweight_dev(self._ctx, _dev_ptr(x_dev)))

    # -- device-resident solver
    @staticmethod
    def _cg_options(epsg, epsf, epsx, maxits, lbfgs_corrections=0):
        return CgOptions(float(epsg), float(epsf), float(epsx), int(maxits), int(lbfgs_corrections))

    def cg_minimize(self, x0, epsg=0.0, epsf=0.0, epsx=0.0, maxits=0):
        """RunCGSolverAnalyticalDiff with the solver vectors on the device.  Returns (x, report dict)."""
        x = np.array(_f64(x0).reshape(-1), copy=True)
        assert x.size == self.num_active
        opt, rep = self._cg_options(epsg, epsf, epsx, maxits), CgReport()
        self._check(self._lib.srb_cg_minimize(self._ctx, _host_ptr(x), C.byref(opt), C.byref(rep)))
        return x.reshape(self.c1 - self.c0, self.H, self.W), _as_dict(rep)

    def lbfgs_minimize(self, x0, corrections=5, epsg=0.0, epsf=0.0, epsx=0.0, maxits=0):
        """RunLBFGSSolverAnalyticalDiff with the solver vectors on the device.  Returns (x, report dict)."""
        x = np.array(_f64(x0).reshape(-1), copy=True)
        assert x.size == self.num_active
        opt, rep = self._cg_options(epsg, epsf, epsx, maxits, corrections), CgReport()
        self._check(self._lib.srb_lbfgs_minimize(self._ctx, _host_ptr(x), C.byref(opt), C.byref(rep)))
        return x.reshape(self.c1 - self.c0, self.H, self.W), _as_dict(rep)

    def cg_minimize_inplace(self, x, epsg=0.0, epsf=0.0, epsx=0.0, maxits=0):
        """srb_cg_minimize on the caller's own buffer (float64, contiguous, num_active elements; pin it with
        pin_host for full PCIe rate): the initial estimate on entry, the solution on return.  Returns the report."""
        assert isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous and x.size == self.num_active
        opt, rep = self._cg_options(epsg, epsf, epsx, maxits), CgReport()
        self._check(self._lib.srb_cg_minimize(self._ctx, _host_ptr(x), C.byref(opt), C.byref(rep)))
        return _as_dict(rep)

    def cg_minimize_dev(self, x_dev, epsg=0.0, epsf=0.0, epsx=0.0, maxits=0):
        opt, rep = self._cg_options(epsg, epsf, epsx, maxits), CgReport()
        self._check(self._lib.srb_cg_minimize_dev(self._ctx, _dev_ptr(x_dev), C.byref(opt), C.byref(rep)))
        return _as_dict(rep)

    def solve_irls(self, x0, epsg=0.0, epsf=0.0, epsx=0.0, maxits=0, max_irls_iterations=20,
                   irls_cost_difference_threshold=1.0e-5, lbfgs_corrections=0):
        """IRLSMapSolver::RunIRLSLoop on the device.  Returns (x, report dict).  The outer-loop defaults are
        the reference's (irls_map_solver.h:27,35: 20 iterations, cost difference 1e-5); unlimited iterations
        together with a zero threshold would never terminate and is refused by srb_solve_irls."""
        x = np.array(_f64(x0).reshape(-1), copy=True)
        assert x.size == self.num_active
        opt, rep = self._cg_options(epsg, epsf, epsx, maxits, lbfgs_corrections), IrlsReport()
        self._check(self._lib.srb_solve_irls(self._ctx, _host_ptr(x), C.byref(opt), int(max_irls_iterations),
                                             float(irls_cost_difference_threshold), C.byref(rep)))
        return x.reshape(self.c1 - self.c0, self.H, self.W), _as_dict(rep)

    def set_path(self, path):
        self._check(self._lib.srb_set_path(self._ctx, int(path)))

    def set_strict_cost(self, on=True):
        """Reference-order kernels sum their costs in the reference's own sequential order (bit-identical cost)."""
        self._check(self._lib.srb_set_strict_cost(self._ctx, 1 if on else 0))

    @property
    def active_path(self):
        return self._lib.srb_active_path(self._ctx)

    @property
    def zlayout_active(self):
        return bool(self._lib.srb_zlayout_active(self._ctx))

    def set_regularizer_rows(self, r0, r1):
        self._check(self._lib.srb_set_regularizer_rows(self._ctx, int(r0), int(r1)))

    # -- hot path
    def eval(self, x, want_grad=True, out=None):
        """ObjectiveFunction::ComputeAllTerms with host buffers.  Returns (cost, gradient|None)."""
        xa = _f64(x).reshape(-1)
        assert xa.size == self.num_active
        g = None
        if want_grad:
            g = out if out is not None else np.empty(self.num_active)
        cost = C.c_double()
        self._check(self._lib.srb_eval(self._ctx, _host_ptr(xa), _host_ptr(g), C.byref(cost)))
        return cost.value, (None if g is None else g.reshape(self.c1 - self.c0, self.H, self.W))

    def eval_dev(self, x_dev, g_dev=None, want_cost=True):
        cost = C.c_double()
        self._check(self._lib.srb_eval_dev(self._ctx, _dev_ptr(x_dev), _dev_ptr(g_dev),
                                           C.byref(cost) if want_cost else None))
        return cost.value if want_cost else None

    def eval_partial_dev(self, x_dev, gc_dev):
        self._check(self._lib.srb_eval_partial_dev(self._ctx, _dev_ptr(x_dev), _dev_ptr(gc_dev)))

    # -- pipelined multi-GPU form
    def num_units(self):
        """(number of units, HR rows per unit) -- see srb_num_units."""
        n, r = C.c_int(), C.c_int()
        self._check(self._lib.srb_num_units(self._ctx, C.byref(n), C.byref(r)))
        return n.value, r.value

    def unit_range(self, u0, u1):
        b, e = C.c_ulonglong(), C.c_ulonglong()
        self._check(self._lib.srb_unit_range(self._ctx, int(u0), int(u1), C.byref(b), C.byref(e)))
        return b.value, e.value

    def eval_units_dev(self, x_dev, gc_dev, u0, u1):
        self._check(self._lib.srb_eval_units_dev(self._ctx, _dev_ptr(x_dev), _dev_ptr(gc_dev), int(u0), int(u1)))

    def eval_finish_dev(self, x_dev, gc_dev):
        self._check(self._lib.srb_eval_finish_dev(self._ctx, _dev_ptr(x_dev), _dev_ptr(gc_dev)))

    def eval_unit_range_dev(self, x_dev, g_dev, u0, u1, cost_dev=None):
        """Units [u0, u1) of the whole objective + their cost (row-band partition, srb_eval_unit_range_dev)."""
        self._check(self._lib.srb_eval_unit_range_dev(self._ctx, _dev_ptr(x_dev), _dev_ptr(g_dev), int(u0), int(u1),
                                                      _dev_ptr(cost_dev)))

    def halo_rows(self):
        """HR rows of x either side of a gradient row band that its evaluation reads (srb_halo_rows)."""
        return self._lib.srb_halo_rows(self._ctx)

    def total_units(self):
        """(channel, 32-row tile) units of the active range, whether or not they can be pipelined."""
        return (self.c1 - self.c0) * ((self.H + 31) // 32)

    def set_profiling(self, on=True):
        self._check(self._lib.srb_set_profiling(self._ctx, 1 if on else 0))

    # -- multi-GPU peer path (reduce-scatter fused into the tile kernel, gather over NVLink)
    def peer_sizes(self, world):
        a, b = C.c_ulonglong(), C.c_ulonglong()
        self._check(self._lib.srb_peer_sizes(self._ctx, int(world), C.byref(a), C.byref(b)))
        return a.value, b.value

    def peer_setup(self, rank, world, slot_ptrs, out_ptrs):
        sa = (C.c_void_p * world)(*slot_ptrs)
        oa = (C.c_void_p * world)(*out_ptrs)
        self._check(self._lib.srb_peer_setup(self._ctx, int(rank), int(world), sa, oa))

    def peer_scatter_dev(self, x_dev):
        self._check(self._lib.srb_peer_scatter_dev(self._ctx, _dev_ptr(x_dev)))

    def peer_gather_dev(self):
        self._check(self._lib.srb_peer_gather_dev(self._ctx))

    def peer_status(self):
        """Raises SrbError(SRB_ERR_STATE) if a flag barrier of the peer path timed out since the last check."""
        self._check(self._lib.srb_peer_status(self._ctx))

    def memcpy_d2h(self, dst, src_ptr, nbytes):
        self._check(self._lib.srb_memcpy_d2h(self._ctx, dst.ctypes.data_as(C.c_void_p), C.c_void_p(src_ptr), int(nbytes)))

    def data_term(self, x, gradient=None):
        """ObjectiveDataTerm::Compute: returns cost; ADDS into `gradient` (in place) if given."""
        xa = _f64(x).reshape(-1)
        if gradient is not None:
            assert gradient.dtype == np.float64 and gradient.flags.c_contiguous
        cost = C.c_double()
        self._check(self._lib.srb_data_term(self._ctx, _host_ptr(xa), _host_ptr(gradient), C.byref(cost)))
        return cost.value

    def irls_term(self, x, gradient=None):
        xa = _f64(x).reshape(-1)
        if gradient is not None:
            assert gradient.dtype == np.float64 and gradient.flags.c_contiguous
        cost = C.c_double()
        self._check(self._lib.srb_irls_term(self._ctx, _host_ptr(xa), _host_ptr(gradient), C.byref(cost)))
        return cost.value

    def reg_apply(self, x):
        xa = _f64(x)
        Cn = xa.shape[0]
        out = np.empty_like(xa)
        self._check(self._lib.srb_reg_apply(self._ctx, _host_ptr(xa), Cn, _host_ptr(out)))
        return out

    def reg_apply_diff(self, x, constants):
        xa, ca = _f64(x), _f64(constants)
        Cn = xa.shape[0]
        v, p = np.empty_like(xa), np.empty_like(xa)
        self._check(self._lib.srb_reg_apply_diff(self._ctx, _host_ptr(xa), _host_ptr(ca), Cn,
                                                 _host_ptr(v), _host_ptr(p)))
        return v, p

    def forward(self, k, hr):
        hr = _f64(hr)
        H, W = hr.shape
        f = 1.0 / float(self.scale)
        out = np.empty((int(H * f), int(W * f)))
        self._check(self._lib.srb_forward(self._ctx, int(k), _host_ptr(hr), H, W, _host_ptr(out)))
        return out

    def forward_all(self, hr):
        """The whole LR stack [N][C][h][w] of an HR image [C][H][W] (ImageModel::ApplyToImage per frame)."""
        hr = _f64(hr)
        assert hr.shape == (self.C, self.H, self.W), hr.shape
        out = np.empty((self.N, self.C, self.h, self.w))
        self._check(self._lib.srb_forward_all(self._ctx, _host_ptr(hr), _host_ptr(out)))
        return out

    def transpose(self, k, lr):
        lr = _f64(lr)
        h, w = lr.shape
        out = np.empty((h * self.scale, w * self.scale))
        self._check(self._lib.srb_transpose(self._ctx, int(k), _host_ptr(lr), h, w, _host_ptr(out)))
        return out

    # -- the steps either side of the hot path
    def resize_linear(self, src, H, W):
        """ImageData::ResizeImage(size, INTERPOLATE_LINEAR): [C][h][w] -> [C][H][W]."""
        src = _f64(src)
        Cn, h, w = src.shape
        out = np.empty((Cn, int(H), int(W)))
        self._check(self._lib.srb_resize_linear(self._ctx, _host_ptr(src), Cn, h, w, int(H), int(W), _host_ptr(out)))
        return out

    def initial_estimate(self, frame=0):
        """Bilinear upsampling of LR observation `frame` of the active channel range (super_resolution.cpp:368-373)."""
        out = np.empty((self.c1 - self.c0, self.H, self.W))
        self._check(self._lib.srb_initial_estimate(self._ctx, int(frame), _host_ptr(out)))
        return out

    def initial_estimate_dev(self, x_dev, frame=0):
        self._check(self._lib.srb_initial_estimate_dev(self._ctx, int(frame), _dev_ptr(x_dev)))

    def score(self, image, truth, k1=0.01, k2=0.03, image_scale=1.0):
        """(PSNR, SSIM) of `image` against `truth` (src/evaluation); host arrays or torch CUDA tensors."""
        psnr, ssim = C.c_double(), C.c_double()
        if hasattr(image, "is_cuda"):
            n = image.numel()
            assert truth.numel() == n
            self._check(self._lib.srb_score_dev(self._ctx, _dev_ptr(image), _dev_ptr(truth), n, k1, k2, image_scale,
                                                C.byref(psnr), C.byref(ssim)))
        else:
            a, b = _f64(image).reshape(-1), _f64(truth).reshape(-1)
            assert a.size == b.size
            self._check(self._lib.srb_score(self._ctx, _host_ptr(a), _host_ptr(b), a.size, k1, k2, image_scale,
                                            C.byref(psnr), C.byref(ssim)))
        return psnr.value, ssim.value

    def add_noise(self, data, sigma, seed=0, stream_id=0):
        """AdditiveNoiseModule::ApplyToImage: returns data + N(0, (sigma/255)^2), Philox4x32-10 counter-based."""
        out = np.array(_f64(data), copy=True)
        self._check(self._lib.srb_add_noise(self._ctx, _host_ptr(out.reshape(-1)), out.size, float(sigma), int(seed),
                                            int(stream_id)))
        return out

    def generate_observations(self, hr, noise_sigma=0.0, seed=0, keep=True, want_lr=True):
        """The whole LR stack of an HR image through the image model (+ noise), optionally kept as the context's
        observations.  hr: host array [C][H][W] or torch CUDA tensor."""
        out = np.empty((self.N, self.C, self.h, self.w)) if want_lr else None
        if hasattr(hr, "is_cuda"):
            assert tuple(hr.shape) == (self.C, self.H, self.W)
            hp, dp = None, _dev_ptr(hr)
        else:
            hr = _f64(hr)
            assert hr.shape == (self.C, self.H, self.W), hr.shape
            hp, dp = _host_ptr(hr), None
        self._check(self._lib.srb_generate_observations(self._ctx, hp, dp, float(noise_sigma), int(seed), _host_ptr(out),
                                                        1 if keep else 0))
        return out

    def envi_read(self, data_path, header, rows=None, cols=None, bands=None):
        """ReadBinaryFileBSQ<float>: the selected range of a float32 BSQ file as [bands][rows][cols] doubles."""
        r0, r1 = rows if rows is not None else (0, header.num_data_rows)
        c0, c1 = cols if cols is not None else (0, header.num_data_cols)
        b0, b1 = bands if bands is not None else (0, header.num_data_bands)
        out = np.empty((max(b1 - b0, 0), max(r1 - r0, 0), max(c1 - c0, 0)))
        self._check(self._lib.srb_envi_read(self._ctx, os.fsencode(data_path), C.byref(header), int(r0), int(r1), int(c0),
                                            int(c1), int(b0), int(b1), _host_ptr(out)))
        return out

    def pca_project(self, pca, image):
        """SpectralPCA::GetPCAImage: [C][...] -> [k][...]."""
        image = _f64(image)
        assert image.shape[0] == pca.num_bands
        P = int(np.prod(image.shape[1:]))
        out = np.empty((pca.num_components,) + image.shape[1:])
        self._check(self._lib.srb_pca_project(self._ctx, pca._p, _host_ptr(image), P, _host_ptr(out)))
        return out

    def pca_reconstruct(self, pca, pca_image):
        """SpectralPCA::ReconstructImage: [k][...] -> [C][...]."""
        pca_image = _f64(pca_image)
        assert pca_image.shape[0] == pca.num_components
        P = int(np.prod(pca_image.shape[1:]))
        out = np.empty((pca.num_bands,) + pca_image.shape[1:])
        self._check(self._lib.srb_pca_reconstruct(self._ctx, pca._p, _host_ptr(pca_image), P, _host_ptr(out)))
        return out

    # -- plumbing
    def stream_handle(self):
        return self._lib.srb_stream(self._ctx)

    def dev_x_ptr(self):
        return self._lib.srb_dev_x(self._ctx)

    def dev_gradient_ptr(self):
        return self._lib.srb_dev_gradient(self._ctx)

    def synchronize(self):
        self._check(self._lib.srb_synchronize(self._ctx))

    def timing(self):
        t = Timing()
        self._check(self._lib.srb_get_timing(self._ctx, C.byref(t)))
        return {name: getattr(t, name) for name, _ in Timing._fields_}


class MultiEngine:
    """One srb_multi: the whole model sharded over `n_gpus` devices of this process (frames in contiguous
    blocks, x and IRLS weights replicated, regularizer split by row bands) behind the call shape of
    `Engine.eval` -- what the reference's single-threaded solver would hold."""

    def __init__(self, lr_shape, scale, psf=None, shifts=None, n_gpus=1, devices=None, partition=None):
        """partition: PARTITION_FRAMES (the contract partition: frames sharded, gradient summed over NVLink),
        PARTITION_ROWS (every device evaluates the whole objective on its HR row bands, no exchange), or None =
        srb_multi_create's choice (SRB_MULTI_PARTITION in the environment, default frames)."""
        self._lib = load_library()
        self._ctx = _ctx_p()
        N, Cn, h, w = (int(v) for v in lr_shape)
        self.N, self.C, self.h, self.w, self.scale = N, Cn, h, w, int(scale)
        self.H, self.W = h * self.scale, w * self.scale
        self.c0, self.c1 = 0, Cn
        self.n_gpus = int(n_gpus)
        self._psf = None if psf is None else _f64(psf)
        self._shifts = None if shifts is None else _f64(shifts).reshape(-1, 2)
        desc = ModelDesc(h, w, Cn, N, self.scale, 0 if self._psf is None else self._psf.shape[0],
                         None if self._psf is None else self._psf.ctypes.data_as(_dp),
                         None if self._shifts is None else self._shifts.ctypes.data_as(_dp))
        dev = None if devices is None else (C.c_int * self.n_gpus)(*[int(d) for d in devices])
        if partition is None:
            st = self._lib.srb_multi_create(C.byref(desc), self.n_gpus, dev, C.byref(self._ctx))
        else:
            st = self._lib.srb_multi_create_partitioned(C.byref(desc), self.n_gpus, dev, int(partition), C.byref(self._ctx))
        if st != 0:
            msg = self._lib.srb_multi_last_error(self._ctx).decode() if self._ctx else "allocation failed"
            if self._ctx:
                self._lib.srb_multi_destroy(self._ctx)
            self._ctx = None
            raise SrbError(st, msg)

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.srb_multi_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, st):
        if st != 0:
            raise SrbError(st, self._lib.srb_multi_last_error(self._ctx).decode())

    @property
    def num_active(self):
        return (self.c1 - self.c0) * self.H * self.W

    def set_observations(self, lr):
        lr = _f64(lr)
        assert lr.shape == (self.N, self.C, self.h, self.w), lr.shape
        self._check(self._lib.srb_multi_set_observations(self._ctx, _host_ptr(lr)))

    def set_channel_range(self, c0, c1):
        self._check(self._lib.srb_multi_set_channel_range(self._ctx, int(c0), int(c1)))
        self.c0, self.c1 = int(c0), int(c1)

    def set_regularizer(self, kind, lam, btv_range=3, btv_decay=0.5):
        self._check(self._lib.srb_multi_set_regularizer(self._ctx, int(kind), float(lam), int(btv_range),
                                                        float(btv_decay)))

    def set_irls_weights(self, weights):
        w = None if weights is None else _f64(weights).reshape(-1)
        if w is not None:
            assert w.size == self.num_active
        self._check(self._lib.srb_multi_set_irls_weights(self._ctx, _host_ptr(w)))

    def reweight(self, x, want_weights=True):
        xa = _f64(x).reshape(-1)
        out = np.empty(self.num_active) if want_weights else None
        self._check(self._lib.srb_multi_reweight(self._ctx, _host_ptr(xa), _host_ptr(out)))
        return None if out is None else out.reshape(self.c1 - self.c0, self.H, self.W)

    def set_path(self, path):
        self._check(self._lib.srb_multi_set_path(self._ctx, int(path)))

    def eval(self, x, want_grad=True, out=None):
        """ObjectiveFunction::ComputeAllTerms over all devices, host buffers.  Returns (cost, gradient|None)."""
        xa = _f64(x).reshape(-1)
        assert xa.size == self.num_active
        g = None
        if want_grad:
            g = out if out is not None else np.empty(self.num_active)
        cost = C.c_double()
        self._check(self._lib.srb_multi_eval(self._ctx, _host_ptr(xa), _host_ptr(g), C.byref(cost)))
        return cost.value, (None if g is None else g.reshape(self.c1 - self.c0, self.H, self.W))

    # -- device-resident solver on all devices (row-band partition)
    def cg_minimize_inplace(self, x, epsg=0.0, epsf=0.0, epsx=0.0, maxits=0):
        """srb_multi_cg_minimize on the caller's buffer (float64, contiguous, num_active elements): the initial
        estimate on entry, the solution on return.  Returns the report dict."""
        assert isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous and x.size == self.num_active
        opt, rep = Engine._cg_options(epsg, epsf, epsx, maxits), CgReport()
        self._check(self._lib.srb_multi_cg_minimize(self._ctx, _host_ptr(x), C.byref(opt), C.byref(rep)))
        return _as_dict(rep)

    def cg_minimize(self, x0, epsg=0.0, epsf=0.0, epsx=0.0, maxits=0):
        x = np.array(_f64(x0).reshape(-1), copy=True)
        rep = self.cg_minimize_inplace(x, epsg, epsf, epsx, maxits)
        return x.reshape(self.c1 - self.c0, self.H, self.W), rep

    def lbfgs_minimize(self, x0, corrections=5, epsg=0.0, epsf=0.0, epsx=0.0, maxits=0):
        x = np.array(_f64(x0).reshape(-1), copy=True)
        assert x.size == self.num_active
        opt, rep = Engine._cg_options(epsg, epsf, epsx, maxits, corrections), CgReport()
        self._check(self._lib.srb_multi_lbfgs_minimize(self._ctx, _host_ptr(x), C.byref(opt), C.byref(rep)))
        return x.reshape(self.c1 - self.c0, self.H, self.W), _as_dict(rep)

    def solve_irls(self, x0, epsg=0.0, epsf=0.0, epsx=0.0, maxits=0, max_irls_iterations=20,
                   irls_cost_difference_threshold=1.0e-5, lbfgs_corrections=0):
        """IRLSMapSolver::RunIRLSLoop on all devices (defaults as Engine.solve_irls).  Returns (x, report dict)."""
        x = np.array(_f64(x0).reshape(-1), copy=True)
        assert x.size == self.num_active
        opt, rep = Engine._cg_options(epsg, epsf, epsx, maxits, lbfgs_corrections), IrlsReport()
        self._check(self._lib.srb_multi_solve_irls(self._ctx, _host_ptr(x), C.byref(opt), int(max_irls_iterations),
                                                   float(irls_cost_difference_threshold), C.byref(rep)))
        return x.reshape(self.c1 - self.c0, self.H, self.W), _as_dict(rep)

    def timing(self):
        t = Timing()
        self._check(self._lib.srb_multi_get_timing(self._ctx, C.byref(t)))
        return {name: getattr(t, name) for name, _ in Timing._fields_}


class SpectralPCA:
    """srb_pca: SpectralPCA's basis (spectral_pca.cpp:155-173), trained on the host from sub-sampled pixel vectors.
    images: list of [C][...] arrays of one shape.  num_pca_bands > 0, or retained_variance in (0, 1]."""

    def __init__(self, images, num_pca_bands=0, retained_variance=0.0):
        self._lib = load_library()
        imgs = [_f64(im) for im in images]
        if not imgs:
            raise SrbError(1, "at least one image is required to compute the PCA basis")
        Cn = imgs[0].shape[0]
        P = int(np.prod(imgs[0].shape[1:]))
        for im in imgs:
            if im.shape != imgs[0].shape:
                raise SrbError(1, "inconsistent image shapes")
        ptrs = (C.c_void_p * len(imgs))(*[im.ctypes.data for im in imgs])
        self._p = _ctx_p()
        st = self._lib.srb_pca_create(ptrs, len(imgs), Cn, P, int(num_pca_bands), float(retained_variance), C.byref(self._p))
        if st != 0:
            raise SrbError(st, "invalid SpectralPCA arguments")
        self.num_bands = self._lib.srb_pca_num_bands(self._p)
        self.num_components = self._lib.srb_pca_num_components(self._p)
        self.mean = np.empty(self.num_bands)
        self.eigenvectors = np.empty((self.num_components, self.num_bands))
        self.eigenvalues = np.empty(self.num_components)
        self._lib.srb_pca_get(self._p, _host_ptr(self.mean), _host_ptr(self.eigenvectors), _host_ptr(self.eigenvalues))

    def close(self):
        if getattr(self, "_p", None):
            self._lib.srb_pca_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def envi_read_header(path):
    """HSIBinaryDataParameters::ReadHeaderFromFile: parses an ENVI .hdr (host only)."""
    h = EnviHeader()
    st = load_library().srb_envi_read_header(os.fsencode(path), C.byref(h))
    if st != 0:
        raise SrbError(st, "could not read ENVI header '%s'" % path)
    return h


def envi_write(path, image):
    """WriteBinaryFileBSQ<float> + .hdr + .config (host only).  image: [bands][rows][cols]."""
    image = _f64(image)
    b, r, c = image.shape
    st = load_library().srb_envi_write(os.fsencode(path), _host_ptr(image), b, r, c)
    if st != 0:
        raise SrbError(st, "could not write ENVI file '%s'" % path)


def pin_host(array):
    st = load_library().srb_pin_host(array.ctypes.data_as(C.c_void_p), array.nbytes)
    if st != 0:
        raise SrbError(st, "cudaHostRegister failed")


def unpin_host(array):
    load_library().srb_unpin_host(array.ctypes.data_as(C.c_void_p))


def dev_alloc(nbytes):
    p = C.c_void_p()
    st = load_library().srb_dev_alloc(C.byref(p), int(nbytes))
    if st != 0:
        raise SrbError(st, "cudaMalloc failed")
    return p.value


def dev_free(ptr):
    load_library().srb_dev_free(C.c_void_p(ptr))


def ipc_export(ptr):
    buf = C.create_string_buffer(64)
    st = load_library().srb_ipc_export(C.c_void_p(ptr), buf)
    if st != 0:
        raise SrbError(st, "cudaIpcGetMemHandle failed")
    return bytes(buf.raw)


def ipc_open(handle):
    p = C.c_void_p()
    st = load_library().srb_ipc_open(C.c_char_p(handle), C.byref(p))
    if st != 0:
        raise SrbError(st, "cudaIpcOpenMemHandle failed (no peer access between these GPUs?)")
    return p.value


def ipc_close(ptr):
    load_library().srb_ipc_close(C.c_void_p(ptr))


def plan(lr_shape, scale, psf=None, shifts=None):
    """srb_plan: what srb_create would decide for this model -- host only, no CUDA device needed.
    Returns a dict; raises SrbError for descriptions the reference would CHECK-fail on."""
    lib = load_library()
    N, Cn, h, w = (int(v) for v in lr_shape)
    psf_a = None if psf is None else _f64(psf)
    sh = None if shifts is None else _f64(shifts).reshape(-1, 2)
    desc = ModelDesc(h, w, Cn, N, int(scale), 0 if psf_a is None else psf_a.shape[0],
                     None if psf_a is None else psf_a.ctypes.data_as(_dp),
                     None if sh is None else sh.ctypes.data_as(_dp))
    info = PlanInfo()
    st = lib.srb_plan(C.byref(desc), C.byref(info))
    if st != 0:
        raise SrbError(st, info.why.decode())
    out = {name: getattr(info, name) for name, _ in PlanInfo._fields_}
    out["why"] = info.why.decode()
    return out


def quantize_shift(d):
    return load_library().srb_quantize_shift(float(d))


def sample_is_special(q, hr_size, psf_half, scale, shift):
    return bool(load_library().srb_sample_is_special(int(q), int(hr_size), int(psf_half), int(scale), float(shift)))
