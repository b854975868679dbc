"""Objectives and settings shared by tests/test_cg_restatement.py and tools/make_cg_golden.py: the
cases on which the CG restatement (super-resolution_b200/csrc/srb_cg.h) is pinned against ALGLIB's
mincg.  They cover the four termination rules the reference can hit (EpsF, EpsX, EpsG, MaxIts), the
automatic EpsX, restarts, all trial-step cases of the line search and the function trimming."""
import ctypes as C

import numpy as np

FG = C.CFUNCTYPE(None, C.c_longlong, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                 C.c_void_p)
ARGTYPES = [C.c_longlong, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double, C.c_int, FG, C.c_void_p,
            C.POINTER(C.c_double)]


def rosenbrock(x):
    f = np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2)
    g = np.zeros_like(x)
    g[:-1] += -400 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2 * (1 - x[:-1])
    g[1:] += 200 * (x[1:] - x[:-1] ** 2)
    return f, g


def _quadratic():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((40, 40))
    A = A @ A.T + 1e-3 * np.eye(40)
    b = rng.standard_normal(40)
    return lambda x: (0.5 * x @ A @ x - b @ x, A @ x - b)


def pole(x):
    """Steep wall near 0: overshooting line-search steps land where the function is trimmed."""
    f = np.sum((x - 1.0) ** 2 + 1e-3 / x ** 2)
    g = 2.0 * (x - 1.0) - 2e-3 / x ** 3
    return float(f), g


def lying_gradient(x):
    """A convex bowl whose reported gradient points uphill: the line search spends its
    evaluation budget without finding a lower point."""
    return float(np.sum(x * x)), -2.0 * x


def kinked(x):
    """|x|-like valley with a discontinuous slope: line searches end on the interval / evaluation
    budget rules instead of the Wolfe conditions."""
    f = np.sum(np.abs(x) + 0.05 * x * x)
    return float(f), np.sign(x) + 0.1 * x


class NanAfter:
    """Rosenbrock that turns into NaN after a number of evaluations (termination type -8)."""

    def __init__(self, limit):
        self.limit = limit
        self.calls = 0

    def reset(self):
        self.calls = 0

    def __call__(self, x):
        self.calls += 1
        f, g = rosenbrock(x)
        if self.calls > self.limit:
            return float("nan"), g
        return f, g


class _SrObjective:
    """The MAP objective itself (oracle): 32 x 40 HR, 2x, 3x3 PSF, 4 frames, TV.  Built on first use so
    that importing this module never touches the oracle library."""
    S, H_LR, W_LR = 2, 16, 20

    def __init__(self):
        self._fg = None

    @property
    def x0(self):
        return np.full(self.H_LR * self.S * self.W_LR * self.S, 0.5)

    def _build(self):
        from oracle import sr_oracle as o
        rng = np.random.default_rng(7)
        s, h, w = self.S, self.H_LR, self.W_LR
        psf = o.gaussian_psf(3, 0.8)
        shifts = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], dtype=np.float64)
        m = o.Model(s, psf, shifts)
        truth = rng.random((h * s, w * s))
        lr = np.stack([o.forward(m, k, truth) for k in range(4)])[:, None] + 0.01 * rng.standard_normal((4, 1, h, w))
        obs = o.upsample_observations(m, lr)

        def fg(x):
            f, g = o.evaluate(m, x.reshape(1, h * s, w * s), obs, reg_kind=o.REG_TV, lam=0.01,
                              weights=np.ones((1, h * s, w * s)))
            return f, g.ravel()
        self._keep = (m, obs)
        return fg

    def __call__(self, x):
        if self._fg is None:
            self._fg = self._build()
        return self._fg(x)


def cases():
    rng = np.random.default_rng(1)
    quad = _quadratic()
    sr = _SrObjective()
    sr_x0 = sr.x0
    return [
        ("rosenbrock10_epsg", rosenbrock, rng.standard_normal(10), dict(epsg=1e-10, maxits=500)),
        ("rosenbrock100_epsf", rosenbrock, -1.2 * np.ones(100), dict(epsf=1e-12, maxits=300)),
        ("quadratic_epsg", quad, np.zeros(40), dict(epsg=1e-9)),
        ("quadratic_auto_epsx", quad, np.ones(40), dict()),
        ("rosenbrock7_epsx", rosenbrock, rng.standard_normal(7), dict(epsx=1e-9)),
        ("rosenbrock33_maxits", rosenbrock, rng.standard_normal(33), dict(maxits=25)),
        ("pole_trimmed", pole, np.full(12, 3.0) + rng.random(12), dict(epsg=1e-12, maxits=200)),
        ("map_objective_tv", sr, sr_x0, dict(epsg=1e-8, epsf=1e-14, epsx=1e-12, maxits=40)),
        ("lying_gradient", lying_gradient, rng.standard_normal(9), dict(epsg=1e-12, maxits=100)),
        ("kinked_valley", kinked, rng.standard_normal(15) * 3, dict(epsg=1e-9, epsx=1e-13, maxits=60)),
        ("nan_after_12_evaluations", NanAfter(12), rng.standard_normal(6), dict(epsg=1e-12, maxits=100)),
    ]


ARGTYPES_LBFGS = [C.c_longlong, C.POINTER(C.c_double), C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, FG,
                  C.c_void_p, C.POINTER(C.c_double)]


def run(fn, x0, fg, epsg=0.0, epsf=0.0, epsx=0.0, maxits=0, lbfgs_m=0):
    """Runs one solver entry point (ref_mincg / srbcg_host_minimize, or with lbfgs_m > 0 ref_minlbfgs /
    srbcg_host_lbfgs).  Returns (x, report, f trace)."""
    fn.restype = C.c_int
    fn.argtypes = ARGTYPES_LBFGS if lbfgs_m > 0 else ARGTYPES
    if hasattr(fg, "reset"):
        fg.reset()
    x = np.array(x0, dtype=np.float64)
    rep = np.zeros(6)
    trace = []

    def cb(n, xp, fp, gp, user):
        f, g = fg(np.ctypeslib.as_array(xp, (n,)))
        fp[0] = f
        np.ctypeslib.as_array(gp, (n,))[:] = g
        trace.append(float(f))
    head = (len(x), x.ctypes.data_as(C.POINTER(C.c_double))) + ((lbfgs_m,) if lbfgs_m > 0 else ())
    fn(*head, epsg, epsf, epsx, maxits, FG(cb), None, rep.ctypes.data_as(C.POINTER(C.c_double)))
    return x, rep, np.array(trace)
