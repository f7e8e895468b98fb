"""ctypes binding of libfsb_host.so: the compiled C++ host layer (flecsolve-shaped headers
under flecsolve_b200/include/) driven from Python by tests and bench.py."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as F

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfsb_host.so")
_lib = None

STOP_REASONS = ["converged_atol", "converged_rtol", "converged_user", "diverged_dtol", "diverged_iters",
                "diverged_breakdown", "unknown"]
SOLVERS = {"cg": 0, "gmres": 1, "bicgstab": 2, "fcg": 3, "cg_device": 4, "cg_sr": 5}
PRECONDS = {None: 0, "none": 0, "identity": 0, "dinv": 1, "jacobi": 1, "relax": 2}


class Info(C.Structure):
    _fields_ = [("status", C.c_int), ("iters", C.c_int), ("restarts", C.c_int),
                ("res_norm_initial", C.c_float), ("res_norm_final", C.c_float),
                ("sol_norm_initial", C.c_float), ("sol_norm_final", C.c_float), ("rhs_norm", C.c_float),
                ("callbacks", C.c_int), ("window_launches", C.c_int), ("solve_ms", C.c_double),
                ("h2d_ms", C.c_double), ("d2h_ms", C.c_double)]

    @property
    def reason(self) -> str:
        return STOP_REASONS[self.status]


class Options(C.Structure):
    _fields_ = [("solver", C.c_int), ("precond", C.c_int), ("omega", C.c_float), ("nrelax", C.c_int),
                ("maxiter", C.c_int), ("rtol", C.c_float), ("atol", C.c_float), ("use_zero_guess", C.c_int),
                ("max_krylov_dim", C.c_int), ("restart", C.c_int), ("pre_side_left", C.c_int),
                ("ev_start", C.c_int), ("ev_stop", C.c_int), ("lag", C.c_int)]


class BdfOptions(C.Structure):
    _fields_ = [("method", C.c_int), ("starting", C.c_int), ("predictor", C.c_int), ("strategy", C.c_int),
                ("controller", C.c_int), ("use_pi_controller", C.c_int), ("norm", C.c_int), ("error_scaling", C.c_int),
                ("time_rtol", C.c_double), ("time_atol", C.c_double), ("problem_scale", C.c_double),
                ("initial_time", C.c_double), ("final_time", C.c_double), ("initial_dt", C.c_double),
                ("max_dt", C.c_double), ("min_dt", C.c_double), ("max_steps", C.c_int), ("max_attempts", C.c_int)]


class BdfResult(C.Structure):
    _fields_ = [("attempts", C.c_int), ("steps", C.c_int), ("rejects", C.c_int), ("final_time", C.c_double),
                ("value_max", C.c_double), ("value_l2", C.c_double), ("inner_iterations", C.c_int), ("solve_ms", C.c_double)]


class RkOptions(C.Structure):
    _fields_ = [("order", C.c_int), ("initial_time", C.c_double), ("final_time", C.c_double), ("initial_dt", C.c_double),
                ("max_dt", C.c_double), ("min_dt", C.c_double), ("max_steps", C.c_int), ("safety_factor", C.c_float),
                ("atol", C.c_float), ("use_fixed_dt", C.c_int)]


BDF_METHODS = {"CN": 0, "BE": 1, "BDF2": 2, "BDF3": 3, "BDF4": 4, "BDF5": 5, "BDF6": 6}
BDF_CONTROLLERS = {"H211b": 0, "PC.4.7": 1, "PC11": 2, "Deadbeat": 3}


def make_bdf_options(method="BDF2", starting="CN", predictor="leapfrog", strategy="truncation-error",
                     controller="PC.4.7", use_pi_controller=True, norm="inf", error_scaling="fixed-resolution",
                     time_rtol=1e-9, time_atol=1e-15, problem_scale=1.0, initial_time=0.0, final_time=1.0,
                     initial_dt=0.0, max_dt=float("inf"), min_dt=0.0, max_steps=1000, max_attempts=0) -> BdfOptions:
    strategies = {"truncation-error": 0, "constant": 1, "final-constant": 2, "limit-relative-change": 3}
    return BdfOptions(BDF_METHODS[method], BDF_METHODS[starting], {"ab2": 0, "leapfrog": 1}[predictor],
                      strategies[strategy], BDF_CONTROLLERS[controller], int(use_pi_controller),
                      {"inf": 0, "l2": 1}[norm], {"fixed-resolution": 0, "fixed-scaling": 1}[error_scaling],
                      time_rtol, time_atol, problem_scale, initial_time, final_time, initial_dt, max_dt, min_dt,
                      max_steps, max_attempts)


def make_options(solver="cg", precond=None, maxiter=1000, rtol=1e-9, atol=0.0, use_zero_guess=False, omega=2 / 3,
                 nrelax=1, max_krylov_dim=-1, restart=False, pre_side="right", ev_start=-1, ev_stop=-1, lag=2) -> Options:
    return Options(SOLVERS[solver], PRECONDS[precond], omega, nrelax, maxiter, rtol, atol, int(use_zero_guess),
                   max_krylov_dim, int(restart), int(pre_side == "left"), ev_start, ev_stop, lag)


_pd = C.POINTER(C.c_double)
_pl = C.POINTER(C.c_int64)


def declare(L: C.CDLL) -> C.CDLL:
    """argument types of the fsbh_* entry points (flecsolve_b200/host/driver.cpp)"""
    L.fsbh_last_error.restype = C.c_char_p
    L.fsbh_session_create.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.fsbh_session_destroy.argtypes = [C.c_void_p]
    L.fsbh_session_b.restype = C.c_void_p
    L.fsbh_session_b.argtypes = [C.c_void_p]
    L.fsbh_session_x.restype = C.c_void_p
    L.fsbh_session_x.argtypes = [C.c_void_p]
    L.fsbh_solve.argtypes = [C.c_void_p, C.POINTER(Options), _pd, _pd, C.POINTER(Info), _pd, C.c_int]
    L.fsbh_adapter_apply.argtypes = [C.c_void_p, C.c_double, _pd, _pd]
    L.fsbh_solve_multi2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Options), _pd, _pd,
                                    C.POINTER(Info), _pd, C.c_int]
    L.fsbh_solve_subset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Options), _pd, _pd,
                                    C.POINTER(Info)]
    L.fsbh_vector_selftest.argtypes = [C.c_void_p, _pd]
    L.fsbh_multivector_selftest.argtypes = [C.c_void_p, _pd]
    _pi = C.POINTER(C.c_int)
    L.fsbh_bdf_rate.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(BdfOptions), C.c_double, C.c_double,
                                C.POINTER(BdfResult), _pd, _pi, _pd, C.c_int]
    L.fsbh_rk_rate.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(RkOptions), C.c_double, C.c_double,
                               C.POINTER(BdfResult), _pd, _pi, _pd, C.c_int]
    L.fsbh_bdf_heat.argtypes = [C.c_void_p, C.POINTER(BdfOptions), C.POINTER(Options), _pd, _pd,
                                C.POINTER(BdfResult), _pd, _pi, _pi, C.c_int]
    L.fsbh_mtx_read.argtypes = [C.c_char_p, _pl, _pl, _pl, _pi, _pl, _pl, _pd]
    L.fsbh_mtx_create.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
    L.fsbh_config_dump.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
    L.fsbh_solve_config.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, _pd, _pd, C.POINTER(Info), _pd, C.c_int]
    L.fsbh_poisson_narray.argtypes = [C.c_void_p, C.c_int, C.POINTER(Options), C.c_uint, _pd, C.POINTER(Info), _pd, C.c_int]
    return L


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        F.lib()  # libfsb.so first (and the NCCL preload)
        if not os.path.exists(LIB_PATH):
            raise F.FsbError(-1, f"{LIB_PATH} is missing: run __graft_entry__.build()")
        _lib = declare(C.CDLL(LIB_PATH))
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise F.FsbError(rc, lib().fsbh_last_error().decode())


def _d(a):
    return a.ctypes.data_as(_pd) if a is not None else None


class Session:
    """One operator (a parallel CSR matrix) with its solution / right-hand-side fields."""

    def __init__(self, ctx: F.Context, A: F.ParCSR):
        self.ctx, self.A = ctx, A
        h = C.c_void_p()
        _check(lib().fsbh_session_create(ctx.h, A.h, C.byref(h)))
        self.h = h
        n, g = A.local_rows, A.num_ghosts
        self.b = F.Vector(ctx, n, g, handle=C.c_void_p(lib().fsbh_session_b(h)))
        self.x = F.Vector(ctx, n, g, handle=C.c_void_p(lib().fsbh_session_x(h)))

    def close(self):
        if self.h:
            _check(lib().fsbh_session_destroy(self.h))
            self.h = None

    def solve(self, b=None, x0=None, history_cap=0, inplace=False, **kw):
        """Host-buffer call when b is given (uploads b and x0, downloads x); device-resident
        otherwise (uses self.b / self.x as they are).  inplace: x0 (a contiguous float64 array, e.g. pinned memory)
        is the in/out buffer itself instead of being copied."""
        opts = make_options(**kw)
        info = Info()
        hist = np.zeros(max(history_cap, 1))
        if b is None:
            _check(lib().fsbh_solve(self.h, C.byref(opts), None, None, C.byref(info), _d(hist), history_cap))
            x = None
        else:
            b = np.ascontiguousarray(b, dtype=np.float64)
            if inplace:
                assert x0 is not None and x0.dtype == np.float64 and x0.flags.c_contiguous
                x = x0
            else:
                x = np.array(x0 if x0 is not None else np.zeros_like(b), dtype=np.float64)
            _check(lib().fsbh_solve(self.h, C.byref(opts), _d(b), _d(x), C.byref(info), _d(hist), history_cap))
        return x, info, hist[:min(info.callbacks, history_cap)]

    def solve_config(self, cfg: str, prefix: str, b, x0=None, precond=None, history_cap=0):
        """solve with the solver `[prefix] type = ...` names (krylov_factory)"""
        info = Info()
        hist = np.zeros(max(history_cap, 1))
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.array(x0 if x0 is not None else np.zeros_like(b), dtype=np.float64)
        _check(lib().fsbh_solve_config(self.h, cfg.encode(), prefix.encode(), PRECONDS[precond], _d(b), _d(x),
                                       C.byref(info), _d(hist), history_cap))
        return x, info, hist[:min(info.callbacks, history_cap)]

    def adapter_apply(self, gamma: float, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        _check(lib().fsbh_adapter_apply(self.h, gamma, _d(x), _d(y)))
        return y

    def bdf_heat(self, u0, bdf: BdfOptions, cap=4096, **solver_kw):
        """u_t = A u by time_integrator::bdf with a Krylov inner solver on (I - gamma A)."""
        so = make_options(**solver_kw)
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        u = np.zeros_like(u0)
        res = BdfResult()
        dts, good, iters = np.zeros(cap), np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
        _check(lib().fsbh_bdf_heat(self.h, C.byref(bdf), C.byref(so), _d(u0), _d(u), C.byref(res), _d(dts),
                                   good.ctypes.data_as(C.POINTER(C.c_int)), iters.ctypes.data_as(C.POINTER(C.c_int)), cap))
        k = min(res.attempts, cap)
        return u, res, dts[:k], good[:k], iters[:k]

    def multivector_selftest(self) -> np.ndarray:
        """closed forms of vectors/test/flecsi_multivector.cc on a four-component vec::multi (see driver.cpp)"""
        out = np.zeros(18)
        _check(lib().fsbh_multivector_selftest(self.h, _d(out)))
        return out

    def vector_selftest(self) -> np.ndarray:
        out = np.zeros(16)
        _check(lib().fsbh_vector_selftest(self.h, _d(out)))
        return out


def read_mtx(path: str):
    """Matrix Market file -> (nrows, ncols, symmetric, rowptr, col, val) with the reference reader's rules
    (float-rounded values, mirrored symmetric entries, stable sort by row)."""
    nr, nc, nz, sym = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
    args = (path.encode(), C.byref(nr), C.byref(nc), C.byref(nz), C.byref(sym))
    _check(lib().fsbh_mtx_read(*args, None, None, None))
    rp, col, val = np.zeros(nr.value + 1, dtype=np.int64), np.zeros(nz.value, dtype=np.int64), np.zeros(nz.value)
    _check(lib().fsbh_mtx_read(*args, rp.ctypes.data_as(_pl), col.ctypes.data_as(_pl), _d(val)))
    return nr.value, nc.value, bool(sym.value), rp, col, val


def mtx_create(ctx: F.Context, path: str) -> F.ParCSR:
    """this rank's equal row block of a Matrix Market file as a device matrix"""
    h = C.c_void_p()
    _check(lib().fsbh_mtx_create(ctx.h, path.encode(), C.byref(h)))
    return F.ParCSR(ctx, h)


CONFIG_KINDS = {"solver": 0, "krylov": 1, "bdf": 2, "heat": 3}


def config_dump(path: str, kind: str, prefix: str) -> dict:
    """read_config(path, <kind>::options(prefix)) -> the settings as a dict"""
    import json
    buf = C.create_string_buffer(4096)
    _check(lib().fsbh_config_dump(path.encode(), CONFIG_KINDS[kind], prefix.encode(), buf, len(buf)))
    return json.loads(buf.value.decode())


def poisson_narray(ctx: F.Context, m: int, seed: int = 7, history_cap=0, **kw):
    """examples/poisson on an narray mesh (fields = padded 2-D arrays, vectors act on the interior)"""
    opts = make_options(**kw)
    info = Info()
    u = np.zeros(m * m)
    hist = np.zeros(max(history_cap, 1))
    _check(lib().fsbh_poisson_narray(ctx.h, m, C.byref(opts), seed, _d(u), C.byref(info), _d(hist), history_cap))
    return u, info, hist[:min(info.callbacks, history_cap)]


def solve_multi2(ctx: F.Context, A0: F.ParCSR, A1: F.ParCSR, b, x0, history_cap=0, **kw):
    opts = make_options(**kw)
    info = Info()
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.array(x0, dtype=np.float64)
    hist = np.zeros(max(history_cap, 1))
    _check(lib().fsbh_solve_multi2(ctx.h, A0.h, A1.h, C.byref(opts), _d(b), _d(x), C.byref(info), _d(hist), history_cap))
    return x, info, hist[:min(info.callbacks, history_cap)]


def solve_subset(ctx: F.Context, A: F.ParCSR, which: int, b, x0, **kw):
    """CG bound to one variable of a two-component vec::multi (solvers/test/cgmulti.cc)."""
    opts = make_options(**kw)
    info = Info()
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.array(x0, dtype=np.float64)
    _check(lib().fsbh_solve_subset(ctx.h, A.h, which, C.byref(opts), _d(b), _d(x), C.byref(info)))
    return x, info


def bdf_rate(ctx: F.Context, A: F.ParCSR, bdf: BdfOptions, lam: float, ic: float, cap=4096):
    """x' = lam x on every entry of a vector on A's topology (time-integrators/test/implicit.cc)."""
    res = BdfResult()
    dts, good, vals = np.zeros(cap), np.zeros(cap, dtype=np.int32), np.zeros(cap)
    _check(lib().fsbh_bdf_rate(ctx.h, A.h, C.byref(bdf), lam, ic, C.byref(res), _d(dts),
                               good.ctypes.data_as(C.POINTER(C.c_int)), _d(vals), cap))
    k = min(res.attempts, cap)
    return res, dts[:k], good[:k], vals[:k]


def rk_rate(ctx: F.Context, A: F.ParCSR, order: int, lam: float, ic: float, initial_dt, max_dt, min_dt, final_time,
            safety_factor, atol, use_fixed_dt, cap=4096):
    """x' = lam x with rk23 / rk45 (time-integrators/test/explicit.cc)."""
    o = RkOptions(order, 0.0, final_time, initial_dt, max_dt, min_dt, 1000, safety_factor, atol, int(use_fixed_dt))
    res = BdfResult()
    dts, good, vals = np.zeros(cap), np.zeros(cap, dtype=np.int32), np.zeros(cap)
    _check(lib().fsbh_rk_rate(ctx.h, A.h, C.byref(o), lam, ic, C.byref(res), _d(dts),
                              good.ctypes.data_as(C.POINTER(C.c_int)), _d(vals), cap))
    k = min(res.attempts, cap)
    return res, dts[:k], good[:k], vals[:k]
