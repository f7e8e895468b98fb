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
SOLVERS = {"cg": 0, "gmres": 1, "bicgstab": 2, "fcg": 3}
PRECONDS = {None: 0, "none": 0, "identity": 0, "dinv": 1, "jacobi": 1, "relax": 2}


class Info(C.Structure):
    _fields_ = [("status", C.c_int), ("iters", C.c_int), ("restarts", C.c_int),
                ("res_norm_initial", C.c_float), ("res_norm_final", C.c_float),
                ("sol_norm_initial", C.c_float), ("sol_norm_final", C.c_float), ("rhs_norm", C.c_float),
                ("callbacks", C.c_int), ("window_launches", C.c_int)]

    @property
    def reason(self) -> str:
        return STOP_REASONS[self.status]


class Options(C.Structure):
    _fields_ = [("solver", C.c_int), ("precond", C.c_int), ("omega", C.c_float), ("nrelax", C.c_int),
                ("maxiter", C.c_int), ("rtol", C.c_float), ("atol", C.c_float), ("use_zero_guess", C.c_int),
                ("max_krylov_dim", C.c_int), ("restart", C.c_int), ("pre_side_left", C.c_int),
                ("ev_start", C.c_int), ("ev_stop", C.c_int)]


def make_options(solver="cg", precond=None, maxiter=1000, rtol=1e-9, atol=0.0, use_zero_guess=False, omega=2 / 3,
                 nrelax=1, max_krylov_dim=-1, restart=False, pre_side="right", ev_start=-1, ev_stop=-1) -> Options:
    return Options(SOLVERS[solver], PRECONDS[precond], omega, nrelax, maxiter, rtol, atol, int(use_zero_guess),
                   max_krylov_dim, int(restart), int(pre_side == "left"), ev_start, ev_stop)


_pd = C.POINTER(C.c_double)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        F.lib()  # libfsb.so first (and the NCCL preload)
        if not os.path.exists(LIB_PATH):
            raise F.FsbError(-1, f"{LIB_PATH} is missing: run __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)
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
        _lib = L
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

    def solve(self, b=None, x0=None, history_cap=0, **kw):
        """Host-buffer call when b is given (uploads b and x0, downloads x); device-resident
        otherwise (uses self.b / self.x as they are)."""
        opts = make_options(**kw)
        info = Info()
        hist = np.zeros(max(history_cap, 1))
        if b is None:
            _check(lib().fsbh_solve(self.h, C.byref(opts), None, None, C.byref(info), _d(hist), history_cap))
            x = None
        else:
            b = np.ascontiguousarray(b, dtype=np.float64)
            x = np.array(x0 if x0 is not None else np.zeros_like(b), dtype=np.float64)
            _check(lib().fsbh_solve(self.h, C.byref(opts), _d(b), _d(x), C.byref(info), _d(hist), history_cap))
        return x, info, hist[:min(info.callbacks, history_cap)]

    def adapter_apply(self, gamma: float, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        _check(lib().fsbh_adapter_apply(self.h, gamma, _d(x), _d(y)))
        return y

    def vector_selftest(self) -> np.ndarray:
        out = np.zeros(16)
        _check(lib().fsbh_vector_selftest(self.h, _d(out)))
        return out


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
