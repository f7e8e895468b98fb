"""ORACLE -- test infrastructure only (see oracle/oracle.cpp header).

ctypes wrapper of the CPU restatement.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from flecsolve_b200/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

_lib = None

STOP_REASONS = ["converged_atol", "converged_rtol", "converged_user", "diverged_dtol", "diverged_iters",
                "diverged_breakdown", "unknown"]


class Info(C.Structure):
    _fields_ = [("status", C.c_int), ("iters", C.c_int), ("restarts", C.c_int),
                ("res_norm_initial", C.c_float), ("res_norm_final", C.c_float),
                ("sol_norm_initial", C.c_float), ("sol_norm_final", C.c_float), ("rhs_norm", C.c_float),
                ("history_len", C.c_int)]

    @property
    def reason(self) -> str:
        return STOP_REASONS[self.status]


_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int64)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            from oracle import build as _b
            _b.build_oracle()
        L = C.CDLL(LIB_PATH)
        L.orc_max_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_stencil_nnz.restype = C.c_int64
        L.orc_stencil_nnz.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64]
        L.orc_stencil_fill.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double, _pi, _pi, _pd]
        L.orc_csr_spmv.argtypes = [C.c_int64, _pi, _pi, _pd, _pd, _pd]
        L.orc_parcsr_create.restype = C.c_void_p
        L.orc_parcsr_create.argtypes = [C.c_int64, C.c_int, _pi, _pi, _pi, _pd]
        L.orc_parcsr_destroy.argtypes = [C.c_void_p]
        L.orc_parcsr_partition.argtypes = [C.c_void_p, _pi]
        L.orc_parcsr_sizes.argtypes = [C.c_void_p, C.c_int, _pi]
        L.orc_parcsr_block.argtypes = [C.c_void_p, C.c_int, C.c_int, _pi, _pi, _pd, _pi]
        L.orc_parcsr_spmv.argtypes = [C.c_void_p, _pd, _pd]
        L.orc_parcsr_dinv.argtypes = [C.c_void_p, _pd]
        L.orc_jacobi_relax.argtypes = [C.c_void_p, C.c_double, C.c_int64, _pd, _pd]
        L.orc_dot.restype = C.c_double
        L.orc_dot.argtypes = [C.c_void_p, _pd, _pd]
        L.orc_set_random.argtypes = [C.c_void_p, _pd, C.c_uint]
        solver_args = [C.c_void_p, _pd, C.c_int, C.c_float, C.c_int]
        tail = [_pd, _pd, C.POINTER(Info), _pd, C.c_int]
        L.orc_cg.argtypes = solver_args + tail
        L.orc_bicgstab.argtypes = solver_args + tail
        L.orc_gmres.argtypes = solver_args + [C.c_int, C.c_int, C.c_int] + tail
        L.orc_vec_op.argtypes = [C.c_int, C.c_int64, _pd, _pd, _pd, C.c_double, C.c_double]
        L.orc_vec_reduce.restype = C.c_double
        L.orc_vec_reduce.argtypes = [C.c_int, C.c_int64, _pd, _pd, C.c_double]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_pd) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_pi) if a is not None else None


def max_threads() -> int:
    return lib().orc_max_threads()


def set_threads(n: int) -> None:
    lib().orc_set_threads(n)


def stencil_csr(kind: int, nx: int, ny: int, nz: int = 1, diag_shift: float = 0.0, scale: float = 1.0):
    """Global CSR (int64 rowptr/col, fp64 val) of the synthetic operators of SURVEY.md section 8d."""
    n = nx * ny * nz
    nnz = lib().orc_stencil_nnz(kind, nx, ny, nz)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    col = np.zeros(nnz, dtype=np.int64)
    val = np.zeros(nnz, dtype=np.float64)
    lib().orc_stencil_fill(kind, nx, ny, nz, diag_shift, scale, _i(rowptr), _i(col), _d(val))
    assert rowptr[-1] == nnz
    return rowptr, col, val


def csr_spmv(rowptr, col, val, x) -> np.ndarray:
    y = np.zeros(len(rowptr) - 1)
    rowptr, col = np.ascontiguousarray(rowptr, dtype=np.int64), np.ascontiguousarray(col, dtype=np.int64)
    val = np.ascontiguousarray(val, dtype=np.float64)
    lib().orc_csr_spmv(len(rowptr) - 1, _i(rowptr), _i(col), _d(val), _d(np.ascontiguousarray(x, dtype=np.float64)), _d(y))
    return y


VEC_OPS = {"copy": 0, "set": 1, "scale": 2, "add": 3, "subtract": 4, "multiply": 5, "divide": 6, "reciprocal": 7,
           "linear_sum": 8, "axpy": 9, "axpby": 10, "abs": 11, "add_scalar": 12}
RED_OPS = {"dot": 0, "l1": 1, "inf": 2, "min": 3, "max": 4, "powsum": 5}


def vec_op(name: str, z, x=None, y=None, a=0.0, b=0.0) -> np.ndarray:
    """In-place on z (float64 array); x / y may be z itself (aliasing as in the reference)."""
    lib().orc_vec_op(VEC_OPS[name], z.size, _d(z), _d(x if x is not None else z), _d(y if y is not None else z), a, b)
    return z


def vec_reduce(name: str, x, y=None, a=0.0) -> float:
    return lib().orc_vec_reduce(RED_OPS[name], x.size, _d(x), _d(y if y is not None else x), a)


class ParCSR:
    """P colours of the reference's parallel CSR simulated in one process."""

    def __init__(self, rowptr, col, val, colours: int = 1, part=None):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int64)
        self.val = np.ascontiguousarray(val, dtype=np.float64)
        self.n = len(self.rowptr) - 1
        self.colours = colours
        p = np.ascontiguousarray(part, dtype=np.int64) if part is not None else None
        self.h = lib().orc_parcsr_create(self.n, colours, _i(p), _i(self.rowptr), _i(self.col), _d(self.val))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_parcsr_destroy(self.h)
            self.h = None

    def partition(self) -> np.ndarray:
        out = np.zeros(self.colours + 1, dtype=np.int64)
        lib().orc_parcsr_partition(self.h, _i(out))
        return out

    def block(self, p: int, which: int):
        """(rowptr, col, val, colmap) of colour p's diag (0) / offd (1) block, local numbering."""
        s = np.zeros(4, dtype=np.int64)
        lib().orc_parcsr_sizes(self.h, p, _i(s))
        nnz = s[2 + which]
        rp = np.zeros(s[0] + 1, dtype=np.int64)
        col = np.zeros(nnz, dtype=np.int64)
        val = np.zeros(nnz, dtype=np.float64)
        cm = np.zeros(s[1], dtype=np.int64)
        lib().orc_parcsr_block(self.h, p, which, _i(rp), _i(col), _d(val), _i(cm))
        return rp, col, val, cm

    def spmv(self, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.n)
        lib().orc_parcsr_spmv(self.h, _d(x), _d(y))
        return y

    def dinv(self) -> np.ndarray:
        d = np.zeros(self.n)
        lib().orc_parcsr_dinv(self.h, _d(d))
        return d

    def jacobi_relax(self, omega: float, nrelax: int, b, x) -> np.ndarray:
        x = np.array(x, dtype=np.float64)
        lib().orc_jacobi_relax(self.h, omega, nrelax, _d(np.ascontiguousarray(b, dtype=np.float64)), _d(x))
        return x

    def dot(self, x, y) -> float:
        return lib().orc_dot(self.h, _d(x), _d(y))

    def set_random(self, seed: int) -> np.ndarray:
        x = np.zeros(self.n)
        lib().orc_set_random(self.h, _d(x), seed)
        return x

    def _solve(self, which, b, x0, dinv, maxiter, rtol, use_zero_guess, history_cap, extra=()):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.array(x0 if x0 is not None else np.zeros(self.n), dtype=np.float64)
        hist = np.zeros(max(history_cap, 1))
        info = Info()
        d = np.ascontiguousarray(dinv, dtype=np.float64) if dinv is not None else None
        fn = getattr(lib(), which)
        fn(self.h, _d(d), maxiter, rtol, int(use_zero_guess), *extra, _d(b), _d(x), C.byref(info), _d(hist), history_cap)
        return x, info, hist[:min(info.history_len, history_cap)]

    def cg(self, b, x0=None, dinv=None, maxiter=1000, rtol=1e-9, use_zero_guess=False, history_cap=0):
        return self._solve("orc_cg", b, x0, dinv, maxiter, rtol, use_zero_guess, history_cap)

    def bicgstab(self, b, x0=None, dinv=None, maxiter=1000, rtol=1e-9, use_zero_guess=False, history_cap=0):
        return self._solve("orc_bicgstab", b, x0, dinv, maxiter, rtol, use_zero_guess, history_cap)

    def gmres(self, b, x0=None, dinv=None, maxiter=100, rtol=1e-9, use_zero_guess=False, max_krylov_dim=-1,
              right_precond=True, restart=False, history_cap=0):
        return self._solve("orc_gmres", b, x0, dinv, maxiter, rtol, use_zero_guess, history_cap,
                           extra=(max_krylov_dim, int(right_precond), int(restart)))


def fvm_diffusion_apply(u, a, bface, beta: float, alpha: float, vol: float, kface, lo, hi) -> np.ndarray:
    """The reference's matrix-free finite-volume diffusion operator  v = -beta div(b grad u) + alpha vol a u  on one
    padded array (physics/volume_diffusion/diffusion.hh): arrays are indexed [k][j][i] (x fastest), as the mdspans there.
      update_flux   (:115-160)  f_x[k][j][i] = b_x[k][j][i] * (dA_x * i_dx) * (u[k][j][i+1] - u[k][j][i])   (same for y, z)
      sum_cell_flux (:162-181)  du[k][j][i] = (f_x[k][j][i] - f_x[k][j][i-1]) + (f_y[..j..] - f_y[..j-1..]) + (f_z[k] - f_z[k-1])
      operate       (:183-203)  v[k][j][i]  = (-beta * du[k][j][i]) + (alpha * vol * a[k][j][i] * u[k][j][i])
    kface[axis] = dA_axis * i_dx_axis.  lo/hi: the dof box per axis (x first); v is returned for the dofs, 0 elsewhere.
    1-D and 2-D boxes are the same expressions without the missing axes (a sum of fewer differences)."""
    u = np.asarray(u, dtype=np.float64)
    dim = u.ndim
    a = np.asarray(a, dtype=np.float64).reshape(u.shape)
    flux = []
    for ax in range(dim):  # ax 0 = x = the LAST numpy axis
        npax = dim - 1 - ax
        b = np.asarray(bface[ax], dtype=np.float64).reshape(u.shape)
        f = np.zeros_like(u)
        upper = [slice(None)] * dim
        lower = [slice(None)] * dim
        upper[npax] = slice(1, None)
        lower[npax] = slice(0, -1)
        f[tuple(lower)] = b[tuple(lower)] * kface[ax] * (u[tuple(upper)] - u[tuple(lower)])
        flux.append(f)
    box = tuple(slice(lo[dim - 1 - d], hi[dim - 1 - d]) for d in range(dim))  # numpy axis d is mesh axis dim-1-d
    du = None
    for ax in range(dim):
        npax = dim - 1 - ax
        shifted = tuple(slice(s.start - 1, s.stop - 1) if d == npax else s for d, s in enumerate(box))
        term = flux[ax][box] - flux[ax][shifted]
        du = term if du is None else du + term
    v = np.zeros_like(u)
    v[box] = (-beta * du) + (alpha * vol * a[box] * u[box])
    return v

