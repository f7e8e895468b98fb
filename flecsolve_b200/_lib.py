"""ctypes binding of libfsb.so (include/fsb.h).  Used by tests and bench.py only;
the product's host layer is the C++ header tree under flecsolve_b200/include/.

There is deliberately no fallback: if the shared library is missing this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfsb.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "fsb.h")

_lib = None


class FsbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fsb error {code}: {msg}")
        self.code = code


def declared_symbols() -> list[str]:
    """Every function name include/fsb.h declares."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fsb_[a-z0-9_]+)\s*\(", src)))


def _preload_nccl() -> None:
    """libfsb.so needs libnccl.so.2.  torch bundles a newer one than the system's; whichever is
    loaded first wins for the whole process (same soname), and torch cannot run on the older
    one -- so load torch's copy first when it exists."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
    except Exception:
        pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FsbError(-1, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        _preload_nccl()
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


_i64 = C.c_int64
_p = C.c_void_p
_dbl = C.c_double
_pd = C.POINTER(C.c_double)
_pi64 = C.POINTER(C.c_int64)
_pi32 = C.POINTER(C.c_int32)


class Coef(C.Structure):
    """fsb_coef: scale * scalar[num] / scalar[den]; (v, 0, 0) is the plain number v."""
    _fields_ = [("scale", C.c_double), ("num", C.c_int32), ("den", C.c_int32)]


class ScalarOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("dst", C.c_int32), ("a", C.c_int32), ("b", C.c_int32)]


class RedOpts(C.Structure):
    _fields_ = [("store", C.c_int32), ("halt_mode", C.c_int), ("halt_threshold", C.c_double), ("n_post", C.c_int32),
                ("post", ScalarOp * 8)]


SOP = {"add": 0, "sub": 1, "mul": 2, "div": 3, "copy": 4}


HALT_NEVER, HALT_IF_SQRT_LT, HALT_IF_LT = 0, 1, 2


def _declare(L: C.CDLL) -> None:
    def f(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    f("fsb_last_error", C.c_char_p)
    f("fsb_version", C.c_int)
    f("fsb_device_count", C.c_int)
    f("fsb_nccl_unique_id", C.c_int, _p)
    f("fsb_ctx_create", C.c_int, C.c_int, C.c_int, C.c_int, _p, C.POINTER(_p))
    f("fsb_ctx_create_group", C.c_int, C.c_int, C.c_int, C.POINTER(_p))
    f("fsb_ctx_destroy", C.c_int, _p)
    f("fsb_ctx_flush", C.c_int, _p)
    f("fsb_ctx_sync", C.c_int, _p)
    f("fsb_ctx_stream", _p, _p)
    f("fsb_ctx_rank", C.c_int, _p)
    f("fsb_ctx_nranks", C.c_int, _p)
    f("fsb_ctx_set_option", C.c_int, _p, C.c_int, _i64)
    f("fsb_ctx_get_stat", C.c_int, _p, C.c_int, _pi64)
    f("fsb_ctx_reset_stats", C.c_int, _p)
    f("fsb_ctx_flush_l2", C.c_int, _p)
    f("fsb_ctx_event_record", C.c_int, _p, C.c_int)
    f("fsb_ctx_event_elapsed_ms", C.c_int, _p, C.c_int, C.c_int, _pd)
    f("fsb_ctx_profile_read", C.c_int, _p, _pd, _pi64)
    f("fsb_ctx_profile_read_split", C.c_int, _p, _pd, _pi64)
    f("fsb_ctx_timeline_read", C.c_int, _p, C.POINTER(C.c_uint64), _i64, _pi64)
    f("fsb_debug_program_info", C.c_int, _pi32, C.c_int, C.c_int, _pi32, _pi32)
    f("fsb_debug_jit_compile", C.c_int, _pi32, C.c_int, C.c_int, C.c_int, _pi64, C.c_char_p, C.c_int)
    f("fsb_vec_create", C.c_int, _p, _i64, _i64, C.POINTER(_p))
    f("fsb_vec_wrap", C.c_int, _p, _p, _i64, _i64, C.POINTER(_p))
    f("fsb_vec_destroy", C.c_int, _p)
    f("fsb_vec_local_size", _i64, _p)
    f("fsb_vec_ghost_size", _i64, _p)
    f("fsb_vec_device_ptr", _p, _p)
    f("fsb_vec_upload", C.c_int, _p, _pd, _i64, _i64)
    f("fsb_vec_download", C.c_int, _p, _pd, _i64, _i64)
    f("fsb_vec_copy", C.c_int, _p, _p)
    f("fsb_vec_set", C.c_int, _p, _dbl)
    f("fsb_vec_scale", C.c_int, _p, _dbl, _p)
    for n in ("add", "sub", "mul", "div"):
        f(f"fsb_vec_{n}", C.c_int, _p, _p, _p)
    f("fsb_vec_recip", C.c_int, _p, _p)
    f("fsb_vec_linear_sum", C.c_int, _p, _dbl, _p, _dbl, _p)
    f("fsb_vec_axpy", C.c_int, _p, _dbl, _p, _p)
    f("fsb_vec_axpby", C.c_int, _p, _dbl, _dbl, _p)
    f("fsb_vec_abs", C.c_int, _p, _p)
    f("fsb_vec_add_scalar", C.c_int, _p, _p, _dbl)
    f("fsb_vec_set_random", C.c_int, _p, C.c_uint)
    f("fsb_vec_dump", C.c_int, _p, C.c_char_p)
    f("fsb_vec_dot", C.c_int, _p, _p, _pi64)
    for n in ("sumsq", "asum", "amax", "min", "max"):
        f(f"fsb_vec_{n}", C.c_int, _p, _pi64)
    f("fsb_vec_powsum", C.c_int, _p, C.c_int, _pi64)
    f("fsb_vec_global_size", C.c_int, _p, _pi64)
    f("fsb_red_get", C.c_int, _p, _i64, _pd)
    f("fsb_red_wait", C.c_int, _p, _i64)
    f("fsb_vec_create_box", C.c_int, _p, C.c_int, _pi64, _pi64, _pi64, C.POINTER(_p))
    f("fsb_vec_box_upload_all", C.c_int, _p, _pd)
    f("fsb_vec_box_download_all", C.c_int, _p, _pd)
    f("fsb_parcsr_create_box_stencil", C.c_int, _p, C.c_int, _pi64, _pi64, _pi64, _dbl, _pd, C.POINTER(_p))
    f("fsb_parcsr_create_box_fvm", C.c_int, _p, C.c_int, _pi64, _pi64, _pi64, _dbl, _dbl, _dbl, _pd, _pd, C.POINTER(_pd),
      C.POINTER(_p))
    f("fsb_scalar_create", C.c_int, _p, C.POINTER(C.c_int32))
    f("fsb_scalar_destroy", C.c_int, _p, C.c_int32)
    f("fsb_scalar_set", C.c_int, _p, C.c_int32, _dbl)
    f("fsb_scalar_get", C.c_int, _p, C.c_int32, _pd)
    f("fsb_vec_linear_sum_c", C.c_int, _p, Coef, _p, Coef, _p)
    f("fsb_vec_dot_opts", C.c_int, _p, _p, C.POINTER(RedOpts), _pi64)
    f("fsb_ctx_halt_arm", C.c_int, _p)
    f("fsb_ctx_halt_disarm", C.c_int, _p, C.POINTER(C.c_int))
    f("fsb_parcsr_create", C.c_int, _p, _i64, _pi64, _pi64, _pi64, _pd, C.POINTER(_p))
    f("fsb_parcsr_create_stencil", C.c_int, _p, C.c_int, _i64, _i64, _i64, _dbl, _dbl, C.POINTER(_p))
    f("fsb_parcsr_destroy", C.c_int, _p)
    for n in ("local_rows", "global_rows", "num_ghosts", "row_begin"):
        f(f"fsb_parcsr_{n}", _i64, _p)
    f("fsb_parcsr_local_nnz", _i64, _p, C.c_int)
    f("fsb_parcsr_info", _i64, _p, C.c_int)
    f("fsb_parcsr_download", C.c_int, _p, C.c_int, _pi64, _pi32, _pd)
    f("fsb_parcsr_download_colmap", C.c_int, _p, _pi64)
    f("fsb_parcsr_spmv", C.c_int, _p, _p, _p)
    f("fsb_parcsr_extract_dinv", C.c_int, _p, _p)
    f("fsb_parcsr_jacobi_relax", C.c_int, _p, _dbl, _i64, _p, _p, _p)
    f("fsb_parcsr_halo_exchange", C.c_int, _p, _p)


def check(rc: int) -> None:
    if rc != 0:
        raise FsbError(rc, lib().fsb_last_error().decode())


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(_pd)


def device_count() -> int:
    return lib().fsb_device_count()


STAT = {"launches": 0, "fused_statements": 1, "halo_exchanges": 2, "allreduces": 3, "host_syncs": 4,
        "unmatched_groups": 5, "wait_ns": 6, "flush_ns": 7, "jit_groups": 8, "bind_ns": 9, "ew_launch_ns": 10, "spmv_launch_ns": 11,
        "armed_hits": 13, "armed_misses": 14}
OPT = {"fusion": 0, "spmv_rows_per_cta": 1, "spmv_threads": 2, "trace": 3, "profile": 4, "reproducible": 5, "jit": 6,
       "timeline": 7, "spmv_dictionary": 8, "speculate": 9}


class Context:
    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1, unique_id: bytes | None = None):
        h = _p()
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        check(lib().fsb_ctx_create(device, rank, nranks, buf, C.byref(h)))
        self.h = h
        self.rank, self.nranks = rank, nranks

    @classmethod
    def group(cls, device: int, nranks: int) -> list["Context"]:
        """nranks contexts = the ranks of one group as threads of this process on one device (fsb_ctx_create_group);
        each must then be driven by its own host thread"""
        handles = (_p * nranks)()
        check(lib().fsb_ctx_create_group(device, nranks, handles))
        out = []
        for r in range(nranks):
            c = cls.__new__(cls)
            c.h, c.rank, c.nranks = _p(handles[r]), r, nranks
            out.append(c)
        return out

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().fsb_nccl_unique_id(buf))
        return buf.raw

    def close(self):
        if self.h:
            check(lib().fsb_ctx_destroy(self.h))
            self.h = None

    def sync(self):
        check(lib().fsb_ctx_sync(self.h))

    def flush(self):
        check(lib().fsb_ctx_flush(self.h))

    def flush_l2(self):
        check(lib().fsb_ctx_flush_l2(self.h))

    def event_record(self, slot: int):
        check(lib().fsb_ctx_event_record(self.h, slot))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        out = _dbl()
        check(lib().fsb_ctx_event_elapsed_ms(self.h, a, b, C.byref(out)))
        return out.value

    def timeline_read(self, max_slots: int = 65536):
        """the device-side timeline since the last read: uint64 array [launch, 16] (see fsb_ctx_timeline_read in fsb.h);
        needs set_option('timeline', capacity)"""
        buf = np.zeros((max_slots, 16), dtype=np.uint64)
        n = C.c_int64()
        check(lib().fsb_ctx_timeline_read(self.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), max_slots, C.byref(n)))
        return buf[:n.value].copy()

    def profile_read_split(self):
        """((diag ms, offd ms), (diag launches, offd launches)) since the last read."""
        ms = (C.c_double * 2)()
        cnt = (C.c_int64 * 2)()
        check(lib().fsb_ctx_profile_read_split(self.h, ms, cnt))
        return (ms[0], ms[1]), (cnt[0], cnt[1])

    def profile_read(self):
        """(total SpMV kernel ms, launches) since the last read; needs set_option('profile', 1)."""
        ms, cnt = _dbl(), _i64()
        check(lib().fsb_ctx_profile_read(self.h, C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    @property
    def stream(self) -> int:
        return int(lib().fsb_ctx_stream(self.h) or 0)

    def set_option(self, name: str, value: int):
        check(lib().fsb_ctx_set_option(self.h, OPT[name], value))

    def stat(self, name: str) -> int:
        v = _i64()
        check(lib().fsb_ctx_get_stat(self.h, STAT[name], C.byref(v)))
        return v.value

    def reset_stats(self):
        check(lib().fsb_ctx_reset_stats(self.h))

    def vector(self, n_owned: int, n_ghost: int = 0, data=None) -> "Vector":
        v = Vector(self, n_owned, n_ghost)
        if data is not None:
            v.upload(data)
        return v

    def wrap(self, device_ptr: int, n_owned: int, n_ghost: int = 0) -> "Vector":
        """a vector over caller-owned device memory (fsb_vec_wrap)"""
        h = _p()
        check(lib().fsb_vec_wrap(self.h, C.c_void_p(device_ptr), n_owned, n_ghost, C.byref(h)))
        return Vector(self, n_owned, n_ghost, handle=h)

    def get(self, token: int) -> float:
        out = _dbl()
        check(lib().fsb_red_get(self.h, token, C.byref(out)))
        return out.value

    # structured-grid vectors / operators (fsb.h "narray")
    def box_vector(self, extents, lo, hi) -> "Vector":
        e, l, h = (np.ascontiguousarray(a, dtype=np.int64) for a in (extents, lo, hi))
        hnd = _p()
        check(lib().fsb_vec_create_box(self.h, len(e), e.ctypes.data_as(_pi64), l.ctypes.data_as(_pi64),
                                       h.ctypes.data_as(_pi64), C.byref(hnd)))
        v = Vector(self, int(np.prod(h - l)), int(np.prod(e) - np.prod(h - l)), handle=hnd)
        v.box_extents = tuple(int(x) for x in e)
        return v

    def box_stencil(self, extents, lo, hi, center: float, off) -> "ParCSR":
        e, l, h = (np.ascontiguousarray(a, dtype=np.int64) for a in (extents, lo, hi))
        o = np.ascontiguousarray(off, dtype=np.float64)
        hnd = _p()
        check(lib().fsb_parcsr_create_box_stencil(self.h, len(e), e.ctypes.data_as(_pi64), l.ctypes.data_as(_pi64),
                                                  h.ctypes.data_as(_pi64), center, _dptr(o), C.byref(hnd)))
        return ParCSR(self, hnd)

    def box_fvm(self, extents, lo, hi, beta: float, alpha: float, vol: float, kface, a, bface) -> "ParCSR":
        """finite-volume diffusion operator -beta div(b grad u) + alpha vol a u from its coefficient fields (padded arrays)"""
        e, l, h = (np.ascontiguousarray(v, dtype=np.int64) for v in (extents, lo, hi))
        k = np.ascontiguousarray(kface, dtype=np.float64)
        av = np.ascontiguousarray(a, dtype=np.float64).ravel()
        bs = [np.ascontiguousarray(b, dtype=np.float64).ravel() for b in bface]
        assert av.size == int(np.prod(e)) and all(b.size == av.size for b in bs) and len(bs) == len(e) == k.size
        ptrs = (_pd * len(bs))(*[_dptr(b) for b in bs])
        hnd = _p()
        check(lib().fsb_parcsr_create_box_fvm(self.h, len(e), e.ctypes.data_as(_pi64), l.ctypes.data_as(_pi64),
                                              h.ctypes.data_as(_pi64), beta, alpha, vol, _dptr(k), _dptr(av), ptrs, C.byref(hnd)))
        return ParCSR(self, hnd)

    # device scalars (fsb.h "device scalars")
    def scalar(self, value: float | None = None) -> int:
        s = C.c_int32()
        check(lib().fsb_scalar_create(self.h, C.byref(s)))
        if value is not None:
            check(lib().fsb_scalar_set(self.h, s.value, value))
        return s.value

    def scalar_destroy(self, s: int): check(lib().fsb_scalar_destroy(self.h, s))
    def scalar_set(self, s: int, value: float): check(lib().fsb_scalar_set(self.h, s, value))

    def scalar_get(self, s: int) -> float:
        out = _dbl()
        check(lib().fsb_scalar_get(self.h, s, C.byref(out)))
        return out.value

    def halt_arm(self): check(lib().fsb_ctx_halt_arm(self.h))

    def halt_disarm(self) -> bool:
        v = C.c_int()
        check(lib().fsb_ctx_halt_disarm(self.h, C.byref(v)))
        return bool(v.value)


class Vector:
    def __init__(self, ctx: Context, n_owned: int, n_ghost: int = 0, handle=None):
        self.ctx = ctx
        if handle is None:
            handle = _p()
            check(lib().fsb_vec_create(ctx.h, n_owned, n_ghost, C.byref(handle)))
        self.h = handle
        self.n, self.n_ghost = n_owned, n_ghost

    def destroy(self):
        if self.h:
            check(lib().fsb_vec_destroy(self.h))
            self.h = None

    def upload(self, a, offset: int = 0):
        a = np.ascontiguousarray(a, dtype=np.float64)
        check(lib().fsb_vec_upload(self.h, _dptr(a), a.size, offset))
        return self

    def upload_all(self, a):
        """structured-grid vectors: the whole padded array, boundary layers included"""
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.size == self.n + self.n_ghost
        check(lib().fsb_vec_box_upload_all(self.h, _dptr(a)))
        return self

    def download_all(self) -> np.ndarray:
        out = np.empty(self.n + self.n_ghost, dtype=np.float64)
        check(lib().fsb_vec_box_download_all(self.h, _dptr(out)))
        return out

    def download(self, n: int | None = None, offset: int = 0) -> np.ndarray:
        n = self.n if n is None else n
        out = np.empty(n, dtype=np.float64)
        check(lib().fsb_vec_download(self.h, _dptr(out), n, offset))
        return out

    @property
    def device_ptr(self) -> int:
        return int(lib().fsb_vec_device_ptr(self.h) or 0)

    # element-wise (destination = self), mirroring vec::core
    def copy(self, x): check(lib().fsb_vec_copy(self.h, x.h))
    def set_scalar(self, a): check(lib().fsb_vec_set(self.h, a))
    def zero(self): check(lib().fsb_vec_set(self.h, 0.0))
    def scale(self, a, x=None): check(lib().fsb_vec_scale(self.h, a, (x or self).h))
    def add(self, x, y): check(lib().fsb_vec_add(self.h, x.h, y.h))
    def subtract(self, x, y): check(lib().fsb_vec_sub(self.h, x.h, y.h))
    def multiply(self, x, y): check(lib().fsb_vec_mul(self.h, x.h, y.h))
    def divide(self, x, y): check(lib().fsb_vec_div(self.h, x.h, y.h))
    def reciprocal(self, x): check(lib().fsb_vec_recip(self.h, x.h))
    def linear_sum(self, a, x, b, y): check(lib().fsb_vec_linear_sum(self.h, a, x.h, b, y.h))
    def axpy(self, a, x, y): check(lib().fsb_vec_axpy(self.h, a, x.h, y.h))
    def axpby(self, a, b, x): check(lib().fsb_vec_axpby(self.h, a, b, x.h))
    def abs(self, x): check(lib().fsb_vec_abs(self.h, x.h))
    def add_scalar(self, x, a): check(lib().fsb_vec_add_scalar(self.h, x.h, a))
    def set_random(self, seed): check(lib().fsb_vec_set_random(self.h, seed))
    def dump(self, prefix: str): check(lib().fsb_vec_dump(self.h, prefix.encode()))

    # reductions: *_token variants queue, the plain ones block
    def _tok(self, fn, *args) -> int:
        t = _i64()
        check(fn(*args, C.byref(t)))
        return t.value

    def dot_token(self, x): return self._tok(lib().fsb_vec_dot, self.h, x.h)

    def dot_opts_token(self, x, store=0, halt_mode=HALT_NEVER, halt_threshold=0.0, post=()):
        """post: [(op, dst, a, b), ...] scalar statements evaluated on the device once the value is stored"""
        o = RedOpts(store, halt_mode, halt_threshold)
        o.n_post = len(post)
        for k, (op, dst, a, b) in enumerate(post):
            o.post[k] = ScalarOp(SOP[op], dst, a, b)
        t = _i64()
        check(lib().fsb_vec_dot_opts(self.h, x.h, C.byref(o), C.byref(t)))
        return t.value

    def linear_sum_c(self, a, x, b, y):
        """a, b: float or (scale, num, den) naming device scalars"""
        ca = Coef(*a) if isinstance(a, tuple) else Coef(a, 0, 0)
        cb = Coef(*b) if isinstance(b, tuple) else Coef(b, 0, 0)
        check(lib().fsb_vec_linear_sum_c(self.h, ca, x.h, cb, y.h))
    def sumsq_token(self): return self._tok(lib().fsb_vec_sumsq, self.h)
    def dot(self, x): return self.ctx.get(self.dot_token(x))
    def l2norm(self): return float(np.sqrt(self.ctx.get(self.sumsq_token())))
    def l1norm(self): return self.ctx.get(self._tok(lib().fsb_vec_asum, self.h))
    def inf_norm(self): return self.ctx.get(self._tok(lib().fsb_vec_amax, self.h))
    def min(self): return self.ctx.get(self._tok(lib().fsb_vec_min, self.h))
    def max(self): return self.ctx.get(self._tok(lib().fsb_vec_max, self.h))
    def lp_norm(self, p: int): return self.ctx.get(self._tok(lib().fsb_vec_powsum, self.h, p)) ** (1.0 / p)

    def global_size(self) -> int:
        v = _i64()
        check(lib().fsb_vec_global_size(self.h, C.byref(v)))
        return v.value


class ParCSR:
    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle

    @classmethod
    def from_csr(cls, ctx: Context, n_global: int, row_part, rowptr, col, val) -> "ParCSR":
        row_part = np.ascontiguousarray(row_part, dtype=np.int64)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int64)
        val = np.ascontiguousarray(val, dtype=np.float64)
        h = _p()
        check(lib().fsb_parcsr_create(ctx.h, n_global, row_part.ctypes.data_as(_pi64), rowptr.ctypes.data_as(_pi64),
                                      col.ctypes.data_as(_pi64), _dptr(val), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def stencil(cls, ctx: Context, kind: int, nx: int, ny: int, nz: int, diag_shift=0.0, scale=1.0) -> "ParCSR":
        h = _p()
        check(lib().fsb_parcsr_create_stencil(ctx.h, kind, nx, ny, nz, diag_shift, scale, C.byref(h)))
        return cls(ctx, h)

    def destroy(self):
        if self.h:
            check(lib().fsb_parcsr_destroy(self.h))
            self.h = None

    @property
    def local_rows(self): return lib().fsb_parcsr_local_rows(self.h)
    @property
    def global_rows(self): return lib().fsb_parcsr_global_rows(self.h)
    @property
    def num_ghosts(self): return lib().fsb_parcsr_num_ghosts(self.h)
    @property
    def row_begin(self): return lib().fsb_parcsr_row_begin(self.h)
    def nnz(self, which=0): return lib().fsb_parcsr_local_nnz(self.h, which)

    INFO = {"window_format": 0, "row_blocks": 1, "window_x": 2, "fused_halo": 3, "wide_offsets": 4, "value_dictionary": 5, "dictionary_slots": 6}

    def info(self, key: str) -> int: return lib().fsb_parcsr_info(self.h, self.INFO[key])

    def vector(self, data=None) -> Vector:
        return self.ctx.vector(self.local_rows, self.num_ghosts, data)

    def download(self, which: int):
        n = self.local_rows
        nnz = self.nnz(which)
        rp = np.zeros(n + 1, dtype=np.int64)
        col = np.zeros(nnz, dtype=np.int32)
        val = np.zeros(nnz, dtype=np.float64)
        check(lib().fsb_parcsr_download(self.h, which, rp.ctypes.data_as(_pi64), col.ctypes.data_as(_pi32), _dptr(val)))
        return rp, col, val

    def colmap(self) -> np.ndarray:
        out = np.zeros(self.num_ghosts, dtype=np.int64)
        if out.size:
            check(lib().fsb_parcsr_download_colmap(self.h, out.ctypes.data_as(_pi64)))
        return out

    def spmv(self, x: Vector, y: Vector): check(lib().fsb_parcsr_spmv(self.h, x.h, y.h))
    def extract_dinv(self, d: Vector): check(lib().fsb_parcsr_extract_dinv(self.h, d.h))
    def jacobi_relax(self, omega, nrelax, b, x, tmp):
        check(lib().fsb_parcsr_jacobi_relax(self.h, omega, nrelax, b.h, x.h, tmp.h))
    def halo_exchange(self, x: Vector): check(lib().fsb_parcsr_halo_exchange(self.h, x.h))
