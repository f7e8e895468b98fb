"""ctypes binding of tests/dropin/_build/libfsb_dropin.so (the reference's own headers over the device policies;
see dropin.cpp).  The option / info structs are those of flecsolve_b200/host.py."""
import ctypes as C
import os

import numpy as np

from flecsolve_b200 import host as H

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libfsb_dropin.so")
_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)


def declare(L: C.CDLL) -> C.CDLL:
    L.fsbd_last_error.restype = C.c_char_p
    L.fsbd_provenance.restype = C.c_char_p
    L.fsbd_solve.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(H.Options), _pd, _pd, C.POINTER(H.Info), _pd, C.c_int]
    L.fsbd_solve_multi2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(H.Options), _pd, _pd, C.POINTER(H.Info),
                                    _pd, C.c_int]
    L.fsbd_bdf_heat.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(H.BdfOptions), C.POINTER(H.Options), _pd, _pd,
                                C.POINTER(H.BdfResult), _pd, _pi, _pi, C.c_int]
    L.fsbd_vector_selftest.argtypes = [C.c_void_p, C.c_void_p, _pd, C.c_int]
    return L


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        from flecsolve_b200 import _lib as F
        F.lib()  # libfsb.so (and the NCCL preload) first
        _lib = declare(C.CDLL(LIB_PATH))
    return _lib


class Dropin:
    """the entry points of dropin.cpp over a loaded library (the device build, or the CPU stand-in build in tests)"""

    def __init__(self, L: C.CDLL):
        self.L = L

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"dropin: {self.L.fsbd_last_error().decode()} (status {rc})")

    def solve(self, ctx_h, A_h, b, x0, history_cap=0, **kw):
        opts, info = H.make_options(**kw), H.Info()
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.array(x0, dtype=np.float64)
        hist = np.zeros(max(history_cap, 1))
        self._check(self.L.fsbd_solve(ctx_h, A_h, C.byref(opts), b.ctypes.data_as(_pd), x.ctypes.data_as(_pd), C.byref(info),
                                      hist.ctypes.data_as(_pd), history_cap))
        return x, info, hist[:min(info.callbacks, history_cap)]

    def solve_multi2(self, ctx_h, A0_h, A1_h, b, x0, history_cap=0, **kw):
        opts, info = H.make_options(**kw), H.Info()
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.array(x0, dtype=np.float64)
        hist = np.zeros(max(history_cap, 1))
        self._check(self.L.fsbd_solve_multi2(ctx_h, A0_h, A1_h, C.byref(opts), b.ctypes.data_as(_pd), x.ctypes.data_as(_pd),
                                             C.byref(info), hist.ctypes.data_as(_pd), history_cap))
        return x, info, hist[:min(info.callbacks, history_cap)]

    def bdf_heat(self, ctx_h, A_h, u0, bdf, cap=4096, **solver_kw):
        so = H.make_options(**solver_kw)
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        u = np.zeros_like(u0)
        res = H.BdfResult()
        dts, good, iters = np.zeros(cap), np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
        self._check(self.L.fsbd_bdf_heat(ctx_h, A_h, C.byref(bdf), C.byref(so), u0.ctypes.data_as(_pd), u.ctypes.data_as(_pd),
                                         C.byref(res), dts.ctypes.data_as(_pd), good.ctypes.data_as(_pi),
                                         iters.ctypes.data_as(_pi), cap))
        k = min(res.attempts, cap)
        return u, res, dts[:k], good[:k], iters[:k]

    def vector_selftest(self, ctx_h, A_h) -> np.ndarray:
        out = np.zeros(64)
        self._check(self.L.fsbd_vector_selftest(ctx_h, A_h, out.ctypes.data_as(_pd), len(out)))
        end = int(np.argmax(out == -1e300))
        return out[:end]
