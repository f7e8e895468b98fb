"""Helpers for the reference-generated golden vectors (tests/golden/reference_solvers.json)."""
import hashlib
import json
import os

import numpy as np

import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    return json.load(open(os.path.join(HERE, "golden", "reference_solvers.json")))["cases"]


def problem(case):
    """(oracle matrix, b, x0, solver name, kwargs) exactly as oracle/refcheck/refcheck.cpp builds them:
    b = A * set_random(bseed), x0 = set_random(xseed)."""
    kind, dims, solver, precond, rtol, maxiter, zero, kdim, restart, bseed, xseed = case
    rp, col, val = O.stencil_csr(kind, *dims)
    M = O.ParCSR(rp, col, val, colours=1)
    u = M.set_random(bseed)
    b = O.csr_spmv(rp, col, val, u)
    x0 = M.set_random(xseed)
    kw = dict(rtol=rtol, maxiter=maxiter, use_zero_guess=bool(zero))
    if solver == "gmres":
        kw.update(max_krylov_dim=kdim, restart=bool(restart))
    return (rp, col, val), M, b, x0, solver, bool(precond), kw


def history(entry):
    return np.array([float.fromhex(h) for h in entry["history"]])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()
