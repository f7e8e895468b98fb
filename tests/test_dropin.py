"""Drop-in proof on the CPU: the REFERENCE's own solver / vector / integrator headers (unmodified, under
/root/reference) compiled against the policy classes of include/fsb_flecsolve/b200.hh (tests/dropin/dropin.cpp) and
linked against the sequential stand-in for the C ABI (tests/hostcheck/fsb_cpu_standin.cpp).  With eager serial
arithmetic behind the ABI the reference's templates must reproduce the reference's own serial runs BIT FOR BIT
(tests/golden/reference_solvers.json was produced by the same templates over vec::seq_vec) -- which shows that the
policy layer adds nothing of its own.  The device build of the same translation unit is tests/test_dropin_gpu.py."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O
from tests import dropin as DI
from tests import golden_util as G
from tests.dropin import build as DB
from tests.test_hostcheck import Standin
from flecsolve_b200 import host as H

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

pytestmark = pytest.mark.skipif(not os.path.isdir(DB.REF) and not os.path.exists(DB.STANDIN_OUT),
                                reason="needs /root/reference (or a prebuilt stand-in library)")


@pytest.fixture(scope="module")
def di():
    L = DI.declare(C.CDLL(DB.build_standin()))
    s = Standin(L)
    yield DI.Dropin(L), s
    s.close()


SOLVER_CASES = [e for e in G.load() if e["case"][2] != "spmv"]


@pytest.mark.parametrize("entry", SOLVER_CASES, ids=lambda e: "-".join(map(str, e["case"][:4])).replace(" ", ""))
def test_reference_templates_over_the_policies_reproduce_the_reference(di, entry):
    d, s = di
    (rp, col, val), M, b, x0, solver, precond, kw = G.problem(entry["case"])
    A = s.from_csr(len(rp) - 1, rp, col, val)
    x, info, hist = d.solve(s.ctx.h, A.h, b, x0, solver=solver, precond="dinv" if precond else None, history_cap=2000, **kw)
    ref = entry["info"]
    assert (info.status, info.iters, info.restarts) == (ref["status"], ref["iters"], ref["restarts"])
    assert np.array_equal(hist, G.history(entry))
    assert G.sha(x) == entry["x_sha256"]
    for k in ("res_norm_final", "sol_norm_initial", "sol_norm_final", "rhs_norm"):
        assert np.float32(getattr(info, k)) == np.float32(ref[k]), k


def test_reference_multivector_and_fcg(di):
    d, s = di
    rp, col, val = O.stencil_csr(7, 8, 7, 6)
    n = len(rp) - 1
    rp1, col1, val1 = O.stencil_csr(7, 8, 7, 6, 1e-3, 1.0)
    A0, A1 = s.from_csr(n, rp, col, val), s.from_csr(n, rp1, col1, val1)
    B0, B1 = sp.csr_matrix((val, col, rp)), sp.csr_matrix((val1, col1, rp1))
    rng = np.random.default_rng(3)
    b = np.concatenate([B0 @ rng.random(n), B1 @ rng.random(n)])
    for solver in ("cg", "bicgstab"):
        x, info, _ = d.solve_multi2(s.ctx.h, A0.h, A1.h, b, np.full(2 * n, 2.0), solver=solver, rtol=1e-9, maxiter=500)
        r = np.concatenate([b[:n] - B0 @ x[:n], b[n:] - B1 @ x[n:]])
        assert info.reason == "converged_rtol" and np.linalg.norm(r) <= 2e-9 * np.linalg.norm(b), solver
    x, info, _ = d.solve(s.ctx.h, A0.h, b[:n], np.zeros(n), solver="fcg", precond="dinv", rtol=1e-9, maxiter=500)
    assert info.reason == "converged_rtol" and np.linalg.norm(b[:n] - B0 @ x) <= 1.5e-9 * np.linalg.norm(b[:n])


def test_reference_vector_closed_forms(di):
    d, s = di
    for n in (32, 1001):
        out = d.vector_selftest(s.ctx.h, s.topology(n).h)
        assert len(out) == 22
        scale = max(1.0, float(n) ** 2 * 1e-6)
        assert np.all(np.abs(out) <= 1e-8 * scale * max(1.0, n ** 1.0)), out


BDF = json.load(open(os.path.join(HERE, "golden", "reference_bdf.json")))


@pytest.mark.parametrize("entry", BDF["heat"], ids=lambda e: f"{e['case'][0]}-{e['case'][10]}-{'x'.join(map(str, e['case'][8]))}")
def test_reference_integrator_over_the_policies(di, entry):
    d, s = di
    method, rtol, atol, dt0, dtmax, dtmin, tf, ic, dims, length, solver, irtol, imax, kdim, maxatt = entry["case"]
    ref = entry["result"]
    h = length / (dims[0] + 1)
    rp, col, val = O.stencil_csr(7, *dims, 0.0, -1.0 / (h * h))
    A = s.from_csr(len(rp) - 1, rp, col, val)
    opts = H.make_bdf_options(method=method, time_rtol=rtol, time_atol=atol, initial_dt=dt0, max_dt=dtmax, min_dt=dtmin,
                              final_time=tf, error_scaling="fixed-resolution", norm="inf", max_steps=1000, max_attempts=maxatt)
    nx, ny, nz = dims
    g = np.arange(nx * ny * nz)
    i, j, k = g % nx, (g // nx) % ny, g // (nx * ny)
    mid = lambda a, m: (5 * a >= 2 * m) & (5 * a < 3 * m)
    u0 = np.where(mid(i, nx) & mid(j, ny) & mid(k, nz), ic, 0.0)
    u, res, dts, good, iters = d.bdf_heat(s.ctx.h, A.h, u0, opts, solver=solver, rtol=irtol, maxiter=imax,
                                          max_krylov_dim=kdim, restart=True)
    # the same integrator code, the same arithmetic order: the step history is the reference's own
    assert (res.steps, res.rejects) == (ref["nsteps"], ref["rejects"])
    assert np.array_equal(good, np.array([st[1] for st in ref["steps"]]))
    assert np.array_equal(dts, np.array([float.fromhex(st[0]) for st in ref["steps"]]))
    assert np.array_equal(iters, np.array([st[2] for st in ref["steps"]]))
    assert res.value_max == float.fromhex(ref["value_max"])


def test_dropin_is_compiled_from_the_reference_tree_only():
    """the include path of the drop-in build names the stubs, /root/reference and include/ -- not this repo's host layer"""
    assert not any("flecsolve_b200" in d for d in DB.INCLUDE_DIRS) and DB.REF in DB.INCLUDE_DIRS
    assert open(os.path.join(ROOT, "tests", "dropin", "build.py")).read().count('"-I"') == 2  # both from INCLUDE_DIRS
    if not os.path.exists(DB.STANDIN_OUT):
        pytest.skip("library not built")
    syms = subprocess.run(["nm", "-C", DB.STANDIN_OUT], stdout=subprocess.PIPE, text=True).stdout
    # the reference's Krylov and integrator templates were instantiated on the device vector type
    for needle in ("flecsolve::op::cg<", "flecsolve::op::gmres<", "flecsolve::op::bicgstab<",
                   "flecsolve::time_integrator::bdf::integrator<"):
        assert any(needle in line and "flecsolve::vec::data::b200" in line for line in syms.splitlines()), needle
    L = DI.declare(C.CDLL(DB.STANDIN_OUT))
    assert b"unmodified from /root/reference" in L.fsbd_provenance()
