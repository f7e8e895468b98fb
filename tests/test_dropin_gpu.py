"""Drop-in proof on the device: tests/dropin/_build/libfsb_dropin.so is the REFERENCE's own, unmodified
flecsolve/solvers/{cg,gmres,bicgstab}.hh, vectors/{core,multi}.hh, operators/*.hh and time-integrators/bdf.hh
compiled against the policy classes of include/fsb_flecsolve/b200.hh and linked with libfsb.so (built in the
container where /root/reference lives; only the shared object travels).  Same bars as the tests of this repo's
own host layer (tests/test_solvers_gpu.py::test_device_matches_reference_golden, tests/test_bdf_gpu.py)."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O
from flecsolve_b200 import _lib as F
from flecsolve_b200 import host as H
from tests import dropin as DI
from tests import golden_util as G

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not DI.available(), reason="tests/dropin/_build/libfsb_dropin.so not built "
                                                            "(python tests/dropin/build.py where /root/reference exists)")]
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def di():
    return DI.Dropin(DI.lib())


def test_provenance(di):
    assert b"unmodified from /root/reference" in di.L.fsbd_provenance()


@pytest.mark.parametrize("entry", [e for e in G.load() if e["case"][2] != "spmv"],
                         ids=lambda e: "-".join(map(str, e["case"][:4])).replace(" ", ""))
def test_reference_solver_templates_on_the_device_match_the_reference_goldens(ctx, di, entry):
    (rp, col, val), M, b, x0, solver, precond, kw = G.problem(entry["case"])
    n = len(rp) - 1
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    info = entry["info"]
    launches0 = ctx.stat("launches")
    x, dinfo, hist = di.solve(ctx.h, A.h, b, x0, solver=solver, precond="dinv" if precond else None, history_cap=2000, **kw)
    assert ctx.stat("launches") > launches0  # the work ran in this library's kernels
    assert dinfo.status == info["status"]
    tol = 0.02 if solver != "bicgstab" else 0.1
    assert abs(dinfo.iters - info["iters"]) <= max(1, tol * info["iters"]), (dinfo.iters, info["iters"])
    ref_hist = G.history(entry)
    m = min(len(hist), len(ref_hist), 25)
    assert np.allclose(hist[:m], ref_hist[:m], rtol=1e-7)
    assert abs(np.linalg.norm(x) - float.fromhex(entry["x_norm2"])) <= 1e-6 * float.fromhex(entry["x_norm2"])
    # and the same run through this repo's own host layer: identical call sequence => identical kernels => same bits
    S = H.Session(ctx, A)
    x2, info2, hist2 = S.solve(b, x0, solver=solver, precond="dinv" if precond else None, history_cap=2000, **kw)
    assert (info2.status, info2.iters) == (dinfo.status, dinfo.iters)
    assert np.array_equal(hist, hist2) and np.array_equal(x, x2)
    S.close(); A.destroy()


def test_reference_cg_template_fuses_like_the_own_layer(ctx, di):
    """flecsolve's CG loop (cg.hh:91-131) through the policies: 4 launches per iteration at steady state"""
    nn = 32
    A = F.ParCSR.stencil(ctx, 7, nn, nn, nn)
    n = nn ** 3
    rng = np.random.default_rng(0)
    b = rng.standard_normal(n)
    x, info, _ = di.solve(ctx.h, A.h, b, np.zeros(n), solver="cg", precond="dinv", rtol=0.0, maxiter=60, ev_start=10, ev_stop=50)
    assert info.window_launches == 4 * 40
    A.destroy()


def test_reference_multivector_on_the_device(ctx, di):
    rp, col, val = O.stencil_csr(7, 12, 11, 10)
    n = len(rp) - 1
    rp1, col1, val1 = O.stencil_csr(7, 12, 11, 10, 1e-3, 1.0)
    A0 = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    A1 = F.ParCSR.from_csr(ctx, n, [0, n], rp1, col1, val1)
    B0, B1 = sp.csr_matrix((val, col, rp)), sp.csr_matrix((val1, col1, rp1))
    rng = np.random.default_rng(3)
    b = np.concatenate([B0 @ rng.random(n), B1 @ rng.random(n)])
    for solver in ("cg", "bicgstab"):
        x, info, hist = di.solve_multi2(ctx.h, A0.h, A1.h, b, np.full(2 * n, 2.0), solver=solver, rtol=1e-9, maxiter=500,
                                        history_cap=600)
        r = np.concatenate([b[:n] - B0 @ x[:n], b[n:] - B1 @ x[n:]])
        assert info.reason == "converged_rtol" and np.linalg.norm(r) <= 2e-9 * np.linalg.norm(b), solver
        x2, info2, hist2 = H.solve_multi2(ctx, A0, A1, b, np.full(2 * n, 2.0), solver=solver, rtol=1e-9, maxiter=500,
                                          history_cap=600)
        assert info2.iters == info.iters and np.array_equal(hist, hist2) and np.array_equal(x, x2), solver
    A0.destroy(); A1.destroy()


def test_reference_vector_closed_forms_on_the_device(ctx, di):
    for n in (32, 70001):
        rp = np.arange(n + 1, dtype=np.int64)
        A = F.ParCSR.from_csr(ctx, n, [0, n], rp, np.arange(n, dtype=np.int64), np.ones(n))
        out = di.vector_selftest(ctx.h, A.h)
        assert len(out) == 22
        # element-wise and min/max/inf checks are exact; sums of n terms of size <= n^2 to 1e-13 relative
        assert np.all(out[:15] == 0.0), out[:15]
        N = float(n)
        assert abs(out[15]) <= 1e-13 * N * N and abs(out[16]) <= 1e-13 * N ** 1.5 and out[17] == 0.0
        assert abs(out[18]) <= 1e-13 * N ** 3 and np.all(out[19:] == 0.0), out[15:]
        A.destroy()


BDF = json.load(open(os.path.join(HERE, "golden", "reference_bdf.json")))


@pytest.mark.parametrize("entry", BDF["heat"], ids=lambda e: f"{e['case'][0]}-{e['case'][10]}-{'x'.join(map(str, e['case'][8]))}")
def test_reference_integrator_on_the_device(ctx, di, entry):
    method, rtol, atol, dt0, dtmax, dtmin, tf, ic, dims, length, solver, irtol, imax, kdim, maxatt = entry["case"]
    ref = entry["result"]
    h = length / (dims[0] + 1)
    rp, col, val = O.stencil_csr(7, *dims, 0.0, -1.0 / (h * h))
    n = len(rp) - 1
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    opts = H.make_bdf_options(method=method, time_rtol=rtol, time_atol=atol, initial_dt=dt0, max_dt=dtmax, min_dt=dtmin,
                              final_time=tf, error_scaling="fixed-resolution", norm="inf", max_steps=1000, max_attempts=maxatt)
    nx, ny, nz = dims
    g = np.arange(n)
    i, j, k = g % nx, (g // nx) % ny, g // (nx * ny)
    mid = lambda a, m: (5 * a >= 2 * m) & (5 * a < 3 * m)
    u0 = np.where(mid(i, nx) & mid(j, ny) & mid(k, nz), ic, 0.0)
    u, res, dts, good, iters = di.bdf_heat(ctx.h, A.h, u0, opts, solver=solver, rtol=irtol, maxiter=imax,
                                           max_krylov_dim=kdim, restart=True)
    assert (res.steps, res.rejects) == (ref["nsteps"], ref["rejects"])
    assert np.array_equal(good, np.array([s[1] for s in ref["steps"]]))
    assert np.allclose(dts, [float.fromhex(s[0]) for s in ref["steps"]], rtol=1e-6)
    assert np.all(np.abs(iters - np.array([s[2] for s in ref["steps"]])) <= 1)
    assert abs(res.value_max - float.fromhex(ref["value_max"])) <= 1e-6 * abs(res.value_max)
    A.destroy()
