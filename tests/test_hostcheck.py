"""The C++ host layer (flecsolve_b200/include/flecsolve/** driven through flecsolve_b200/host/driver.cpp) on the
CPU: the same driver source is linked against a sequential stand-in for the C ABI
(tests/hostcheck/fsb_cpu_standin.cpp -- test infrastructure, never part of the product) so that the solver /
integrator / factory / multi-vector / narray templates can be held against the reference-generated golden
vectors without a GPU.  With the stand-in following the reference's serial arithmetic, the Krylov templates
must reproduce the reference's own runs BIT FOR BIT (tests/golden/reference_solvers.json): every residual
norm, the final iterate, solve_info.  The device kernels are not exercised here (tests/*_gpu.py do that)."""
import ctypes as C
import json
import os
import types

import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O
from flecsolve_b200 import _lib as F
from flecsolve_b200 import host as H
from tests import golden_util as G
from tests.hostcheck import build as HB

HERE = os.path.dirname(os.path.abspath(__file__))
_pl = C.POINTER(C.c_int64)
_pd = C.POINTER(C.c_double)


class Standin:
    """handles of the stand-in library with the attributes host.py's wrappers look at (.h, sizes)"""

    def __init__(self, L):
        self.L = L
        L.fsb_ctx_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.fsb_parcsr_create.argtypes = [C.c_void_p, C.c_int64, _pl, _pl, _pl, _pd, C.POINTER(C.c_void_p)]
        L.fsb_parcsr_destroy.argtypes = [C.c_void_p]
        L.fsb_ctx_destroy.argtypes = [C.c_void_p]
        h = C.c_void_p()
        assert L.fsb_ctx_create(0, 0, 1, None, C.byref(h)) == 0
        self.ctx = types.SimpleNamespace(h=h)
        self._mats = []

    def from_csr(self, n, rp, col, val):
        rp, col = (np.ascontiguousarray(a, dtype=np.int64) for a in (rp, col))
        val = np.ascontiguousarray(val, dtype=np.float64)
        part = np.array([0, n], dtype=np.int64)
        h = C.c_void_p()
        assert self.L.fsb_parcsr_create(self.ctx.h, n, part.ctypes.data_as(_pl), rp.ctypes.data_as(_pl),
                                        col.ctypes.data_as(_pl), val.ctypes.data_as(_pd), C.byref(h)) == 0
        A = types.SimpleNamespace(h=h, local_rows=n, num_ghosts=0, ctx=self.ctx)
        self._mats.append(A)
        return A

    def topology(self, n):
        return self.from_csr(n, np.arange(n + 1), np.arange(n), np.ones(n))

    def close(self):
        for A in self._mats:
            self.L.fsb_parcsr_destroy(A.h)
        self.L.fsb_ctx_destroy(self.ctx.h)


@pytest.fixture()
def hc(monkeypatch):
    L = H.declare(C.CDLL(HB.build()))
    monkeypatch.setattr(H, "_lib", L)  # host.py's wrappers now call the stand-in build of the same driver
    s = Standin(L)
    yield s
    s.close()


SOLVER_CASES = [e for e in G.load() if e["case"][2] != "spmv"]


@pytest.mark.parametrize("entry", SOLVER_CASES, ids=lambda e: "-".join(map(str, e["case"][:4])).replace(" ", ""))
def test_krylov_templates_reproduce_the_reference_bit_for_bit(hc, entry):
    (rp, col, val), M, b, x0, solver, precond, kw = G.problem(entry["case"])
    A = hc.from_csr(len(rp) - 1, rp, col, val)
    S = H.Session(hc.ctx, A)
    x, info, hist = S.solve(b, x0, solver=solver, precond="dinv" if precond else None, history_cap=2000, **kw)
    ref = entry["info"]
    assert (info.status, info.iters, info.restarts) == (ref["status"], ref["iters"], ref["restarts"])
    ref_hist = G.history(entry)
    assert len(hist) == len(ref_hist) and np.array_equal(hist, ref_hist)
    assert G.sha(x) == entry["x_sha256"]
    for k in ("res_norm_initial", "res_norm_final", "sol_norm_initial", "sol_norm_final", "rhs_norm"):
        if not (solver == "cg" and k == "res_norm_initial"):
            assert np.float32(getattr(info, k)) == np.float32(ref[k]), k
    S.close()


@pytest.mark.parametrize("lag", [0, 2, 5])
@pytest.mark.parametrize("entry", [e for e in SOLVER_CASES if e["case"][2] == "cg"],
                         ids=lambda e: "-".join(map(str, e["case"][:4])).replace(" ", ""))
def test_cg_device_is_the_same_iteration_as_cg(hc, entry, lag):
    """with eager sequential arithmetic the device-scalar variant has nothing left to differ in: same bits"""
    (rp, col, val), M, b, x0, solver, precond, kw = G.problem(entry["case"])
    A = hc.from_csr(len(rp) - 1, rp, col, val)
    S = H.Session(hc.ctx, A)
    x, info, hist = S.solve(b, x0, solver="cg_device", precond="dinv" if precond else None, history_cap=2000, lag=lag, **kw)
    ref = entry["info"]
    assert (info.status, info.iters) == (ref["status"], ref["iters"])
    assert np.array_equal(hist, G.history(entry)) and G.sha(x) == entry["x_sha256"]
    assert np.float32(info.res_norm_final) == np.float32(ref["res_norm_final"])
    S.close()


@pytest.mark.parametrize("lag", [0, 2])
@pytest.mark.parametrize("entry", [e for e in SOLVER_CASES if e["case"][2] == "cg"],
                         ids=lambda e: "-".join(map(str, e["case"][:4])).replace(" ", ""))
def test_single_reduction_cg_follows_cg(hc, entry, lag):
    """Chronopoulos-Gear CG (solvers/cg_sr.hh): CG's iterates in exact arithmetic; in floating point the residual
    history agrees to ~1e-6 while the residual is above rounding, and the iteration count to +- 2"""
    (rp, col, val), M, b, x0, solver, precond, kw = G.problem(entry["case"])
    A = hc.from_csr(len(rp) - 1, rp, col, val)
    S = H.Session(hc.ctx, A)
    x, info, hist = S.solve(b, x0, solver="cg_sr", precond="dinv" if precond else None, history_cap=2000, lag=lag, **kw)
    ref = entry["info"]
    assert info.status == ref["status"] and abs(info.iters - ref["iters"]) <= 2
    rh = G.history(entry)
    m = min(len(hist), len(rh)) * 2 // 3
    assert np.allclose(hist[:m], rh[:m], rtol=1e-6)
    As = sp.csr_matrix((val, col, rp))
    assert np.linalg.norm(b - As @ x) <= 2 * np.float32(kw["rtol"]) * np.linalg.norm(b)
    S.close()


def test_fcg_and_maxiter_and_early_exit(hc):
    rp, col, val = O.stencil_csr(7, 9, 8, 7)
    n = len(rp) - 1
    A = hc.from_csr(n, rp, col, val)
    S = H.Session(hc.ctx, A)
    As = sp.csr_matrix((val, col, rp))
    b = As @ np.linspace(1, 2, n)
    x, info, _ = S.solve(b, np.zeros(n), solver="fcg", precond="dinv", rtol=1e-9, maxiter=500)
    assert info.reason == "converged_rtol" and np.linalg.norm(b - As @ x) <= 1.5e-9 * np.linalg.norm(b)
    for solver in ("cg", "cg_device", "bicgstab", "fcg"):
        _, info, hist = S.solve(b, np.zeros(n), solver=solver, rtol=1e-12, maxiter=3, history_cap=10)
        assert info.reason == "diverged_iters" and info.iters == 0 and len(hist) == 3, solver
        _, info, _ = S.solve(b, x, solver=solver, rtol=1e-6, maxiter=100)
        assert info.reason == "converged_rtol" and info.iters == 0 and info.callbacks == 0, solver
    S.close()


BDF = json.load(open(os.path.join(HERE, "golden", "reference_bdf.json")))


@pytest.mark.parametrize("entry", BDF["rate"], ids=lambda e: f"{e['case'][0]}-{e['case'][11]}")
def test_bdf_template_matches_reference(hc, entry):
    method, rtol, atol, dt0, dtmax, dtmin, tf, lam, ic, n, use_pi, controller, predictor = entry["case"]
    ref = entry["result"]
    opts = H.make_bdf_options(method=method, time_rtol=rtol, time_atol=atol, initial_dt=dt0, max_dt=dtmax, min_dt=dtmin,
                              final_time=tf, use_pi_controller=bool(use_pi), controller=controller, predictor=predictor,
                              error_scaling="fixed-resolution", norm="inf", max_steps=1000)
    res, dts, good, vals = H.bdf_rate(hc.ctx, hc.topology(n), opts, lam, ic)
    assert (res.steps, res.rejects, res.attempts) == (ref["nsteps"], ref["rejects"], len(ref["steps"]))
    assert np.array_equal(good, np.array([s[1] for s in ref["steps"]]))
    rdt = np.array([float.fromhex(s[0]) for s in ref["steps"]])
    rval = np.array([float.fromhex(s[2]) for s in ref["steps"]])
    assert np.allclose(dts, rdt, rtol=1e-8, atol=0)
    finite = np.isfinite(rval)
    assert np.array_equal(finite, np.isfinite(vals)) and np.allclose(vals[finite], rval[finite], rtol=1e-9, atol=0)
    assert abs(res.final_time - float.fromhex(ref["final_time"])) <= 1e-12


RK = json.load(open(os.path.join(HERE, "golden", "reference_rk.json")))["cases"]


@pytest.mark.parametrize("entry", RK, ids=lambda e: f"rk{e['case'][0]}-{'fixed' if e['case'][7] else 'var'}-n{e['case'][10]}")
def test_rk_templates_match_reference(hc, entry):
    order, dt0, dtmax, dtmin, tf, safety, atol, fixed, lam, ic, n = entry["case"]
    ref = entry["result"]
    res, dts, good, vals = H.rk_rate(hc.ctx, hc.topology(n), order, lam, ic, dt0, dtmax, dtmin, tf, safety, atol, bool(fixed))
    assert res.steps == ref["nsteps"] and res.attempts == len(ref["steps"])
    assert np.array_equal(good, np.array([s[1] for s in ref["steps"]]))
    assert np.allclose(dts, [float.fromhex(s[0]) for s in ref["steps"]], rtol=1e-10)
    assert np.allclose(vals, [float.fromhex(s[2]) for s in ref["steps"]], rtol=1e-12)
    assert abs(res.final_time - float.fromhex(ref["final_time"])) <= 1e-12


@pytest.mark.parametrize("entry", BDF["heat"], ids=lambda e: f"{e['case'][0]}-{e['case'][10]}-{'x'.join(map(str, e['case'][8]))}")
def test_heat_driver_matches_reference(hc, entry):
    method, rtol, atol, dt0, dtmax, dtmin, tf, ic, dims, length, solver, irtol, imax, kdim, maxatt = entry["case"]
    ref = entry["result"]
    h = length / (dims[0] + 1)
    rp, col, val = O.stencil_csr(7, *dims, 0.0, -1.0 / (h * h))
    S = H.Session(hc.ctx, hc.from_csr(len(rp) - 1, rp, col, val))
    opts = H.make_bdf_options(method=method, time_rtol=rtol, time_atol=atol, initial_dt=dt0, max_dt=dtmax, min_dt=dtmin,
                              final_time=tf, error_scaling="fixed-resolution", norm="inf", max_steps=1000, max_attempts=maxatt)
    nx, ny, nz = dims
    g = np.arange(nx * ny * nz)
    i, j, k = g % nx, (g // nx) % ny, g // (nx * ny)
    mid = lambda a, m: (5 * a >= 2 * m) & (5 * a < 3 * m)
    u0 = np.where(mid(i, nx) & mid(j, ny) & mid(k, nz), ic, 0.0)
    u, res, dts, good, iters = S.bdf_heat(u0, opts, solver=solver, rtol=irtol, maxiter=imax, max_krylov_dim=kdim, restart=True)
    assert (res.steps, res.rejects) == (ref["nsteps"], ref["rejects"])
    assert np.array_equal(good, np.array([s[1] for s in ref["steps"]]))
    assert np.allclose(dts, [float.fromhex(s[0]) for s in ref["steps"]], rtol=1e-6)
    assert np.all(np.abs(iters - np.array([s[2] for s in ref["steps"]])) <= 1)
    assert abs(res.value_max - float.fromhex(ref["value_max"])) <= 1e-6 * abs(res.value_max)
    S.close()


def test_multi_vector_subset_selftest_and_adapter(hc):
    rp, col, val = O.stencil_csr(7, 8, 7, 6)
    n = len(rp) - 1
    rp1, col1, val1 = O.stencil_csr(7, 8, 7, 6, 1e-3, 1.0)
    A0, A1 = hc.from_csr(n, rp, col, val), hc.from_csr(n, rp1, col1, val1)
    B0, B1 = sp.csr_matrix((val, col, rp)), sp.csr_matrix((val1, col1, rp1))
    rng = np.random.default_rng(3)
    b = np.concatenate([B0 @ rng.random(n), B1 @ rng.random(n)])
    for solver in ("cg", "bicgstab"):
        x, info, _ = H.solve_multi2(hc.ctx, A0, A1, b, np.full(2 * n, 2.0), solver=solver, rtol=1e-9, maxiter=500)
        r = np.concatenate([b[:n] - B0 @ x[:n], b[n:] - B1 @ x[n:]])
        assert info.reason == "converged_rtol" and np.linalg.norm(r) <= 2e-9 * np.linalg.norm(b), solver
    # CG bound to one variable of a two-component vector iterates exactly like the single-vector solve
    xs, sinfo = H.solve_subset(hc.ctx, A0, 1, np.concatenate([np.zeros(n), b[:n]]), np.zeros(2 * n), solver="cg", rtol=1e-9,
                               maxiter=500)
    S = H.Session(hc.ctx, A0)
    x1, info1, _ = S.solve(b[:n], np.zeros(n), solver="cg", rtol=1e-9, maxiter=500)
    assert sinfo.iters == info1.iters and np.array_equal(xs[n:], x1) and not xs[:n].any()
    assert np.abs(S.vector_selftest()[:11]).max() <= 1e-12  # vectors/test/flecsi_vector.cc closed forms
    xin = rng.standard_normal(n)
    assert np.array_equal(S.adapter_apply(0.25, xin), -0.25 * (B0 @ xin) + xin) or np.allclose(
        S.adapter_apply(0.25, xin), xin - 0.25 * (B0 @ xin), rtol=1e-13)
    S.close()


def test_config_chosen_solver_and_narray_example(hc, tmp_path):
    rp, col, val = O.stencil_csr(5, 12, 11, 1)
    n = len(rp) - 1
    S = H.Session(hc.ctx, hc.from_csr(n, rp, col, val))
    b = sp.csr_matrix((val, col, rp)) @ np.linspace(1, 2, n)
    x0 = np.random.default_rng(7).random(n)
    for kind, extra, direct in (("cg", "", dict(solver="cg")), ("bicgstab", "", dict(solver="bicgstab")),
                                ("gmres", "max-krylov-dim = 20\nrestart = true", dict(solver="gmres", max_krylov_dim=20, restart=True)),
                                ("cg-device", "lag = 1", dict(solver="cg_device", lag=1)),
                                ("cg-sr", "lag = 1", dict(solver="cg_sr", lag=1))):
        cfg = tmp_path / f"{kind}.cfg"
        cfg.write_text(f"[ls]\ntype = {kind}\n[ls.options]\nmaxiter = 300\nuse-zero-guess = false\nrtol = 1e-8\n{extra}\n")
        x, info, hist = S.solve_config(str(cfg), "ls", b, x0, precond="dinv", history_cap=300)
        xd, dinfo, dhist = S.solve(b, x0, precond="dinv", rtol=1e-8, maxiter=300, history_cap=300, **direct)
        assert info.reason == "converged_rtol" and info.iters == dinfo.iters, kind
        assert np.array_equal(hist, dhist) and np.array_equal(x, xd), kind
    S.close()
    # examples/poisson on its narray mesh vs the same system as a flat CSR matrix: identical arithmetic here
    m = 24
    u, info, hist = H.poisson_narray(hc.ctx, m, seed=7, solver="cg", rtol=1e-9, maxiter=1000, history_cap=1000)
    rp, col, val = O.stencil_csr(5, m, m, 1)
    S = H.Session(hc.ctx, hc.from_csr(m * m, rp, col, val))
    hgrid = 1.0 / (m + 1)
    xs = (np.arange(m) + 1) * hgrid
    X, Y = np.meshgrid(xs, xs, indexing="xy")
    exact = (np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y)).ravel()
    x0 = O.ParCSR(rp, col, val, colours=1).set_random(7)
    xf, finfo, fhist = S.solve(8 * np.pi ** 2 * exact * hgrid * hgrid, x0, solver="cg", rtol=1e-9, maxiter=1000, history_cap=1000)
    assert info.reason == "converged_rtol" and info.iters == finfo.iters
    assert np.allclose(hist, fhist, rtol=1e-10) and np.abs(u - xf).max() <= 1e-10
    assert np.abs(u - exact).max() < 8e-3
    S.close()


def test_example_applications_run_on_the_stand_in(tmp_path):
    """examples/poisson and examples/heat_equation: flecsolve-shaped user programs (read_config, krylov_factory,
    bdf::integrator, narray mesh); linked against the stand-in their host logic runs here end to end"""
    import re
    import subprocess
    ex = os.path.join(os.path.dirname(HERE), "examples")
    exe = HB.build_example("poisson_standin", os.path.join("poisson", "poisson.cc"))
    r = subprocess.run([exe, "40", os.path.join(ex, "poisson", "poisson.cfg")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"converged after (\d+) iterations.*max error.*= ([0-9.e+-]+)", r.stdout)
    assert m and 50 < int(m.group(1)) < 300 and float(m.group(2)) < 3e-3, r.stdout
    exe = HB.build_example("implicit_standin", os.path.join("heat_equation", "implicit.cc"))
    r = subprocess.run([exe, "16", os.path.join(ex, "heat_equation", "implicit.cfg")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"(\d+) steps \((\d+) attempts, (\d+) rejected\), max u = ([0-9.]+), heat ([0-9.]+) -> ([0-9.]+)", r.stdout)
    assert m, r.stdout
    assert int(m.group(1)) >= 10 and 0.0 < float(m.group(4)) <= 50.0 and float(m.group(6)) <= float(m.group(5)) * (1 + 1e-9)


def test_multivector_closed_forms_of_the_reference_test(hc):
    """vectors/test/flecsi_multivector.cc:84-241 on a four-component vec::multi, n = 32 as there"""
    S = H.Session(hc.ctx, hc.topology(32))
    out = S.multivector_selftest()
    assert np.all(out[:11] < 1e-8), out[:11]  # the reference's tolerance
    assert out[11] == -7  # tmp.min() after tmp = y - 7: component 0 at gid 0
    assert np.all(out[12:18] == 1.0), out[12:18]  # combined reductions are exactly the combination of the components'
    S.close()
