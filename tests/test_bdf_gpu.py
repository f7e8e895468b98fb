"""time_integrator::bdf of the C++ host layer over device vectors vs the reference's integrator
(golden step histories from oracle/_ref/refcheck_bdf, tests/golden/reference_bdf.json)."""
import json
import os

import numpy as np
import pytest

import oracle as O
from flecsolve_b200 import _lib as F
from flecsolve_b200 import host as H

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_bdf.json")))


def _topology(ctx, n):
    """any matrix with n rows gives vectors of length n on its topology"""
    rp = np.arange(n + 1, dtype=np.int64)
    return F.ParCSR.from_csr(ctx, n, [0, n], rp, np.arange(n, dtype=np.int64), np.ones(n))


@pytest.mark.parametrize("entry", GOLD["rate"], ids=lambda e: f"{e['case'][0]}-{e['case'][11]}")
def test_scalar_decay_matches_reference(ctx, entry):
    method, rtol, atol, dt0, dtmax, dtmin, tf, lam, ic, n, use_pi, controller, predictor = entry["case"]
    ref = entry["result"]
    A = _topology(ctx, n)
    opts = H.make_bdf_options(method=method, time_rtol=rtol, time_atol=atol, initial_dt=dt0, max_dt=dtmax, min_dt=dtmin,
                              final_time=tf, use_pi_controller=bool(use_pi), controller=controller, predictor=predictor,
                              error_scaling="fixed-resolution", norm="inf", max_steps=1000)
    res, dts, good, vals = H.bdf_rate(ctx, A, opts, lam, ic)
    assert (res.steps, res.rejects, res.attempts) == (ref["nsteps"], ref["rejects"], len(ref["steps"]))
    rdt = np.array([float.fromhex(s[0]) for s in ref["steps"]])
    rgood = np.array([s[1] for s in ref["steps"]])
    rval = np.array([float.fromhex(s[2]) for s in ref["steps"]])
    assert np.array_equal(good, rgood)
    # same accept/reject sequence; step sizes and iterates agree to rounding (the controller feeds
    # rounding differences back into dt, so long rejection-heavy runs drift a few ulps per attempt)
    assert np.allclose(dts, rdt, rtol=1e-8, atol=0)
    # candidates of rejected blow-up attempts are non-finite on both sides (inf here, nan there: fmax vs std::max)
    finite = np.isfinite(rval)
    assert np.array_equal(finite, np.isfinite(vals))
    assert np.allclose(vals[finite], rval[finite], rtol=1e-9, atol=0)
    assert abs(res.final_time - float.fromhex(ref["final_time"])) <= 1e-12
    ref_value = float.fromhex(ref["value"])
    if np.isfinite(ref_value):
        assert abs(res.value_max - ref_value) <= 1e-9 * abs(res.value_max)
    else:  # the run ended on a blown-up attempt (BDF6 is not zero-stable): non-finite on both sides
        assert not np.isfinite(res.value_max)
    if method == "BDF5" and lam == -1.0:  # the reference's KAT: 48 steps, 14 rejects, error < 1e-7
        assert res.steps == 48 and res.rejects == 14
        assert abs(ic * np.exp(lam * tf) - res.value_max) < 1e-7
    A.destroy()


def _heat_u0(dims, ic):
    nx, ny, nz = dims
    g = np.arange(nx * ny * nz)
    i, j, k = g % nx, (g // nx) % ny, g // (nx * ny)
    mid = lambda a, m: (5 * a >= 2 * m) & (5 * a < 3 * m)
    return np.where(mid(i, nx) & mid(j, ny) & mid(k, nz), ic, 0.0)


@pytest.mark.parametrize("entry", GOLD["heat"], ids=lambda e: f"{e['case'][0]}-{e['case'][10]}-{'x'.join(map(str, e['case'][8]))}")
def test_heat_equation_matches_reference(ctx, entry):
    method, rtol, atol, dt0, dtmax, dtmin, tf, ic, dims, length, solver, irtol, imax, kdim, maxatt = entry["case"]
    ref = entry["result"]
    h = length / (dims[0] + 1)
    A = F.ParCSR.stencil(ctx, 7, *dims, diag_shift=0.0, scale=-1.0 / (h * h))
    S = H.Session(ctx, A)
    opts = H.make_bdf_options(method=method, time_rtol=rtol, time_atol=atol, initial_dt=dt0, max_dt=dtmax, min_dt=dtmin,
                              final_time=tf, error_scaling="fixed-resolution", norm="inf", max_steps=1000,
                              max_attempts=maxatt)
    u, res, dts, good, iters = S.bdf_heat(_heat_u0(dims, ic), opts, solver=solver, rtol=irtol, maxiter=imax,
                                          max_krylov_dim=kdim, restart=True)
    rgood = np.array([s[1] for s in ref["steps"]])
    rdt = np.array([float.fromhex(s[0]) for s in ref["steps"]])
    riters = np.array([s[2] for s in ref["steps"]])
    assert (res.steps, res.rejects) == (ref["nsteps"], ref["rejects"])
    assert np.array_equal(good, rgood)
    assert np.allclose(dts, rdt, rtol=1e-6)
    assert np.all(np.abs(iters - riters) <= 1) and abs(res.inner_iterations - ref["inner_iterations"]) <= 2
    assert abs(res.value_max - float.fromhex(ref["value_max"])) <= 1e-6 * abs(res.value_max)
    assert abs(res.value_l2 - float.fromhex(ref["value_l2"])) <= 1e-6 * abs(res.value_l2)
    head = np.array([float.fromhex(v) for v in ref["u_head"]])
    n = u.size
    assert np.allclose(u[n // 2:n // 2 + 8], head, rtol=1e-5, atol=1e-9)
    # physics: heat is conserved up to the Dirichlet boundary loss and the maximum principle holds
    assert u.min() >= -1e-9 and u.max() <= ic
    S.close(); A.destroy()


RK = json.load(open(os.path.join(HERE, "golden", "reference_rk.json")))["cases"]


@pytest.mark.parametrize("entry", RK, ids=lambda e: f"rk{e['case'][0]}-{'fixed' if e['case'][7] else 'var'}-n{e['case'][10]}")
def test_explicit_integrators_match_reference(ctx, entry):
    order, dt0, dtmax, dtmin, tf, safety, atol, fixed, lam, ic, n = entry["case"]
    ref = entry["result"]
    A = _topology(ctx, n)
    res, dts, good, vals = H.rk_rate(ctx, A, order, lam, ic, dt0, dtmax, dtmin, tf, safety, atol, bool(fixed))
    assert res.steps == ref["nsteps"] and res.attempts == len(ref["steps"])
    assert np.array_equal(good, np.array([s[1] for s in ref["steps"]]))
    assert np.allclose(dts, [float.fromhex(s[0]) for s in ref["steps"]], rtol=1e-10)
    assert np.allclose(vals, [float.fromhex(s[2]) for s in ref["steps"]], rtol=1e-12)
    assert abs(res.final_time - float.fromhex(ref["final_time"])) <= 1e-12
    assert abs(res.value_max - float.fromhex(ref["value"])) <= 1e-12 * abs(res.value_max)
    A.destroy()
