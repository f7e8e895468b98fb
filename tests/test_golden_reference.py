"""The oracle against golden vectors produced by the REFERENCE'S OWN solver templates and serial
SpMV (tests/golden/make_golden.py -> oracle/_ref/refcheck, built from /root/reference's headers).
This is what pins the oracle: bit-identical iterates, residual histories and solve_info."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle as O
from tests import golden_util as G

CASES = G.load()


@pytest.mark.parametrize("entry", CASES, ids=lambda e: "-".join(map(str, e["case"][:4])).replace(" ", ""))
def test_oracle_is_bit_identical_to_reference(entry):
    (rp, col, val), M, b, x0, solver, precond, kw = G.problem(entry["case"])
    assert G.sha(b) == entry["b_sha256"]  # same right-hand side (mt19937 + serial SpMV)
    info = entry["info"]
    if solver == "spmv":
        y = O.csr_spmv(rp, col, val, x0)
        assert G.sha(y) == entry["x_sha256"]  # matrices/seq.hh:178-194 bit for bit
        assert G.sha(O.ParCSR(rp, col, val, colours=1).spmv(x0)) == entry["x_sha256"]
        return
    fn = getattr(M, solver)
    x, oinfo, hist = fn(b, x0=x0, dinv=M.dinv() if precond else None, history_cap=2000, **kw)
    assert (oinfo.status, oinfo.iters, oinfo.restarts) == (info["status"], info["iters"], info["restarts"])
    ref_hist = G.history(entry)
    assert len(hist) == len(ref_hist) and np.array_equal(hist, ref_hist)  # every residual norm, every bit
    assert G.sha(x) == entry["x_sha256"]
    for k in ("res_norm_initial", "res_norm_final", "sol_norm_initial", "sol_norm_final", "rhs_norm"):
        if not (solver == "cg" and k == "res_norm_initial"):  # cg leaves it unset unless it exits early (cg.hh:77-82)
            assert np.float32(getattr(oinfo, k)) == np.float32(info[k]), k


def test_live_reference_binary_if_present():
    """When oracle/_ref/refcheck exists (built in the container, shipped with the snapshot) run one
    case live so the golden file cannot go stale silently."""
    exe = os.path.join(os.path.dirname(G.HERE), "oracle", "_ref", "refcheck")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/refcheck not built (needs /root/reference)")
    entry = next(e for e in CASES if e["case"][2] == "bicgstab")
    kind, dims, solver, precond, rtol, maxiter, zero, kdim, restart, bseed, xseed = entry["case"]
    out = tempfile.mktemp(suffix=".bin")
    r = subprocess.run([exe, str(kind), *map(str, dims), solver, str(precond), repr(float(rtol)), str(maxiter),
                        str(zero), str(kdim), str(restart), str(bseed), str(xseed), out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout)
    raw = np.fromfile(out, dtype=np.float64)
    os.unlink(out)
    assert info["iters"] == entry["info"]["iters"]
    assert G.sha(raw[info["n"]:2 * info["n"]]) == entry["x_sha256"]
