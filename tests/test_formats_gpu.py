"""Matrix Market file -> device matrix, and solvers chosen by a configuration file (krylov_factory),
against the direct calls."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O
from flecsolve_b200 import _lib as F
from flecsolve_b200 import host as H

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", ["general5.mtx", "sym6.mtx", "lap30_sym.mtx"])
def test_matrix_from_file_multiplies_like_the_host_csr(ctx, name):
    path = os.path.join(HERE, "golden", "mtx", name)
    nr, nc, sym, rp, col, val = H.read_mtx(path)
    A = H.mtx_create(ctx, path)
    assert A.local_rows == nr and A.nnz(0) == len(col)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(nr)
    xv, yv = A.vector(x), A.vector()
    A.spmv(xv, yv)
    assert np.array_equal(yv.download(), O.csr_spmv(rp, col, val, x))  # file order inside a row is kept
    xv.destroy(); yv.destroy(); A.destroy()


CFG = """[linear-solver]
type = {type}
[linear-solver.options]
maxiter = 400
use-zero-guess = false
rtol = 1e-8
{extra}
"""


@pytest.mark.parametrize("precond", [None, "dinv"])
@pytest.mark.parametrize("kind,extra,direct", [
    ("cg", "", dict(solver="cg")),
    ("gmres", "max-krylov-dim = 20\nrestart = true", dict(solver="gmres", max_krylov_dim=20, restart=True)),
    ("bicgstab", "", dict(solver="bicgstab")),
    ("cg-device", "lag = 1", dict(solver="cg_device", lag=1)),
    ("cg-sr", "lag = 1", dict(solver="cg_sr", lag=1)),
])
def test_solver_named_by_config_file(ctx, tmp_path, kind, extra, direct, precond):
    path = os.path.join(HERE, "golden", "mtx", "lap30_sym.mtx")
    nr, nc, sym, rp, col, val = H.read_mtx(path)
    A = H.mtx_create(ctx, path)
    S = H.Session(ctx, A)
    As = sp.csr_matrix((val, col, rp))
    b = As @ np.linspace(1, 2, nr)
    x0 = np.random.default_rng(7).random(nr)
    cfg = tmp_path / "solver.cfg"
    cfg.write_text(CFG.format(type=kind, extra=extra))
    x, info, hist = S.solve_config(str(cfg), "linear-solver", b, x0, precond=precond, history_cap=400)
    xd, dinfo, dhist = S.solve(b, x0, precond=precond, rtol=1e-8, maxiter=400, history_cap=400, **direct)
    assert info.reason == dinfo.reason == "converged_rtol"
    assert (info.iters, info.callbacks) == (dinfo.iters, dinfo.callbacks)
    assert np.array_equal(hist, dhist) and np.array_equal(x, xd)  # the same solver object, built two ways
    assert np.linalg.norm(b - As @ x) <= 2e-8 * np.linalg.norm(b)
    S.close(); A.destroy()
