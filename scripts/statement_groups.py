#!/usr/bin/env python
"""Which kernel launches do the solver / integrator templates turn into, and which of them have ahead-of-time
kernels?  Runs each driver on the CPU stand-in of the C ABI with call tracing on (tests/hostcheck) and replays the
trace through the deferred queue's grouping rules.  Writes profiles/r1_statement_groups.txt.  No GPU needed."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SCENARIO = r'''
import ctypes as C, sys, numpy as np, scipy.sparse as sp
sys.path.insert(0, {root!r})
import oracle as O
from flecsolve_b200 import host as H
from tests.hostcheck import build as HB
from tests.test_hostcheck import Standin
L = H.declare(C.CDLL(HB.build())); H._lib = L
s = Standin(L)
rp, col, val = O.stencil_csr(7, 10, 9, 8); n = len(rp) - 1
As = sp.csr_matrix((val, col, rp)); b = As @ np.linspace(1, 2, n); x0 = np.zeros(n)
S = H.Session(s.ctx, s.from_csr(n, rp, col, val))
{body}
'''

BODIES = {
    "cg + jacobi": "S.solve(b, x0, solver='cg', precond='dinv', rtol=1e-9, maxiter=500)",
    "cg_device + jacobi": "S.solve(b, x0, solver='cg_device', precond='dinv', rtol=1e-9, maxiter=500)",
    "fcg + jacobi": "S.solve(b, x0, solver='fcg', precond='dinv', rtol=1e-9, maxiter=500)",
    "gmres(20, restart) + jacobi": "S.solve(b, x0, solver='gmres', precond='dinv', rtol=1e-9, maxiter=300, max_krylov_dim=20, restart=True)",
    "gmres(20, restart)": "S.solve(b, x0, solver='gmres', rtol=1e-9, maxiter=300, max_krylov_dim=20, restart=True)",
    "bicgstab + jacobi": "S.solve(b, x0, solver='bicgstab', precond='dinv', rtol=1e-9, maxiter=500)",
    "bicgstab, two-component vec::multi": "A1 = s.from_csr(n, *O.stencil_csr(7, 10, 9, 8, 1e-3, 1.0)); H.solve_multi2(s.ctx, S.A, A1, np.concatenate([b, b]), np.full(2 * n, 2.0), solver='bicgstab', rtol=1e-9, maxiter=500)",
    "cg, two-component vec::multi": "A1 = s.from_csr(n, *O.stencil_csr(7, 10, 9, 8, 1e-3, 1.0)); H.solve_multi2(s.ctx, S.A, A1, np.concatenate([b, b]), np.full(2 * n, 2.0), solver='cg', rtol=1e-9, maxiter=500)",
    "heat: BDF2 (CN start) + gmres(50)": "h = 10.0 / 11; A2 = s.from_csr(n, *O.stencil_csr(7, 10, 9, 8, 0.0, -1.0 / (h * h))); S2 = H.Session(s.ctx, A2); S2.bdf_heat(np.where(np.arange(n) % 7 == 0, 50.0, 0.0), H.make_bdf_options(method='BDF2', time_rtol=1e-2, time_atol=1e-4, initial_dt=1e-2, max_dt=1e-2, min_dt=1e-6, final_time=0.1, error_scaling='fixed-resolution', norm='inf', max_steps=1000), solver='gmres', rtol=1e-6, maxiter=1000, max_krylov_dim=50, restart=True)",
    "scalar decay: BDF5": "H.bdf_rate(s.ctx, s.topology(64), H.make_bdf_options(method='BDF5', time_rtol=1e-7, time_atol=1e-5, initial_dt=0.5, max_dt=0.5, min_dt=1e-3, final_time=1.0, error_scaling='fixed-resolution', norm='inf', max_steps=1000), -1.0, 3.0)",
    "scalar decay: rk45": "H.rk_rate(s.ctx, s.topology(64), 45, -1.0, 3.0, 0.1, 0.5, 1e-4, 1.0, 0.9, 1e-8, False)",
    "scalar decay: rk23": "H.rk_rate(s.ctx, s.topology(64), 23, -1.0, 3.0, 0.1, 0.5, 1e-4, 1.0, 0.9, 1e-8, False)",
}


def main():
    from tests.hostcheck import groups as G
    out = ["# Kernel launches per driver, from a call trace of the CPU stand-in replayed through the deferred queue's grouping",
           "# rules (scripts/statement_groups.py; small 10x9x8 systems -- the group SHAPES do not depend on the size).",
           "# aot = an ahead-of-time instantiation of ew_program_kernel exists; otherwise the generic program kernel (or, with",
           "# FSB_JIT=1, a run-time instantiation) runs the group.  dev = device-resident coefficients / halt flag in play.", ""]
    for name, body in BODIES.items():
        with tempfile.NamedTemporaryFile(suffix=".trace", delete=False) as f:
            path = f.name
        env = dict(os.environ, FSB_STANDIN_TRACE=path)
        r = subprocess.run([sys.executable, "-c", SCENARIO.format(root=ROOT, body=body)], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            out.append(f"== {name}: FAILED\n{r.stderr[-800:]}")
            continue
        launches = G.replay(open(path).read().splitlines())
        os.unlink(path)
        total = sum(launches.values())
        generic = sum(c for (d, reg, dev), c in launches.items() if not reg)
        out.append(f"== {name}: {total} launches, {generic} without an ahead-of-time kernel ({100.0 * generic / max(total, 1):.1f} %)")
        for (d, reg, dev), c in sorted(launches.items(), key=lambda kv: -kv[1]):
            out.append(f"   {c:6d}  {'aot' if reg else '---'} {'dev' if dev else '   '}  {d}")
        out.append("")
    text = "\n".join(out) + "\n"
    open(os.path.join(ROOT, "profiles", "r1_statement_groups.txt"), "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
