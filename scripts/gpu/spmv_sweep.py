"""Scratch: one SpMV configuration (env-selected) timed; prints GB/s (algorithmic bytes: 12 nnz + 4 (N+1) + 16 N).
usage: spmv_sweep.py <kind> <n> [nz] [mode]   mode: plain | dotx (dot with x, CG's <Ap,p>) | doty (sum y^2) | dotu (dot with a third vector)
       | jacobi (one weighted-Jacobi sweep)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from flecsolve_b200 import _lib as F
kind = int(sys.argv[1]); nn = int(sys.argv[2]); nz = int(sys.argv[3]) if len(sys.argv) > 3 else nn
mode = sys.argv[4] if len(sys.argv) > 4 else "plain"
ctx = F.Context(0)
A = F.ParCSR.stencil(ctx, kind, nn, nn, nz)
N, nnz = A.local_rows, A.nnz(0)
p, w, u = A.vector(), A.vector(), A.vector()
p.set_scalar(1.0); u.set_scalar(2.0)
def timeit(fn, reps=30):
    for _ in range(3): fn()
    ctx.sync(); ctx.event_record(0)
    for _ in range(reps): fn()
    ctx.event_record(1)
    return ctx.event_elapsed_ms(0, 1) / reps
def step():
    if mode == "jacobi":
        A.jacobi_relax(2 / 3, 1, u, p, w)
        return
    A.spmv(p, w)
    if mode == "dotx": w.dot_token(p)
    elif mode == "doty": w.sumsq_token()
    elif mode == "dotu": w.dot_token(u)
    ctx.flush()
ms = timeit(step)
byt = 12 * nnz + 4 * (N + 1) + 16 * N
cfg = {k: os.environ.get(k, '-') for k in ('FSB_SPMV_ROWS', 'FSB_SPMV_THREADS', 'FSB_SPMV_STAGES', 'FSB_SPMV_CTAS_PER_SM', 'FSB_SPMV_WINDOW')}
print(f"{kind}pt {nn}x{nn}x{nz} {mode:6s} rows={cfg['FSB_SPMV_ROWS']} st={cfg['FSB_SPMV_STAGES']} win={cfg['FSB_SPMV_WINDOW']}: {ms:.4f} ms {byt/ms/1e6:.0f} GB/s")
