"""Scratch: one SpMV configuration (env-selected) timed; prints GB/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from flecsolve_b200 import _lib as F
kind = int(sys.argv[1]); nn = int(sys.argv[2]); nz = int(sys.argv[3]) if len(sys.argv) > 3 else nn
ctx = F.Context(0)
A = F.ParCSR.stencil(ctx, kind, nn, nn, nz)
N, nnz = A.local_rows, A.nnz(0)
p, w = A.vector(), A.vector()
p.set_scalar(1.0)
def timeit(fn, reps=30):
    for _ in range(3): fn()
    ctx.sync(); ctx.event_record(0)
    for _ in range(reps): fn()
    ctx.event_record(1)
    return ctx.event_elapsed_ms(0, 1) / reps
ms = timeit(lambda: (A.spmv(p, w), ctx.flush()))
byt = 12 * nnz + 4 * (N + 1) + 16 * N
cfg = {k: os.environ.get(k, '-') for k in ('FSB_SPMV_ROWS', 'FSB_SPMV_THREADS', 'FSB_SPMV_STAGES', 'FSB_SPMV_CTAS_PER_SM')}
print(f"{kind}pt {nn}x{nn}x{nz} rows={cfg['FSB_SPMV_ROWS']} thr={cfg['FSB_SPMV_THREADS']} st={cfg['FSB_SPMV_STAGES']} ctas={cfg['FSB_SPMV_CTAS_PER_SM']}: {ms:.4f} ms {byt/ms/1e6:.0f} GB/s")
