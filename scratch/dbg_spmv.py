import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
from flecsolve_b200 import _lib as F
ctx = F.Context(0)
for kind, dims in ((27, (24, 24, 24)), (27, (9, 10, 11)), (7, (40, 40, 40))):
    rp, col, val = O.stencil_csr(kind, *dims)
    n = len(rp) - 1
    x = np.random.default_rng(kind).standard_normal(n)
    ref = O.csr_spmv(rp, col, val, x)
    for name, A in (("host", F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)), ("gen", F.ParCSR.stencil(ctx, kind, *dims))):
        xv, yv = A.vector(x), A.vector()
        A.spmv(xv, yv)
        y = yv.download()
        bad = np.nonzero(y != ref)[0]
        print(kind, dims, name, "mismatches", bad.size, "of", n, "first", bad[:10], "maxrel", (np.abs(y - ref) / np.maximum(np.abs(ref), 1e-300)).max() if n else 0)
        if bad.size:
            r = bad[0]
            print("   row", r, "len", rp[r + 1] - rp[r], "y", y[r].hex(), "ref", ref[r].hex())
            # python sequential sum
            s = 0.0
            for k in range(rp[r], rp[r + 1]):
                s = s + val[k] * x[col[k]]
            print("   python seq", float(s).hex(), "rows mod:", bad[:20] % 151, bad[:20] % 128)
