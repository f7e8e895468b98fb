"""Generates tests/golden/reference_solvers.json by RUNNING THE REFERENCE'S OWN CODE here:
oracle/_ref/refcheck is compiled from /root/reference's headers (flecsolve/matrices/seq.hh,
flecsolve/vectors/seq.hh, flecsolve/solvers/{cg,gmres,bicgstab}.hh) against stub FleCSI/Boost
headers (oracle/refcheck/).  Run in the build container:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFCHECK = os.path.join(ROOT, "oracle", "_ref", "refcheck")

CASES = [
    # kind, dims, solver, precond, rtol, maxiter, zero_guess, kdim, restart, bseed, xseed
    (5, (24, 24, 1), "cg", 0, 1e-9, 1000, 0, -1, 0, 0, 7),
    (7, (12, 11, 10), "cg", 0, 1e-9, 1000, 0, -1, 0, 0, 7),
    (7, (12, 11, 10), "cg", 1, 1e-9, 1000, 0, -1, 0, 0, 7),
    (7, (16, 16, 16), "cg", 1, 1e-9, 1000, 1, -1, 0, 3, 7),
    (27, (10, 9, 8), "cg", 1, 1e-9, 1000, 0, -1, 0, 0, 7),
    (7, (12, 11, 10), "gmres", 0, 1e-6, 100, 0, -1, 0, 0, 1),
    (7, (12, 11, 10), "gmres", 1, 1e-6, 100, 0, -1, 0, 0, 1),
    (27, (8, 8, 8), "gmres", 0, 1e-8, 200, 0, 10, 1, 0, 1),
    (5, (20, 20, 1), "gmres", 1, 1e-7, 300, 0, 15, 1, 0, 1),
    (7, (12, 11, 10), "bicgstab", 0, 1e-9, 500, 0, -1, 0, 0, 2),
    (27, (10, 9, 8), "bicgstab", 1, 1e-9, 500, 0, -1, 0, 0, 2),
    (5, (24, 24, 1), "bicgstab", 0, 1e-9, 500, 1, -1, 0, 5, 2),
    (7, (9, 8, 7), "spmv", 0, 0.0, 0, 0, -1, 0, 0, 1),
    (27, (7, 6, 5), "spmv", 0, 0.0, 0, 0, -1, 0, 0, 1),
    (5, (13, 9, 1), "spmv", 0, 0.0, 0, 0, -1, 0, 0, 1),
]


def run_case(c):
    kind, dims, solver, precond, rtol, maxiter, zero, kdim, restart, bseed, xseed = c
    out = tempfile.mktemp(suffix=".bin")
    r = subprocess.run([REFCHECK, str(kind), *map(str, dims), solver, str(precond), repr(float(rtol)), str(maxiter),
                        str(zero), str(kdim), str(restart), str(bseed), str(xseed), out],
                       capture_output=True, text=True, check=True)
    info = json.loads(r.stdout)
    raw = np.fromfile(out, dtype=np.float64)
    os.unlink(out)
    n = info["n"]
    b, x, hist = raw[:n], raw[n:2 * n], raw[2 * n:]
    return info, b, x, hist


def main():
    if not os.path.exists(REFCHECK):
        sys.path.insert(0, ROOT)
        import runpy
        runpy.run_path(os.path.join(ROOT, "oracle", "refcheck", "build.py"), run_name="__main__")
    golden = []
    for c in CASES:
        info, b, x, hist = run_case(c)
        entry = {"case": list(c[:1]) + [list(c[1])] + list(c[2:]), "info": info,
                 "history": [float(h).hex() for h in hist],
                 "x_sha256": hashlib.sha256(x.tobytes()).hexdigest(), "b_sha256": hashlib.sha256(b.tobytes()).hexdigest(),
                 "x_head": [float(v).hex() for v in x[:8]], "x_norm2": float(np.linalg.norm(x)).hex()}
        golden.append(entry)
        print(c[:3], "iters", info["iters"], "status", info["status"], "hist", len(hist))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_solvers.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py (oracle/_ref/refcheck = reference headers + stubs)",
                   "cases": golden}, f, indent=0)


if __name__ == "__main__":
    main()
