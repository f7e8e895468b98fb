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


REFCHECK_BDF = os.path.join(ROOT, "oracle", "_ref", "refcheck_bdf")

# method, rtol, atol, initial_dt, max_dt, min_dt, final_time, lambda, ic, n, use_pi, controller, predictor
BDF_RATE_CASES = [
    ("BDF2", 1e-7, 1e-5, 0.5, 0.5, 0.001, 1.0, -1.0, 3.0, 1, 1, "PC.4.7", "leapfrog"),  # test/implicit.cfg [bdf-2]
    ("BDF5", 1e-7, 1e-5, 0.5, 0.5, 0.001, 1.0, -1.0, 3.0, 1, 1, "PC.4.7", "leapfrog"),  # test/implicit.cfg [bdf-5]
    ("BDF3", 1e-6, 1e-6, 0.25, 0.5, 0.001, 2.0, -2.0, 1.5, 7, 1, "H211b", "leapfrog"),
    ("BDF4", 1e-6, 1e-6, 0.25, 0.5, 0.001, 1.0, -0.5, 2.0, 3, 0, "Deadbeat", "leapfrog"),
    ("BDF6", 1e-7, 1e-6, 0.1, 0.5, 0.001, 1.0, -1.0, 1.0, 2, 1, "PC11", "leapfrog"),
    ("BE", 1e-3, 1e-4, 0.1, 0.5, 0.001, 1.0, -1.0, 1.0, 2, 1, "PC.4.7", "leapfrog"),
]
# heat: method, rtol, atol, initial_dt, max_dt, min_dt, final_time, ic, (nx, ny, nz), length, solver, inner_rtol, maxiter, kdim, max_attempts
BDF_HEAT_CASES = [
    ("BDF2", 1e-2, 1e-4, 1e-2, 1e-2, 1e-6, 0.1, 50.0, (16, 16, 16), 10.0, "gmres", 1e-6, 10000, 50, 0),
    ("BDF2", 1e-2, 1e-4, 1e-2, 1e-2, 1e-6, 0.1, 50.0, (12, 10, 8), 10.0, "cg", 1e-6, 10000, 50, 0),
    ("BDF3", 1e-3, 1e-5, 5e-3, 1e-2, 1e-6, 0.05, 50.0, (14, 14, 14), 10.0, "gmres", 1e-8, 10000, 30, 0),
]


def heat_scale(dims, length):
    """-alpha / h^2 with alpha = 1, h = length / (nx + 1): F = scale * stencil is the Laplacian."""
    h = length / (dims[0] + 1)
    return -1.0 / (h * h)


def run_bdf_rate(c):
    r = subprocess.run([REFCHECK_BDF, c[0], *[repr(float(v)) for v in c[1:9]], str(c[9]), str(c[10]), c[11], c[12]],
                       capture_output=True, text=True, check=True)
    return json.loads(r.stdout)


def run_bdf_heat(c):
    method, rtol, atol, dt0, dtmax, dtmin, tf, ic, dims, length, solver, irtol, imax, kdim, maxatt = c
    out = tempfile.mktemp(suffix=".bin")
    r = subprocess.run([REFCHECK_BDF, method, *[repr(float(v)) for v in (rtol, atol, dt0, dtmax, dtmin, tf, 0.0, ic)], "1",
                        "1", "PC.4.7", "leapfrog", "heat", *map(str, dims), repr(heat_scale(dims, length)), solver,
                        repr(float(irtol)), str(imax), str(kdim), str(maxatt), out], capture_output=True, text=True, check=True)
    d = json.loads(r.stdout)
    u = np.fromfile(out, dtype=np.float64)
    os.unlink(out)
    d["u_sha256"] = hashlib.sha256(u.tobytes()).hexdigest()
    d["u_head"] = [float(v).hex() for v in u[len(u) // 2:len(u) // 2 + 8]]
    return d


REFCHECK_RK = os.path.join(ROOT, "oracle", "_ref", "refcheck_rk")
# order, initial_dt, max_dt, min_dt, final_time, safety, atol, fixed, lambda, ic, n  (test/explicit.cfg [variable] / [fixed])
RK_CASES = [
    (23, 0.5, 0.5, 0.001, 1.0, 0.5, 1e-5, 0, -1.0, 3.0, 1),
    (45, 0.5, 0.5, 0.001, 1.0, 0.5, 1e-5, 0, -1.0, 3.0, 1),
    (23, 0.05, 0.5, 0.001, 1.0, 0.05, 1e-10, 1, -1.0, 3.0, 1),
    (45, 0.05, 0.5, 0.001, 1.0, 0.05, 1e-10, 1, -1.0, 3.0, 1),
    (23, 0.25, 0.5, 0.001, 2.0, 0.8, 1e-6, 0, -2.0, 1.5, 5),
    (45, 0.25, 0.5, 0.001, 2.0, 0.8, 1e-7, 0, -2.0, 1.5, 5),
]


def main_rk():
    out = {"generator": "tests/golden/make_golden.py (oracle/_ref/refcheck_rk = the reference's rk23.hh / rk45.hh + stubs)",
           "cases": []}
    for c in RK_CASES:
        r = subprocess.run([REFCHECK_RK, str(c[0]), *[repr(float(v)) for v in c[1:7]], str(c[7]), repr(float(c[8])),
                            repr(float(c[9])), str(c[10])], capture_output=True, text=True, check=True)
        d = json.loads(r.stdout)
        out["cases"].append({"case": list(c), "result": d})
        print("rk", c[0], "fixed" if c[7] else "variable", "steps", d["nsteps"], "attempts", len(d["steps"]))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_rk.json"), "w") as f:
        json.dump(out, f, indent=0)


def main_bdf():
    out = {"generator": "tests/golden/make_golden.py (oracle/_ref/refcheck_bdf = the reference's bdf.hh, bdf.cc, gmres.hh, "
                        "cg.hh, operator_adapter.hh + stubs)", "rate": [], "heat": []}
    for c in BDF_RATE_CASES:
        d = run_bdf_rate(c)
        out["rate"].append({"case": list(c), "result": d})
        print("bdf rate", c[0], "steps", d["nsteps"], "rejects", d["rejects"], "attempts", len(d["steps"]))
    for c in BDF_HEAT_CASES:
        d = run_bdf_heat(c)
        out["heat"].append({"case": [c[0], *c[1:8], list(c[8]), *c[9:]], "result": d})
        print("bdf heat", c[0], c[8], c[10], "steps", d["nsteps"], "rejects", d["rejects"], "inner", d["inner_iterations"])
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_bdf.json"), "w") as f:
        json.dump(out, f, indent=0)


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
    main_bdf()
    main_rk()
