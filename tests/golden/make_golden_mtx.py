"""Writes the Matrix Market fixtures under tests/golden/mtx/ and the CSR the REFERENCE'S reader makes
of them (oracle/_ref/refcheck_mtx, compiled from /root/reference/flecsolve/matrices/io/matrix_market.hh)
into tests/golden/reference_mtx.json.  Run in the build container: python tests/golden/make_golden_mtx.py"""
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "tests", "golden", "mtx")
REFCHECK = os.path.join(ROOT, "oracle", "_ref", "refcheck_mtx")


def write(name, banner, comments, n, m, entries):
    path = os.path.join(OUT, name)
    with open(path, "w") as f:
        f.write(banner + "\n")
        for c in comments:
            f.write("%" + c + "\n")
        f.write(f"{n} {m} {len(entries)}\n")
        for i, j, v in entries:
            f.write(f"{i} {j} {v}\n")
    return path


def fixtures():
    os.makedirs(OUT, exist_ok=True)
    files = []
    # general, rows out of order, values that are not floats exactly, a repeated entry
    files.append(write("general5.mtx", "%%MatrixMarket matrix coordinate real general", [" a comment"], 5, 5, [
        (3, 1, "0.1"), (1, 1, "4"), (1, 2, "-1.000000001"), (5, 5, "2.5e-3"), (2, 2, "3.14159265358979"),
        (2, 1, "-1"), (4, 4, "1e10"), (3, 3, "7"), (1, 2, "0.25"), (5, 1, "-0.333333333333"), (4, 2, "6.02e23")]))
    # symmetric lower triangle: off-diagonal entries are mirrored right behind the original
    files.append(write("sym6.mtx", "%%MatrixMarket matrix coordinate real symmetric", [], 6, 6, [
        (1, 1, "2"), (2, 1, "-1"), (2, 2, "2"), (3, 2, "-1"), (3, 3, "2"), (4, 3, "-1"), (4, 4, "2.0000001"),
        (5, 4, "-1"), (5, 5, "2"), (6, 5, "-0.1"), (6, 6, "2"), (6, 1, "0.7")]))
    # SPD 2-D Laplacian 6x5 written as a symmetric file in random order (a solvable system)
    rng = np.random.default_rng(0)
    nx, ny = 6, 5
    ent = []
    for j in range(ny):
        for i in range(nx):
            g = i + nx * j
            ent.append((g + 1, g + 1, "4.5"))
            if i > 0:
                ent.append((g + 1, g, "-1"))
            if j > 0:
                ent.append((g + 1, g - nx + 1, "-1.25"))
    ent = [ent[k] for k in rng.permutation(len(ent))]
    files.append(write("lap30_sym.mtx", "%%MatrixMarket matrix coordinate real symmetric", [" 6x5 grid"], nx * ny, nx * ny, ent))
    return files


def main():
    gold = {}
    for path in fixtures():
        r = subprocess.run([REFCHECK, path], capture_output=True, text=True, check=True)
        gold[os.path.basename(path)] = json.loads(r.stdout)
    with open(os.path.join(ROOT, "tests", "golden", "reference_mtx.json"), "w") as f:
        json.dump(gold, f, indent=0)
    print({k: (v["nrows"], v["nnz"], v["symmetric"]) for k, v in gold.items()})


if __name__ == "__main__":
    main()
