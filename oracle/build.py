"""Builds the CPU restatement (oracle/oracle.cpp -> oracle/_build/liboracle.so).  Test infrastructure: only tests/,
__graft_entry__ (build + smoke checker) and bench.py's CPU legs use it; nothing under flecsolve_b200/ does."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
# the image exports CXX=/opt/gcc/bin/g++ (a wrapper without libgomp.spec); use the system compiler
CXX = os.environ.get("FSB_CXX") or ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")


def _walk(d: str, exts: tuple[str, ...]) -> list[str]:
    out = []
    for base, _, files in os.walk(d):
        out += [os.path.join(base, f) for f in files if f.endswith(exts)]
    return sorted(out)


def build_oracle(force: bool = False) -> str:
    out = os.path.join(HERE, "_build", "liboracle.so")
    srcs = [s for s in _walk(HERE, (".cpp",))
            if "/_ref/" not in s and "/_build/" not in s and "/stubs/" not in s and "/refcheck/" not in s]
    if not srcs:
        return out
    deps = srcs + _walk(HERE, (".h", ".hh"))
    stale = not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in deps)
    if force or stale:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cmd = [CXX, "-std=c++17", "-O3", "-march=native", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-Wall", "-o", out, *srcs]
        print("+", " ".join(cmd), flush=True)
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("oracle build failed")
    return out


if __name__ == "__main__":
    build_oracle(force="--force" in sys.argv)
