"""Build recipes for the native parts (run here on CPU; nvcc cross-compiles sm_100a).

  libfsb.so       CUDA kernels + C ABI (include/fsb.h)               nvcc
  libfsb_host.so  C++ host layer (flecsolve-shaped headers) drivers   g++
  oracle/_build/liboracle.so   CPU restatement (test infrastructure)  g++

Everything is built in-tree so the shared objects travel to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flecsolve_b200")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# the image exports CXX=/opt/gcc/bin/g++ (a wrapper without libgomp.spec); use the system compiler
CXX = os.environ.get("FSB_CXX") or ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")

NVCC_FLAGS = [
    "-std=c++20",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "--expt-relaxed-constexpr",
    "--threads", "4",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _walk(d: str, exts: tuple[str, ...]) -> list[str]:
    out = []
    for base, _, files in os.walk(d):
        out += [os.path.join(base, f) for f in files if f.endswith(exts)]
    return sorted(out)


def _run(cmd: list[str]) -> None:
    print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError(f"build failed: {' '.join(cmd[:3])} ...")
    if r.stdout.strip():
        print(r.stdout)


def lib_path() -> str:
    return os.path.join(PKG, "libfsb.so")


def host_lib_path() -> str:
    return os.path.join(PKG, "libfsb_host.so")


def build_lib(force: bool = False, verbose_ptxas: bool = False) -> str:
    out = lib_path()
    srcs = [os.path.join(CSRC, f) for f in ("capi.cu", "fuser.cu", "spmv.cu", "parcsr.cu", "halo.cu")]
    deps = _walk(CSRC, (".cu", ".cuh", ".h")) + [os.path.join(ROOT, "include", "fsb.h")]
    if force or _newer(out, deps):
        flags = list(NVCC_FLAGS)
        if verbose_ptxas:
            flags += ["-Xptxas", "-v"]
        _run([NVCC, *flags, "-o", out, *srcs, "-lnccl"])
    return out


def build_host(force: bool = False) -> str:
    out = host_lib_path()
    srcs = _walk(HOST, (".cpp",))
    if not srcs:
        return out
    deps = srcs + _walk(os.path.join(PKG, "include"), (".hh", ".h")) + [os.path.join(ROOT, "include", "fsb.h")]
    if force or _newer(out, deps) or _newer(out, [lib_path()]):
        _run([CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-Wno-unused-local-typedefs",
              "-I", os.path.join(PKG, "include"), "-I", os.path.join(ROOT, "include"),
              "-o", out, *srcs, "-L", PKG, "-lfsb", "-Wl,-rpath,$ORIGIN"])
    return out


def build_oracle(force: bool = False) -> str:
    odir = os.path.join(ROOT, "oracle")
    out = os.path.join(odir, "_build", "liboracle.so")
    srcs = _walk(odir, (".cpp",))
    srcs = [s for s in srcs if "/_ref/" not in s and "/_build/" not in s and "/stubs/" not in s and "/refcheck/" not in s]
    if not srcs:
        return out
    deps = srcs + _walk(odir, (".h", ".hh"))
    if force or _newer(out, deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        _run([CXX, "-std=c++17", "-O3", "-march=native", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
              "-Wall", "-o", out, *srcs])
    return out


def build_all(force: bool = False) -> None:
    build_lib(force)
    build_host(force)
    build_oracle(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
