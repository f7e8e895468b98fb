"""Builds tests/hostcheck/_build/libhostcheck.so: flecsolve_b200/host/driver.cpp (the C++ host layer and its
drivers, unchanged) linked against the sequential CPU stand-in for the C ABI (fsb_cpu_standin.cpp) instead
of libfsb.so.  Test infrastructure: used by tests/test_hostcheck.py only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libhostcheck.so")


def build() -> str:
    srcs = [os.path.join(ROOT, "flecsolve_b200", "host", "driver.cpp"), os.path.join(HERE, "fsb_cpu_standin.cpp")]
    deps = list(srcs) + [os.path.join(ROOT, "include", "fsb.h")]
    for base, _, files in os.walk(os.path.join(ROOT, "flecsolve_b200", "include")):
        deps += [os.path.join(base, f) for f in files]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cxx = os.environ.get("FSB_CXX") or ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")
    # -ffp-contract=off: products and sums rounded separately, like the reference's baseline x86-64 build
    cmd = [cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wno-unused-local-typedefs",
           "-I", os.path.join(ROOT, "flecsolve_b200", "include"), "-I", os.path.join(ROOT, "include"), *srcs,
           "-Wl,-z,defs", "-o", OUT]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("hostcheck build failed:\n" + r.stdout[-6000:])
    return OUT


def build_example(name: str, source: str) -> str:
    """an example application (examples/) linked against the stand-in: its host logic can then run without a GPU"""
    out = os.path.join(HERE, "_build", name)
    srcs = [os.path.join(ROOT, "examples", source), os.path.join(HERE, "fsb_cpu_standin.cpp")]
    deps = list(srcs)
    for base, _, files in os.walk(os.path.join(ROOT, "flecsolve_b200", "include")):
        deps += [os.path.join(base, f) for f in files]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cxx = os.environ.get("FSB_CXX") or ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")
    cmd = [cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Wno-unused-local-typedefs", "-Wno-unused-variable",
           "-I", os.path.join(ROOT, "flecsolve_b200", "include"), "-I", os.path.join(ROOT, "include"), *srcs, "-o", out]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("example build failed:\n" + r.stdout[-6000:])
    return out


if __name__ == "__main__":
    print(build())
