"""Builds tests/dropin/_build/libfsb_dropin.so: the reference's own solver / vector / integrator headers
(unmodified, where they lie under /root/reference) instantiated on device vectors through the policy classes
of include/fsb_flecsolve/b200.hh and linked against libfsb.so.  Container only (needs /root/reference); the
shared object is git-ignored and travels to the GPU box with the snapshot, where nothing reads /root/reference."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
OUT = os.path.join(HERE, "_build", "libfsb_dropin.so")
STUBS = os.path.join(ROOT, "oracle", "refcheck", "stubs")
# include path: the stubs, the REFERENCE tree, the C ABI + policy header -- not this repo's own host layer
INCLUDE_DIRS = [STUBS, REF, os.path.join(ROOT, "include")]
# translation units of the reference itself that the integrator needs
REF_UNITS = ["flecsolve/time-integrators/bdf.cc", "flecsolve/time-integrators/bdf_parameters.cc", "flecsolve/vectors/util.cc"]


STANDIN_OUT = os.path.join(HERE, "_build", "libfsb_dropin_standin.so")


def build_standin() -> str:
    """the same translation unit linked against the sequential CPU stand-in for the C ABI (tests/hostcheck/): the
    reference's templates over the policy classes can then be held against the reference's golden runs without a GPU"""
    if not os.path.isdir(REF):
        return STANDIN_OUT
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    src = os.path.join(HERE, "dropin.cpp")
    standin = os.path.join(ROOT, "tests", "hostcheck", "fsb_cpu_standin.cpp")
    deps = [src, standin, os.path.join(ROOT, "include", "fsb.h"), os.path.join(ROOT, "include", "fsb_flecsolve", "b200.hh")]
    if os.path.exists(STANDIN_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(STANDIN_OUT) for d in deps):
        return STANDIN_OUT
    os.makedirs(os.path.dirname(STANDIN_OUT), exist_ok=True)
    cmd = [cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wno-unused-local-typedefs",
           "-Wno-unused-variable", f'-DFSBD_REFERENCE_ROOT="{REF}"', *[a for d in INCLUDE_DIRS for a in ("-I", d)],
           src, standin, *[os.path.join(REF, u) for u in REF_UNITS], "-Wl,-z,defs", "-o", STANDIN_OUT]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("dropin stand-in build failed:\n" + r.stdout[-6000:])
    return STANDIN_OUT


def main() -> int:
    if not os.path.isdir(REF):
        print("[dropin] /root/reference not present: keeping the prebuilt library (if any)")
        return 0
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    src = os.path.join(HERE, "dropin.cpp")
    pkg = os.path.join(ROOT, "flecsolve_b200")
    deps = [src, os.path.join(ROOT, "include", "fsb.h"), os.path.join(ROOT, "include", "fsb_flecsolve", "b200.hh"),
            os.path.join(pkg, "libfsb.so")]
    deps += [os.path.join(b, f) for b, _, fs in os.walk(STUBS) for f in fs]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps if os.path.exists(d)):
        return 0
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-Wno-unused-local-typedefs", "-Wno-unused-variable",
           f'-DFSBD_REFERENCE_ROOT="{REF}"', *[a for d in INCLUDE_DIRS for a in ("-I", d)], src,
           *[os.path.join(REF, u) for u in REF_UNITS], "-o", OUT, "-L", pkg, "-lfsb", "-Wl,-rpath,$ORIGIN/../../../flecsolve_b200"]
    print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-6000:])
        return 1
    if r.stdout.strip():
        print(r.stdout[-3000:])
    return 0


if __name__ == "__main__":
    sys.exit(main())
