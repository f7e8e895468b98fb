"""Builds oracle/_ref/refcheck from the reference's own headers (container only: needs
/root/reference).  The binary is git-ignored but travels to the GPU box with the snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
OUT = os.path.join(ROOT, "oracle", "_ref", "refcheck")


def main() -> int:
    if not os.path.isdir(REF):
        print("[refcheck] /root/reference not present: keeping the prebuilt binary (if any)")
        return 0
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    stubs = [os.path.join(b, f) for b, _, fs in os.walk(os.path.join(HERE, "stubs")) for f in fs]
    # (binary, our driver, reference translation units it needs)
    targets = [
        ("refcheck", "refcheck.cpp", []),
        ("refcheck_bdf", "refcheck_bdf.cpp", ["flecsolve/time-integrators/bdf.cc",
                                              "flecsolve/time-integrators/bdf_parameters.cc",
                                              "flecsolve/vectors/util.cc"]),
        ("refcheck_rk", "refcheck_rk.cpp", []),
        ("refcheck_mtx", "refcheck_mtx.cpp", []),
    ]
    rc = 0
    for name, driver, ref_units in targets:
        out = os.path.join(os.path.dirname(OUT), name)
        src = os.path.join(HERE, driver)
        deps = [src] + stubs
        if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
            continue
        os.makedirs(os.path.dirname(out), exist_ok=True)
        # -ffp-contract=off and no -march: the reference's default x86-64 build has no FMA contraction
        cmd = [cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(HERE, "stubs"), "-I", REF, src,
               *[os.path.join(REF, u) for u in ref_units], "-o", out]
        print("+", " ".join(cmd), flush=True)
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-4000:])
            rc = 1
    return rc


if __name__ == "__main__":
    sys.exit(main())
