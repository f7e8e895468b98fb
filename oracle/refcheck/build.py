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
    src = os.path.join(HERE, "refcheck.cpp")
    deps = [src] + [os.path.join(b, f) for b, _, fs in os.walk(os.path.join(HERE, "stubs")) for f in fs]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return 0
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    # -ffp-contract=off and no -march: the reference's default x86-64 build has no FMA contraction
    cmd = [cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(HERE, "stubs"), "-I", REF, src, "-o", OUT]
    print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-4000:])
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
