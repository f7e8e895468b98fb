"""The example applications (examples/): flecsolve-shaped user code over the header layer.
Building them needs no GPU and is part of the CPU suite; running them is part of `-m gpu`."""
import os
import re
import subprocess

import pytest

from flecsolve_b200 import _lib as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "examples", "_build")


def _build():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), f"CXX={cxx}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def test_examples_build():
    _build()
    assert all(os.path.exists(os.path.join(OUT, name)) for name in ("poisson", "implicit", "diffusion"))


def test_equilibrium_diffusion_on_the_sequential_stand_in(tmp_path):
    """examples/equilibrium_diffusion linked against the CPU stand-in for the C ABI (tests/hostcheck): the program's own
    logic -- two finite-volume blocks on a vec::multi, boundary conditions through the face coefficients, CG from
    diffusion.cfg -- drains v1 and keeps v2's constant, with no GPU involved"""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = str(tmp_path / "diffusion_cpu")
    r = subprocess.run([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-Wno-unused-local-typedefs", "-Wno-unused-variable",
                        "-I", os.path.join(ROOT, "flecsolve_b200", "include"), "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "equilibrium_diffusion", "diffusion.cc"),
                        os.path.join(ROOT, "tests", "hostcheck", "fsb_cpu_standin.cpp"), "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    r = subprocess.run([exe, "24", os.path.join(ROOT, "examples", "equilibrium_diffusion", "diffusion.cfg")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"converged after (\d+) iterations.*v1 in \[([0-9.e+-]+), ([0-9.e+-]+)\], v2 in \[([0-9.]+), ([0-9.]+)\]", r.stdout)
    assert m, r.stdout
    assert 10 < int(m.group(1)) <= 500 and abs(float(m.group(2))) < 1e-3 and abs(float(m.group(3))) < 1e-3, r.stdout
    assert abs(float(m.group(4)) - 2.0) < 1e-9 and abs(float(m.group(5)) - 2.0) < 1e-9, r.stdout


@pytest.mark.parametrize("src,args,pattern", [
    ("poisson/poisson.cc", ["32", "poisson/poisson.cfg"], r"converged after (\d+) iterations.*max error.*= ([0-9.e+-]+)"),
    ("heat_equation/implicit.cc", ["10", "heat_equation/implicit.cfg"],
     r"(\d+) steps \((\d+) attempts, (\d+) rejected\), max u = ([0-9.]+), heat ([0-9.]+) -> ([0-9.]+)"),
])
def test_other_examples_on_the_sequential_stand_in(tmp_path, src, args, pattern):
    """examples/poisson and examples/heat_equation against the CPU stand-in: their own logic, no GPU"""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = str(tmp_path / "example_cpu")
    r = subprocess.run([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-Wno-unused-local-typedefs", "-Wno-unused-variable",
                        "-I", os.path.join(ROOT, "flecsolve_b200", "include"), "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", src), os.path.join(ROOT, "tests", "hostcheck", "fsb_cpu_standin.cpp"),
                        "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    r = subprocess.run([exe, args[0], os.path.join(ROOT, "examples", args[1])], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(pattern, r.stdout)
    assert m, r.stdout
    if "poisson" in src:
        assert 20 < int(m.group(1)) < 400 and float(m.group(2)) < 5e-3, r.stdout
    else:
        assert int(m.group(1)) >= 5 and 0.0 < float(m.group(4)) <= 50.0 and float(m.group(6)) <= float(m.group(5)) * (1 + 1e-9), r.stdout


@pytest.mark.gpu
def test_examples_run():
    if F.device_count() == 0:
        pytest.skip("no CUDA device")
    _build()
    r = subprocess.run([os.path.join(OUT, "poisson"), "64", os.path.join(ROOT, "examples", "poisson", "poisson.cfg")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"converged after (\d+) iterations.*max error.*= ([0-9.e+-]+)", r.stdout)
    assert m and 100 < int(m.group(1)) < 400 and float(m.group(2)) < 1.2e-3, r.stdout
    r = subprocess.run([os.path.join(OUT, "implicit"), "24", os.path.join(ROOT, "examples", "heat_equation", "implicit.cfg")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"(\d+) steps \((\d+) attempts, (\d+) rejected\), max u = ([0-9.]+), heat ([0-9.]+) -> ([0-9.]+)", r.stdout)
    assert m, r.stdout
    assert int(m.group(1)) >= 10 and 0.0 < float(m.group(4)) <= 50.0 and float(m.group(6)) <= float(m.group(5)) * (1 + 1e-9)
    # examples/equilibrium_diffusion: v1 (Dirichlet sides) drains towards 0, v2 (zero-flux sides) keeps its constant
    r = subprocess.run([os.path.join(OUT, "diffusion"), "48", os.path.join(ROOT, "examples", "equilibrium_diffusion", "diffusion.cfg")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"converged after (\d+) iterations.*v1 in \[([0-9.e+-]+), ([0-9.e+-]+)\], v2 in \[([0-9.]+), ([0-9.]+)\]", r.stdout)
    assert m, r.stdout
    assert 10 < int(m.group(1)) <= 500 and abs(float(m.group(2))) < 1e-3 and abs(float(m.group(3))) < 1e-3, r.stdout
    assert abs(float(m.group(4)) - 2.0) < 1e-9 and abs(float(m.group(5)) - 2.0) < 1e-9, r.stdout

