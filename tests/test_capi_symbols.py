"""The C-ABI library loads and exports every symbol include/fsb.h declares (no compute calls)."""
import ctypes
import os
import subprocess

import pytest

from flecsolve_b200 import _lib as F
from flecsolve_b200 import host as H


def test_header_symbols_exported():
    names = F.declared_symbols()
    assert len(names) > 50
    out = subprocess.run(["nm", "-D", "--defined-only", F.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in fsb.h but not exported: {missing}"


def test_ctypes_bindings_cover_header():
    lib = F.lib()
    for n in F.declared_symbols():
        fn = getattr(lib, n)
        assert fn.argtypes is not None or n in ("fsb_last_error", "fsb_version", "fsb_device_count"), n


def test_version_and_error_string():
    assert F.lib().fsb_version() >= 100
    assert isinstance(F.lib().fsb_last_error(), bytes)


def test_host_library_loads():
    L = H.lib()
    for n in ("fsbh_session_create", "fsbh_solve", "fsbh_solve_multi2", "fsbh_solve_subset", "fsbh_vector_selftest",
              "fsbh_adapter_apply"):
        assert hasattr(L, n)


def test_no_cpu_fallback():
    """Without a GPU the product refuses to run instead of silently computing on the host."""
    if F.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(F.FsbError) as e:
        F.Context(0)
    assert e.value.code == 4  # FSB_ERR_NOGPU
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under flecsolve_b200/ may reference it."""
    root = os.path.dirname(F.LIB_PATH)
    bad = []
    for base, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hh", ".cpp")) and f != "build.py":
                txt = open(os.path.join(base, f), errors="ignore").read()
                if "import oracle" in txt or "from oracle" in txt or "liboracle" in txt:
                    bad.append(os.path.join(base, f))
    assert not bad, bad
