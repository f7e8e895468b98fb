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


def test_documents_name_real_entry_points():
    """every fsb_* / fsbh_* function INTEGRATION.md, DESIGN.md and README.md mention exists"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    declared = set(F.declared_symbols()) | {"fsb_ctx_s", "fsb_vec_s", "fsb_parcsr_s", "fsb_status", "fsb_option", "fsb_stat",
                                            "fsb_coef", "fsb_red_opts", "fsb_scalar_t", "fsb_token_t", "fsb_ctx_t", "fsb_vec_t",
                                            "fsb_parcsr_t"}
    driver = open(os.path.join(root, "flecsolve_b200", "host", "driver.cpp")).read()
    for doc in ("INTEGRATION.md", "DESIGN.md", "README.md"):
        text = open(os.path.join(root, doc)).read()
        for name in set(re.findall(r"\bfsb_[a-z0-9_]+\b", text)):
            if name.endswith("_") or name in ("fsb_vec_", "fsb_scalar_"):
                continue  # prefixes such as `fsb_vec_*`
            if name in ("fsb_flecsolve", "fsb_dropin", "fsb_host"):
                continue  # directory / library names (include/fsb_flecsolve/, libfsb_dropin.so, libfsb_host.so)
            assert name in declared, f"{doc} mentions {name}, which include/fsb.h does not declare"
        for name in set(re.findall(r"\bfsbh_[a-z0-9_]+\b", text)):
            assert f"int {name}(" in driver or f"{name}(" in driver, f"{doc} mentions {name}, which the driver does not define"


def test_test_infrastructure_stays_out_of_the_product():
    """the CPU stand-in for the C ABI (tests/hostcheck) is for the test suite only"""
    root = os.path.dirname(F.LIB_PATH)
    for base, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hh", ".cpp", ".inc")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "hostcheck" not in txt and "fsb_cpu_standin" not in txt, os.path.join(base, f)
