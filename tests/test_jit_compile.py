"""Run-time instantiation of program kernels (flecsolve_b200/csrc/jit.cu): NVRTC compiles the same
ew_program_body template the ahead-of-time registry uses, for any canonical statement group.  Compiling needs
no GPU, so it is checked here; loading and launching the result is exercised on hardware
(tests/test_zz_jit_hw.py, opt-in: FSB_JIT stays off by default until that has run on a B200)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from flecsolve_b200 import _lib as F

SET, SCALE, LIN2, MUL, DIV, RECIP, ABS, ADDS = range(8)
DOT, ASUM, AMAX, MIN, MAX, POWSUM = range(16, 22)


def compile_group(stmts, dev=False, box=False):
    raw = np.array(stmts, dtype=np.int32).ravel()
    nbytes, log = C.c_int64(), C.create_string_buffer(1 << 16)
    rc = F.lib().fsb_debug_jit_compile(raw.ctypes.data_as(C.POINTER(C.c_int32)), len(stmts), int(dev), int(box),
                                       C.byref(nbytes), log, len(log))
    return rc, nbytes.value, log.value.decode()


GROUPS = {
    "cg_update": [(LIN2, 1, 0, 1), (LIN2, 3, 2, 3), (DOT, -1, 3, 3)],
    # BiCGStab on a 2-component vec::multi, first iteration: no ahead-of-time instantiation
    "bicg2_first": [(SCALE, 34, 33, -1), (SCALE, 26, 25, -1), (SET, 35, -1, -1), (SET, 27, -1, -1), (SET, 36, -1, -1),
                    (SET, 28, -1, -1), (DOT, -1, 26, 25), (DOT, -1, 34, 33)],
    # a BDF source term: six aliased axpys onto one vector (time-integrators/bdf.hh:319-439)
    "bdf_chain": [(LIN2, 0, 1, 2)] + [(LIN2, 0, k, 0) for k in range(3, 8)],
    # every statement and reduction kind at once
    "all_kinds": [(SET, 0, -1, -1), (SCALE, 1, 0, -1), (MUL, 2, 0, 1), (DIV, 3, 2, 1), (RECIP, 4, 3, -1), (ABS, 5, 4, -1),
                  (ADDS, 6, 5, -1), (ASUM, -1, 6, -1), (AMAX, -1, 6, -1), (MIN, -1, 5, -1), (POWSUM, -1, 4, -1)],
    # the slot limits: 12 statements over 12 vectors
    "limits": [(LIN2, k + 1, k, (k + 1) % 12) for k in range(11)] + [(MAX, -1, 11, -1)],
}


@pytest.mark.parametrize("box", [False, True], ids=["flat", "box"])
@pytest.mark.parametrize("dev", [False, True], ids=["imm", "dev"])
@pytest.mark.parametrize("name", sorted(GROUPS))
def test_any_canonical_group_compiles_for_sm_100a(name, dev, box):
    rc, nbytes, log = compile_group(GROUPS[name], dev, box)
    assert rc == 0 and nbytes > 10_000, log[:2000]
    assert "error" not in log.lower()


def test_compiled_kernel_matches_the_registered_instantiation(tmp_path, monkeypatch):
    """same template, same compiler back end: the run-time kernel of CG's update has the resources of the
    ahead-of-time one (44 registers, no local memory, 128-bit loads and stores)"""
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = tmp_path / "k.cubin"
    monkeypatch.setenv("FSB_JIT_DUMP", str(out))
    rc, _, log = compile_group(GROUPS["cg_update"])
    assert rc == 0, log
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", str(out)], capture_output=True, text=True).stdout
    assert "fsb_jit_entry" in res and "STACK:0" in res
    regs = int(res.split("REG:")[1].split()[0])
    assert regs <= 48  # the ahead-of-time instantiation of the same group uses 44
    sass = subprocess.run(["cuobjdump", "-sass", str(out)], capture_output=True, text=True).stdout
    assert sass.count("LDG.E.128") >= 4 and sass.count("STG.E.128") >= 2
    assert "DFMA" in sass and "DMUL" in sass and "DADD" in sass  # fused only where the reference's dot is


def test_statement_lists_beyond_the_limits_are_rejected():
    too_many_vectors = [(LIN2, k + 1, k, k + 20) for k in range(8)]
    rc, _, _ = compile_group(too_many_vectors)
    assert rc != 0 and "limits" in F.lib().fsb_last_error().decode()
