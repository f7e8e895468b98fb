"""Golden vectors of the reference's BDF integrator (tests/golden/reference_bdf.json, produced by
oracle/_ref/refcheck_bdf = flecsolve/time-integrators/bdf.hh + bdf.cc compiled from /root/reference).
CPU-side checks: the goldens satisfy the reference's own known-answer test
(time-integrators/test/implicit.cc:121-134) and the live binary, when present, still reproduces them."""
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_bdf.json")))


def _case(method):
    return next(e for e in GOLD["rate"] if e["case"][0] == method and e["case"][7] == -1.0 and e["case"][8] == 3.0)


def test_reference_kat_bdf2():
    r = _case("BDF2")["result"]
    assert float.fromhex(r["final_time"]) == 1.0
    assert float.fromhex(r["error"]) < 1e-3
    assert r["nsteps"] <= 49 and r["rejects"] <= 7


def test_reference_kat_bdf5():
    r = _case("BDF5")["result"]
    assert float.fromhex(r["final_time"]) == 1.0
    assert float.fromhex(r["error"]) < 1e-7
    assert r["nsteps"] == 48 and r["rejects"] == 14  # EXPECT_EQ in the reference


def test_golden_histories_are_consistent():
    for e in GOLD["rate"] + GOLD["heat"]:
        r = e["result"]
        good = [s[1] for s in r["steps"]]
        assert sum(good) == r["nsteps"] and len(good) - sum(good) == r["rejects"]
        t = sum(float.fromhex(s[0]) for s in r["steps"] if s[1])
        assert abs(t - float.fromhex(r["final_time"])) < 1e-12


def test_live_refcheck_bdf_reproduces_golden():
    exe = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "refcheck_bdf")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/refcheck_bdf not built (needs /root/reference)")
    e = _case("BDF5")
    c = e["case"]
    out = subprocess.run([exe, c[0], *[repr(float(v)) for v in c[1:9]], str(c[9]), str(c[10]), c[11], c[12]],
                         capture_output=True, text=True, check=True).stdout
    assert json.loads(out) == e["result"]


def test_reference_kat_rk():
    """time-integrators/test/explicit.cc:69-98 on the reference-generated goldens."""
    rk = json.load(open(os.path.join(HERE, "golden", "reference_rk.json")))["cases"]
    by = {(c["case"][0], c["case"][7]): c["result"] for c in rk if c["case"][8] == -1.0 and c["case"][9] == 3.0}
    assert float.fromhex(by[(23, 1)]["error"]) < 1e-5 and float.fromhex(by[(45, 1)]["error"]) < 1e-9
    assert by[(23, 0)]["nsteps"] == 32 and float.fromhex(by[(23, 0)]["error"]) < 1e-5
    assert by[(45, 0)]["nsteps"] == 6 and float.fromhex(by[(45, 0)]["error"]) < 1e-6
    for r in by.values():
        assert float.fromhex(r["final_time"]) == 1.0
