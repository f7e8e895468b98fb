import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flecsolve_b200 import _lib as F, host as H
GOLD = json.load(open('tests/golden/reference_bdf.json'))
ctx = F.Context(0)
for entry in GOLD["rate"]:
    method, rtol, atol, dt0, dtmax, dtmin, tf, lam, ic, n, use_pi, controller, predictor = entry["case"]
    if method != "BDF6": continue
    ref = entry["result"]
    rp = np.arange(n + 1, dtype=np.int64)
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, np.arange(n, dtype=np.int64), np.ones(n))
    opts = H.make_bdf_options(method=method, time_rtol=rtol, time_atol=atol, initial_dt=dt0, max_dt=dtmax, min_dt=dtmin,
                              final_time=tf, use_pi_controller=bool(use_pi), controller=controller, predictor=predictor,
                              error_scaling="fixed-resolution", norm="inf", max_steps=1000)
    res, dts, good, vals = H.bdf_rate(ctx, A, opts, lam, ic)
    print(res.steps, res.rejects, res.attempts, "ref", ref["nsteps"], ref["rejects"], len(ref["steps"]))
    for k, s in enumerate(ref["steps"][:len(dts)]):
        rdt, rg, rv = float.fromhex(s[0]), s[1], float.fromhex(s[2])
        flag = "" if (rg == good[k] and abs(rdt - dts[k]) <= 1e-12 * rdt) else "  <-- MISMATCH"
        if k < 14 or flag:
            print(k, rdt, dts[k], rg, good[k], rv, vals[k], flag)
        if flag: break
