"""Device-side timeline of the CG iteration (FSB_OPT_TIMELINE): every kernel stamps %globaltimer at its first CTA's
start and last CTA's end (+ marks for the ghost wait and the cross-rank all-reduce), so the time between kernels --
launch gaps, host round trips -- and the time inside them waiting for peers is measured on every rank, where
ncu / nsys cannot be used.  Run on 1 GPU or under torchrun:
    python scripts/gpu/timeline.py [workload] [solver] [iters]
Prints, per rank, the average over the timed iterations of each kernel of the iteration pattern."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import WORKLOADS, x_true  # noqa: E402
from flecsolve_b200 import _lib as F  # noqa: E402
from flecsolve_b200 import dist as D  # noqa: E402
from flecsolve_b200 import host as H  # noqa: E402

KIND = {1: "spmv", 2: "spmv(offd)", 3: "spmv+halo", 32: "halo push", 33: "halo unpack"}


def name(k):
    k = int(k)
    if k in KIND:
        return KIND[k]
    if k >= 16 and k < 32:
        return f"ew[{k - 16} stmts]"
    return f"kind{k}"


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "poisson7_256"
    solver = sys.argv[2] if len(sys.argv) > 2 else "cg"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    world = D.init(D.world_from_env())
    kind, nx, ny, nz, precond = WORKLOADS[workload]
    ctx = D.make_context(world)
    A = F.ParCSR.stencil(ctx, kind, nx, ny, nz)
    S = H.Session(ctx, A)
    xt = A.vector(x_true(A.row_begin, A.local_rows))
    A.spmv(xt, S.b)
    ctx.sync()
    lag = 2 if solver in ("cg_device", "cg_sr") else 0
    warm = 10

    def run():
        S.x.zero()
        ctx.sync()
        D.barrier(world)
        S.solve(solver=solver, precond=precond, maxiter=warm + iters + lag, rtol=0.0, ev_start=warm, ev_stop=warm + iters, lag=lag)
        ctx.sync()
        return ctx.event_elapsed_ms(0, 1) / iters

    run()
    ms_plain = run()
    ctx.set_option("timeline", 16384)
    ms_tl = run()
    tl = ctx.timeline_read()
    ctx.set_option("timeline", 0)
    t0, t1, k = tl[:, 0].astype(np.int64), tl[:, 1].astype(np.int64), tl[:, 2].astype(np.int64)
    # find the iteration period: the pattern of kinds repeats; use the fused/plain spmv launches as anchors
    anchors = np.flatnonzero((k == 3) | (k == 1))
    anchors = anchors[len(anchors) // 3: len(anchors) // 3 + iters // 2 + 1]  # steady state, away from prologue/epilogue
    per = int(np.median(np.diff(anchors)))
    rows = {}
    for a in anchors[:-1]:
        if not np.array_equal(k[a:a + per], k[anchors[0]:anchors[0] + per]):
            continue
        for j in range(per):
            i = a + j
            d = rows.setdefault(j, {"kind": k[i], "dur": [], "gap": [], "wait": [], "ar": [], "push": [], "pd": []})
            d["dur"].append((t1[i] - t0[i]) / 1e3)
            d["gap"].append((t0[i] - t1[i - 1]) / 1e3)
            w4, w5, w6, w7, w3 = (int(tl[i, c]) for c in (4, 5, 6, 7, 3))
            if w5:
                d["wait"].append(w5 / 1e3)
            if int(tl[i, 8]):
                d["pd"].append([(int(tl[i, c]) - int(t0[i])) / 1e3 for c in (8, 9)])
            if w7 and w6 != 2 ** 64 - 1:
                d["ar"].append((w7 - w6) / 1e3)
            if w3:
                d["push"].append((w3 - int(t0[i])) / 1e3)
    period = np.median(np.diff(t0[anchors])) / 1e3
    out = [f"rank {world.rank}/{world.size} {workload} {solver}: {ms_plain * 1e3:.1f} us/iteration (events), "
           f"{ms_tl * 1e3:.1f} with the timeline on, {period:.1f} between SpMV starts; {per} launches per iteration"]
    out.append(f"  {'kernel':16s} {'gap before us':>14s} {'duration us':>12s} {'push done at':>13s} {'offd phase':>11s} {'all-reduce':>11s}")
    tot_gap = tot_dur = 0.0
    for j in range(per):
        d = rows[j]
        m = lambda v: f"{np.mean(v):.1f}" if v else "-"
        out.append(f"  {name(d['kind']):16s} {m(d['gap']):>14s} {m(d['dur']):>12s} {m(d['push']):>13s} {m(d['wait']):>11s} {m(d['ar']):>11s}"
                   + (f"   push: acks seen {np.mean([p[0] for p in d['pd']]):.1f}, stores issued {np.mean([p[1] for p in d['pd']]):.1f} "
                      f"us after start" if d["pd"] else ""))
        tot_gap += np.mean(d["gap"]); tot_dur += np.mean(d["dur"])
    out.append(f"  {'sum':16s} {tot_gap:14.1f} {tot_dur:12.1f}")
    for r in range(world.size):
        D.barrier(world)
        if r == world.rank and (r == 0 or r == world.size - 1 or r == world.size // 2):
            print("\n".join(out), flush=True)
    S.close(); A.destroy(); ctx.close()
    D.finalize(world)


if __name__ == "__main__":
    main()
