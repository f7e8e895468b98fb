#!/usr/bin/env python
"""Headline benchmark: fp64 Jacobi-preconditioned CG on the synthetic 7-point 3-D Poisson system
256^3 (BASELINE.json configs[1]), row-sharded over N GPUs (z-slabs, strong scaling).

  python bench.py --gpus N --steps K --warmup W            this repo (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N ...            the reference's CPU path (oracle restatement)

A "step" is one CG iteration.  Rank 0 prints ONE JSON line; see README/DESIGN.md for the fields.
The headline (`value`, `e2e`, `roofline`) is always BASELINE.json configs[1] (7-point 256^3); `workloads` carries the
same measurement for the other workloads (27-point 512^3 = configs[4], the strong-scaling target) at the same N.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep NCCL's banner (NCCL_DEBUG=VERSION/INFO) off it
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE"):
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

# measured beside the headline at the same N (one short window each): configs[4], the strong-scaling target
EXTRA_WORKLOADS = {"poisson7_256": ["poisson27_512"]}

WORKLOADS = {
    # name: (stencil kind, nx, ny, nz, preconditioner)
    "poisson7_256": (7, 256, 256, 256, "dinv"),
    "poisson27_512": (27, 512, 512, 512, "dinv"),
    "poisson27_256": (27, 256, 256, 256, "dinv"),
    "poisson7_128": (7, 128, 128, 128, "dinv"),
    "poisson5_2d_256": (5, 256, 256, 1, None),
}


def describe(workload: str) -> str:
    kind, nx, ny, nz, precond = WORKLOADS[workload]
    return (f"{workload}: {kind}-point Poisson {nx}x{ny}x{nz}, {'Jacobi (1/diag) ' if precond else 'un'}preconditioned CG, "
            f"fp64, b = A x_true, x0 = 0")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def x_true(g0: int, n: int) -> np.ndarray:
    """x_true(g) = 1 + (h mod 1000)/1000, h = (uint32)(g * 2654435761)."""
    g = np.arange(g0, g0 + n, dtype=np.uint64)
    h = (g * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    return 1.0 + (h % np.uint64(1000)).astype(np.float64) / 1000.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            # "under load": the upper half of the samples (the timed region keeps the GPU busy)
            s = sorted(sm)
            out.update(sm_mhz=float(np.median(s[len(s) // 2:])), sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def pinned(n: int) -> np.ndarray:
    import torch
    return torch.empty(n, dtype=torch.float64).pin_memory().numpy()


def measure(world, ctx, workload: str, W: int, K: int, repeats: int, solver: str, with_e2e: bool, history_cap: int = 0,
            dictionary: bool = True) -> dict:
    """one workload on the ranks of `world`: device-resident timing of K iterations, SpMV roofline, optionally e2e.
    dictionary=False: the matrix is built with FSB_OPT_SPMV_DICTIONARY off, i.e. as any matrix with more than 256 distinct
    values is held (fp64 value per nonzero)"""
    from flecsolve_b200 import _lib as F
    from flecsolve_b200 import dist as D
    from flecsolve_b200 import host as H

    kind, nx, ny, nz, precond = WORKLOADS[workload]
    t0 = time.time()
    ctx.set_option("spmv_dictionary", 1 if dictionary else 0)
    A = F.ParCSR.stencil(ctx, kind, nx, ny, nz)
    ctx.sync()
    ctx.set_option("spmv_dictionary", 1)
    n_local, n_global = A.local_rows, A.global_rows
    nnz_local = A.nnz(0) + A.nnz(1)
    nnz_global = int(D.sum_over_ranks(world, nnz_local))
    S = H.Session(ctx, A)
    # b = A x_true computed on the device; x0 = 0
    xt = A.vector(x_true(A.row_begin, n_local))
    A.spmv(xt, S.b)
    ctx.sync()
    xt.destroy()
    setup_s = time.time() - t0
    ctx.set_option("profile", 1)
    split = {}
    hist_out = {}

    # ---- device-resident timing: iterations W+1 .. W+K of one solve, CUDA events on the library's stream
    def timed_solve(which):
        S.x.zero()
        ctx.sync()
        ctx.profile_read()
        D.barrier(world)
        # cg_device looks at iteration j (and fires the event marks) after issuing iteration j + lag: issue
        # lag more iterations so that the marks still bracket exactly K of them
        lag = 2 if which in ("cg_device", "cg_sr") else 0
        _, info, hist = S.solve(solver=which, precond=precond, maxiter=W + K + lag, rtol=0.0, ev_start=W, ev_stop=W + K,
                                lag=lag, history_cap=history_cap)
        ctx.sync()
        D.barrier(world)
        ms = ctx.event_elapsed_ms(0, 1)
        (ms_d, ms_o), (n_d, n_o) = ctx.profile_read_split()
        split.update(diag_avg_ms=ms_d / max(n_d, 1), offd_avg_ms=ms_o / max(n_o, 1), launches_per_spmv=1 if n_o == 0 else 2)
        hist_out[which] = np.array(hist)
        return ms, info, ms_d + ms_o, n_d

    timed_solve(solver)  # cold pass: allocations, occupancy queries, NCCL channels
    # the timed region is short (K x ~0.5 ms): run it `repeats` times and report the MEDIAN run
    runs = sorted((timed_solve(solver) for _ in range(max(1, repeats))), key=lambda r: r[0])
    ms, info, spmv_ms, spmv_n = runs[len(runs) // 2]
    # the other CG flavour beside it (cg: flecsolve's template, 3 host reads per iteration;
    # cg_device: scalars on the device, residual norm inspected late -- SURVEY 8(f) N1)
    other = "cg_device" if solver == "cg" else "cg"
    timed_solve(other)
    oruns = sorted((timed_solve(other) for _ in range(max(1, repeats))), key=lambda r: r[0])
    oms, oinfo = oruns[len(oruns) // 2][:2]
    oms = D.max_over_ranks(world, oms)
    # single-reduction CG (solvers/cg_sr.hh): two launches and one synchronisation point per iteration
    timed_solve("cg_sr")
    sruns = sorted((timed_solve("cg_sr") for _ in range(max(1, repeats))), key=lambda r: r[0])
    sms, sinfo = sruns[len(sruns) // 2][:2]
    sms = D.max_over_ranks(world, sms)
    ms = D.max_over_ranks(world, ms)
    ms_per_step = ms / K
    ctx.set_option("profile", 0)

    peak, peak_src = measured_peak_gbs()
    N, nnz = n_local, nnz_local
    wide = nnz >= 2 ** 31 - 16
    off = 8 if wide else 4
    # ALGORITHMIC bytes: fp64 values + int32 columns + row offsets + x + y (+ dot operand = x for CG: free)
    spmv_bytes = 12 * nnz + off * (N + 1) + 16 * N + 8 * A.num_ghosts
    iter_bytes = 12 * nnz + off * N + (108 - 4) * N + 8 * A.num_ghosts  # 12 nnz + 108 N (int32 offsets)
    # bytes of the format the device really streams (window format: 8 + 2 per nonzero, 16-bit block-relative row offsets)
    window = A.info("window_format") == 1
    dictionary = A.info("value_dictionary") if window else 0
    if dictionary:  # 1-byte value index + 16-bit position per slot (rows padded to 8 slots), 3 x 16-bit row meta
        fmt_bytes = 3 * A.info("dictionary_slots") + 2 * (3 * N + A.info("row_blocks")) + 16 * N + 16 * A.num_ghosts
    else:
        fmt_bytes = (10 * nnz + 2 * (2 * N + A.info("row_blocks")) if window else 12 * nnz + off * (N + 1)) + 16 * N + 16 * A.num_ghosts
    spmv_avg_ms = spmv_ms / max(spmv_n, 1)  # per SpMV: all its launches (one, unless the two-launch transports are in use)
    spmv_gbs = spmv_bytes / spmv_avg_ms / 1e6 if spmv_avg_ms > 0 else 0.0
    iter_gbs = iter_bytes / ms_per_step / 1e6
    out = {
        "workload": describe(workload), "value": 1000.0 / ms_per_step, "ms_per_step": ms_per_step,
        "rows": n_global, "nnz": nnz_global, "partition": f"{world.size} z-slab(s) of {n_local} rows",
        "gpu_launches": int(info.window_launches), "setup_seconds": setup_s,
        "device_format": (f"window + value dictionary: the matrix holds {dictionary} distinct values, so the device streams a 1-byte "
                          "value index + 16-bit position in the row block's staged x segments per nonzero (3 B, rows padded to 8 "
                          "slots) and multiplies the same fp64 values in the same order (general matrices: 10 B per nonzero, "
                          "see `general_values`)" if dictionary else
                          "window: fp64 value + 16-bit position in the row block's staged x segments per nonzero (10 B), "
                          "16-bit block-relative row offsets" if window else
                          f"csr: fp64 value + int32 column per nonzero (12 B), int{off * 8} row offsets"),
        "value_dictionary": int(dictionary),
        "spmv_launches_per_product": split.get("launches_per_spmv"),
        "fused_halo": bool(A.info("fused_halo")),
        "roofline": {
            "bound": "hbm", "kernel": ("spmv_window_kernel" if window else "spmv_stream_kernel") + " (y = A p fused with p.Ap"
                                      + (", ghost exchange and all-reduce inside" if A.info("fused_halo") else "") + ")",
            "achieved": spmv_gbs, "peak": peak, "unit": "GB/s", "frac": spmv_gbs / peak,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": spmv_bytes, "avg_launch_ms": spmv_avg_ms,
            "launches_timed": spmv_n, "share_of_step": spmv_avg_ms / ms_per_step if ms_per_step > 0 else None,
            "format_bytes_per_launch": fmt_bytes, "achieved_format_gbs": fmt_bytes / spmv_avg_ms / 1e6 if spmv_avg_ms > 0 else 0.0,
            "frac_format": fmt_bytes / spmv_avg_ms / 1e6 / peak if spmv_avg_ms > 0 else 0.0,
            "note": "achieved = SURVEY 8(d) algorithmic bytes (12 B/nnz CSR) / time; the device streams format_bytes "
                    "(value dictionary: ~3.4 B/nnz, window format: 10 B/nnz), so achieved may exceed the copy peak while "
                    "achieved_format cannot",
            "iteration": {"algorithmic_bytes": iter_bytes, "achieved": iter_gbs, "frac": iter_gbs / peak,
                          "frac_of_nominal_8TBs": iter_gbs / 8000.0},
        },
        "solver": solver,
        "other_solver": {"solver": other, "value": 1000.0 * K / oms, "ms_per_step": oms / K,
                         "gpu_launches": int(oinfo.window_launches),
                         "iteration_frac_of_peak": iter_bytes / (oms / K) / 1e6 / peak},
        "single_reduction_solver": {"solver": "cg_sr", "value": 1000.0 * K / sms, "ms_per_step": sms / K,
                                    "gpu_launches": int(sinfo.window_launches),
                                    "what": "Chronopoulos-Gear CG, scalars on the device: 2 launches / iteration, 12 nnz + 116 N bytes"},
    }
    if history_cap:
        out["_history"] = hist_out.get(solver)

    if with_e2e:
        # ---- end to end: full solve to rtol 1e-9 through the host-buffer call (H2D b, x0; D2H x), pinned buffers in place
        b_host, x_host = pinned(n_local), pinned(n_local)
        b_host[:] = S.b.download()
        best = None
        for rep in range(3):  # first rep warms the pinned buffers / page tables; the faster of the next two is reported
            x_host[:] = 0.0
            ctx.sync()
            D.barrier(world)
            t1 = time.perf_counter()
            x_out, einfo, _ = S.solve(b_host, x_host, inplace=True, solver="cg", precond=precond, maxiter=5000, rtol=1e-9)
            dt = D.max_over_ranks(world, time.perf_counter() - t1)
            if rep > 0 and (best is None or dt < best[0]):
                best = (dt, einfo.iters, einfo.h2d_ms, einfo.solve_ms, einfo.d2h_ms)
        e2e_s, e2e_iters, h2d_ms, solve_ms, d2h_ms = best
        err = D.max_over_ranks(world, float(np.abs(x_out - x_true(A.row_begin, n_local)).max()))
        out["e2e"] = {
            "value": e2e_iters / e2e_s if e2e_s > 0 and e2e_iters > 0 else 0.0, "unit": "iterations/s",
            "h2d_bytes_per_step": 16 * n_local / max(e2e_iters, 1), "d2h_bytes_per_step": 8 * n_local / max(e2e_iters, 1),
            "iterations": e2e_iters, "seconds": e2e_s,
            "what": "fsbh_solve with pinned host b, x0 in and x out, rtol 1e-9f, wall clock around the call",
            "max_abs_error_vs_x_true": err,
            "breakdown_ms": {"h2d": h2d_ms, "solver_call": solve_ms, "of_which_iterations_at_device_rate": e2e_iters * ms_per_step,
                             "d2h": d2h_ms, "binding_overhead": e2e_s * 1e3 - h2d_ms - solve_ms - d2h_ms,
                             "note": "solver_call - iterations = prologue (|b|, residual, first preconditioner apply), "
                                     "diagnostics and the final |x|; rank 0's numbers"},
        }
    S.close()
    A.destroy()
    return out


def run_ours(args):
    from flecsolve_b200 import dist as D

    world = D.init(D.world_from_env())
    if world.size != args.gpus:
        if world.size == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    ctx = D.make_context(world)
    W, K = args.warmup, args.steps
    sampler = ClockSampler(world.local_rank).start() if world.is_root else None
    cpu_iters = args.cpu_iters
    head = measure(world, ctx, args.workload, W, K, args.repeats, args.solver, with_e2e=True, history_cap=W + K)
    clocks = sampler.stop() if sampler else {}
    gpu_hist = head.pop("_history")

    def general_values(full, wl, w, k):
        """the same workload held as a matrix with more than 256 distinct values is (no value dictionary)"""
        if not full.get("value_dictionary"):
            return None
        g = measure(world, ctx, wl, w, k, 1, args.solver, with_e2e=False, dictionary=False)
        return {"what": "same solver and system, matrix built with FSB_OPT_SPMV_DICTIONARY off: the SpMV streams the fp64 "
                        "values (window format, 10 B per nonzero) as for any matrix with more than 256 distinct values",
                "value": g["value"], "ms_per_step": g["ms_per_step"], "device_format": g["device_format"],
                "spmv_avg_launch_ms": g["roofline"]["avg_launch_ms"], "spmv_frac_of_peak": g["roofline"]["frac"],
                "spmv_format_gbs": g["roofline"]["achieved_format_gbs"],
                "iteration_frac_of_peak": g["roofline"]["iteration"]["frac"],
                "other_solver": {"solver": g["other_solver"]["solver"], "value": g["other_solver"]["value"]}}

    head["general_values"] = general_values(head, args.workload, W, K)
    extra = {}
    if not args.no_extra_workloads:
        for wl in EXTRA_WORKLOADS.get(args.workload, []):
            r = measure(world, ctx, wl, 5, 30, 1, args.solver, with_e2e=False)
            r.pop("_history", None)
            r["general_values"] = general_values(r, wl, 5, 30)
            extra[wl] = r

    line = {
        "metric": "cg_iterations_per_second",
        "value": head["value"],
        "unit": "iterations/s",
        "n_gpus": world.size,
        "steps": K,
        "warmup": W,
        "ms_per_step": head["ms_per_step"],
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": head["workload"],
            "rows": head["rows"], "nnz": head["nnz"], "partition": head["partition"],
            "device_format": head["device_format"],
            "value_dictionary": head.get("value_dictionary", 0),
            "l2_policy": "inputs larger than L2 (per-iteration working set %.2f GB >> 126 MB)"
                         % (head["roofline"]["iteration"]["algorithmic_bytes"] / 1e9),
            "timed": f"iterations {W + 1}..{W + K} of one solve (rtol 0), CUDA events on the library stream, max over ranks; "
                     f"median of {max(1, args.repeats)} such solves",
            "host_templates": "this repo's flecsolve-shaped host layer (same call sequence as the reference's op::cg; the "
                              "reference's own cg.hh over the same kernels is timed in `reference_templates` when "
                              "tests/dropin/_build/libfsb_dropin.so is present)",
        },
        "roofline": head["roofline"],
        "e2e": head["e2e"],
        "gpu_launches": head["gpu_launches"],
        "solver": head["solver"],
        "other_solver": head["other_solver"],
        "single_reduction_solver": head["single_reduction_solver"],
        "general_values": head.get("general_values"),
        "spmv_launches_per_product": head["spmv_launches_per_product"],
        "fused_halo": head["fused_halo"],
        "clocks": clocks,
        "setup_seconds": head["setup_seconds"],
        "workloads": extra,
    }
    tp = os.path.join(ROOT, "profiles", "r2_spmv_ncu_summary.json")
    line["roofline"]["traffic"] = None
    if os.path.exists(tp) and args.workload == "poisson7_256" and world.size == 1:
        try:
            t = json.load(open(tp))
            # the capture names the instantiation it measured; only quote it for the format this run used
            if bool(t.get("value_dictionary")) == bool(head.get("value_dictionary")):
                line["roofline"]["traffic"] = t.get("dram_bytes_per_launch")
                line["roofline"]["traffic_source"] = ("not measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of "
                                                      "the same kernel instantiation on the same workload from the committed "
                                                      "ncu --set full capture profiles/r2_spmv_ncu_summary.json (" +
                                                      str(t.get("kernel")) + ")")
            if t.get("general_format"):
                line["roofline"]["traffic_general_format"] = t["general_format"]
        except Exception:
            pass
    if world.is_root and world.size == 1:
        ref = reference_templates(ctx, args.workload, W, K)
        if ref:
            line["reference_templates"] = ref
    if world.is_root and world.size == 1 and not args.no_cpu_baseline:
        base = cpu_baseline(args.workload, sample_iters=cpu_iters)
        ref_hist = np.array(base.pop("history"))
        m = min(len(ref_hist), len(gpu_hist))
        if m:
            rel = np.abs(gpu_hist[:m] - ref_hist[:m]) / ref_hist[:m]
            line["parity"] = {"what": "residual norms of the first iterations: this run's device solve vs the CPU restatement "
                                      "of the reference on the same full-size system",
                              "iters_compared": int(m), "max_rel_diff": float(rel.max()), "tolerance": 1e-8,
                              "ok": bool(rel.max() <= 1e-8)}
        serial = reference_serial(args.workload)
        if serial:  # the reference's own serial code beside the port's one-thread figure: is the port representative?
            base["reference_serial"] = serial
        line["cpu_baseline"] = base
    ctx.close()
    if world.is_root:
        print(json.dumps(line), flush=True)
    D.finalize(world)


def reference_templates(ctx, workload: str, W: int, K: int):
    """the same window timed through the REFERENCE's own cg.hh (tests/dropin: unmodified /root/reference headers over the
    policy classes of include/fsb_flecsolve/b200.hh, prebuilt where /root/reference exists); host buffers in and out"""
    try:
        from tests import dropin as DI
        if not DI.available():
            return None
        from flecsolve_b200 import _lib as F
        kind, nx, ny, nz, precond = WORKLOADS[workload]
        A = F.ParCSR.stencil(ctx, kind, nx, ny, nz)
        n = A.local_rows
        xt, bv = A.vector(x_true(A.row_begin, n)), A.vector()
        A.spmv(xt, bv)
        b = bv.download()
        xt.destroy(); bv.destroy()
        d = DI.Dropin(DI.lib())
        best = None
        for _ in range(3):
            _, info, _ = d.solve(ctx.h, A.h, b, np.zeros(n), solver="cg", precond=precond, rtol=0.0, maxiter=W + K,
                                 ev_start=W, ev_stop=W + K)
            ms = ctx.event_elapsed_ms(0, 1) / K
            best = ms if best is None else min(best, ms)
        A.destroy()
        return {"value": 1000.0 / best, "ms_per_step": best, "gpu_launches": int(info.window_launches),
                "what": "flecsolve's own solvers/cg.hh + vectors/core.hh + operators/*.hh (unmodified) over "
                        "include/fsb_flecsolve/b200.hh, same kernels, same window of iterations"}
    except Exception as e:  # never let the extra leg break the benchmark line
        return {"unavailable": str(e)[:200]}


def host_threads() -> int:
    """all host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its workers)"""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(workload: str, sample_iters: int, warm: int = 2) -> dict:
    """The reference's CPU path (oracle restatement: size_t indices, diag/offd split + add pass,
    one pass per vector op, 3 blocking reductions per iteration) on the host cores of this box."""
    import oracle as O
    kind, nx, ny, nz, precond = WORKLOADS[workload]
    threads = host_threads()
    O.set_threads(threads)
    rp, col, val = O.stencil_csr(kind, nx, ny, nz)
    M = O.ParCSR(rp, col, val, colours=threads)  # one colour per thread, like one MPI rank per core
    n = len(rp) - 1
    b = M.spmv(x_true(0, n))
    dinv = M.dinv() if precond else None
    M.cg(b, dinv=dinv, maxiter=warm, rtol=0.0)  # page in
    t0 = time.perf_counter()
    _, info, hist = M.cg(b, dinv=dinv, maxiter=sample_iters, rtol=0.0, history_cap=sample_iters)
    dt = time.perf_counter() - t0
    out = {"value": sample_iters / dt, "unit": "iterations/s", "cores": threads, "kind": "port",
           "sample": f"first {sample_iters} CG iterations of the same system (x0 = 0), {threads} OpenMP threads = colours, "
                     f"{dt:.1f} s", "seconds": dt, "host_cores": os.cpu_count(), "omp_max_threads_in_use": O.max_threads(),
           "history": [float(h) for h in hist]}
    if threads > 1:  # the serial figure beside it (one colour, one thread), on a shorter sample
        one_iters = max(3, sample_iters // 8)
        M1 = O.ParCSR(rp, col, val, colours=1)
        O.set_threads(1)
        try:
            M1.cg(b, dinv=dinv, maxiter=1, rtol=0.0)
            t0 = time.perf_counter()
            M1.cg(b, dinv=dinv, maxiter=one_iters, rtol=0.0)
            d1 = time.perf_counter() - t0
        finally:
            O.set_threads(threads)
        out["one_thread"] = {"value": one_iters / d1, "iterations": one_iters, "seconds": d1}
    return out


def reference_serial(workload: str, iters: int = 8) -> dict | None:
    """The reference's OWN serial code on one core: oracle/_ref/refcheck = flecsolve/solvers/cg.hh + matrices/seq.hh +
    vectors/seq.hh compiled from the reference tree against stub FleCSI headers (oracle/refcheck/build.py), run on the same
    system.  The binary reports no timings, so two runs (2 and 2 + iters iterations) are timed and subtracted."""
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "refcheck")
    if not os.path.exists(exe):
        return None
    kind, nx, ny, nz, precond = WORKLOADS[workload]
    if nx * ny * nz > 2 ** 25:  # the serial assembly of the larger systems takes minutes
        return None
    try:
        with tempfile.TemporaryDirectory() as tmp:
            out = os.path.join(tmp, "x.bin")
            times = []
            for maxiter in (2, 2 + iters):
                t0 = time.perf_counter()
                r = subprocess.run([exe, str(kind), str(nx), str(ny), str(nz), "cg", "1" if precond else "0", "0", str(maxiter),
                                    "1", "0", "0", "1", "2", out], capture_output=True, text=True, timeout=600)
                times.append(time.perf_counter() - t0)
                if r.returncode != 0:
                    return None
        dt = times[1] - times[0]
        if dt <= 0:
            return None
        return {"value": iters / dt, "unit": "iterations/s", "cores": 1, "kind": "reference",
                "sample": f"the reference's own cg.hh + seq.hh (oracle/_ref/refcheck), {iters} iterations of the same system: "
                          f"{times[1]:.1f} s for {2 + iters} iterations minus {times[0]:.1f} s for 2 (set-up included in both)"}
    except Exception:
        return None


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path is not buildable here
    (needs FleCSI/MPI/Boost), so this times its line-faithful restatement under oracle/ on ALL host cores."""
    from flecsolve_b200 import dist as D
    world = D.world_from_env()
    if world.rank != 0:
        return
    # at least 20 and at most --cpu-iters-cap iterations: a handful of iterations mostly times page faults
    iters = min(max(args.steps, 20), args.cpu_iters_cap)
    base = cpu_baseline(args.workload, sample_iters=iters, warm=max(args.warmup, 1) if args.warmup <= 5 else 3)
    base.pop("history", None)
    if base["cores"] * 2 < (os.cpu_count() or 1):
        print(f"[bench] warning: reference arm runs on {base['cores']} of {os.cpu_count()} host cores", file=sys.stderr)
    line = {
        "impl": "reference",
        "metric": "cg_iterations_per_second", "value": base["value"], "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / base["value"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": describe(args.workload),
                   "implementation": "reference CPU path restated (oracle/oracle.cpp: size_t indices, diag/offd split + add "
                                     "pass, one pass per vector op, blocking reductions), one colour per OpenMP thread",
                   "sampled_steps": iters},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    serial = reference_serial(args.workload)
    if serial:
        line["reference_serial"] = serial
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="poisson7_256", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-iters", type=int, default=40, help="CG iterations of the CPU baseline sample")
    ap.add_argument("--cpu-iters-cap", type=int, default=60)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-workloads", action="store_true", help="skip the 27-point 512^3 block")
    ap.add_argument("--repeats", type=int, default=3, help="timed solves of W+K iterations; the median one is reported")
    ap.add_argument("--solver", default="cg", choices=["cg", "cg_device"],
                    help="cg: op::cg, the same call sequence as the reference's template (headline); cg_device: scalars kept on the device")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
