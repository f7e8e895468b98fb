"""Turns the raw captures a GPU run left in gpurun_out/ into the tracked summaries under profiles/.
Usage (in the build container, after scripts/gpu/final_n1.sh ran under gpurun): python scripts/summarize_profiles.py"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(src, dst, title):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    agg = defaultdict(list)
    for r in rows:
        agg[r[4].split("(")[0].replace("void ", "").replace("fsb::", "")].append(float(r[-1]))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(title + "\n")
        f.write(f"# {len(rows)} launches, {tot / 1e6:.3f} ms of kernel time in the window; times are per launch, cold cache, serialised\n")
        f.write(f"{'kernel':78s} {'n':>5s} {'avg us':>9s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:78s} {len(v):5d} {sum(v) / len(v) / 1000:9.1f} {sum(v) / tot:7.3f}\n")
    return agg


def full(reps, dst, title, key="spmv_stream"):
    out = [title]
    first = {}
    for rep in reps:
        pre = rep[:-len(".ncu-rep")] + ".raw.csv"  # extracted on the GPU box (scripts/gpu/r2_final.sh) to keep the download small
        if os.path.exists(pre):
            text = open(pre).read()
        else:
            text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(text.splitlines()))
        rows = [x for x in rows if len(x) > 20]
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for row in rows[2:]:
            name = row[idx["Kernel Name"]]
            out.append(f"\n== {name}   [{os.path.basename(rep)}]")
            for w in WANT:
                if w in idx:
                    out.append(f"  {w:84s} {row[idx[w]]} {units[idx[w]]}")
            if key in name and "spmv" not in first:
                def gb(key):
                    v, u = float(row[idx[key]]), units[idx[key]]
                    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
                first["spmv"] = {"kernel": name, "dram_bytes_read": gb("dram__bytes_read.sum"),
                                 "dram_bytes_write": gb("dram__bytes_write.sum"),
                                 "dram_bytes_per_launch": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"),
                                 "gpu_time_us_under_ncu": float(row[idx["gpu__time_duration.sum"]]),
                                 "source": os.path.basename(rep)}
    open(dst, "w").write("\n".join(out) + "\n")
    return first


def round2():
    """round 2: scripts/gpu/r2_final.sh left r2_launches.csv, r2_final_{spmv7,ew,spmv27_512}.ncu-rep, r2_bench_*.json"""
    import shutil
    launches(os.path.join(G, "r2_launches.csv"), os.path.join(P, "r2_launches_cg_poisson7_256.txt"),
             "# round 2: ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300, python bench.py --steps 20 --warmup 5 "
             "--no-cpu-baseline --no-extra-workloads --repeats 1\n# (7-pt 256^3 Jacobi-CG; the window covers the CG flavours "
             "bench.py times: op::cg, cg_device, cg_sr)")
    f = full([p for p in (os.path.join(G, "r2_final_spmv7.ncu-rep"), os.path.join(G, "r2_final_spmv27_512.ncu-rep"),
                          os.path.join(G, "r2_final_spmv27_512_general.ncu-rep"), os.path.join(G, "r2_final_ew.ncu-rep"))
              if os.path.exists(p) or os.path.exists(p[:-len(".ncu-rep")] + ".raw.csv")], os.path.join(P, "r2_ncu_full_summary.txt"),
             "# round 2: ncu --set full --clock-control none --import-source on; spmv_window_kernel inside bench.py (7-pt 256^3, "
             "fused <Ap,p>), the same kernel on 27-pt 512^3 (scripts/gpu/spmv_sweep.py 27 512 512 dotx), CG's element-wise kernels",
             key="spmv_window")
    if "spmv" in f:
        s = f["spmv"]
        # template arguments <NSTAGE, DOT, HALO, JAC, VD>: VD = value dictionary
        import re
        m = re.search(r"<([^>]*)>", s["kernel"])
        s["value_dictionary"] = bool(m) and m.group(1).replace(" ", "").split(",")[-1] in ("1", "true")
        s["algorithmic_bytes_per_launch"] = 12 * 117047296 + 4 * (16777216 + 1) + 16 * 16777216
        bench = os.path.join(G, "r2_bench_n1.json")
        fmt = None
        if os.path.exists(bench):
            lines = [l for l in open(bench).read().splitlines() if l.startswith("{")]
            if lines:
                fmt = json.loads(lines[-1])["roofline"].get("format_bytes_per_launch")
        s["format_bytes_per_launch"] = fmt
        s["traffic_over_algorithmic"] = s["dram_bytes_per_launch"] / s["algorithmic_bytes_per_launch"]
        if fmt:
            s["traffic_over_format"] = s["dram_bytes_per_launch"] / fmt
        # the general (10 B per nonzero) window kernel, captured with the dictionary off
        g = full([os.path.join(G, "r2_final_spmv7_general.ncu-rep")], os.path.join(P, "r2_ncu_full_summary_general.txt"),
                 "# round 2: the same capture with FSB_SPMV_DICT=0 (window format, fp64 values: what a general matrix runs)",
                 key="spmv_window") if (os.path.exists(os.path.join(G, "r2_final_spmv7_general.ncu-rep")) or
                                        os.path.exists(os.path.join(G, "r2_final_spmv7_general.raw.csv"))) else {}
        if "spmv" in g:
            gg = g["spmv"]
            s["general_format"] = {"kernel": gg["kernel"], "dram_bytes_per_launch": gg["dram_bytes_per_launch"],
                                   "format_bytes_per_launch": 1506092180,
                                   "traffic_over_format": gg["dram_bytes_per_launch"] / 1506092180,
                                   "traffic_over_algorithmic": gg["dram_bytes_per_launch"] / s["algorithmic_bytes_per_launch"],
                                   "gpu_time_us_under_ncu": gg["gpu_time_us_under_ncu"], "source": gg["source"]}
        json.dump(s, open(os.path.join(P, "r2_spmv_ncu_summary.json"), "w"), indent=1)
        print(s)
    for name in ("r2_bench_n1.json", "r2_bench_ref.json", "r2_bench_n2.json", "r2_bench_n4.json", "r2_bench_n8.json"):
        src = os.path.join(G, name)
        if os.path.exists(src):
            lines = [l for l in open(src).read().splitlines() if l.startswith("{")]
            if lines:
                json.dump(json.loads(lines[-1]), open(os.path.join(P, name), "w"), indent=1)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "r2":
        return round2()
    launches(os.path.join(G, "launches_r1.csv"), os.path.join(P, "r1_launches_cg_poisson7_256.txt"),
             "# round 1: ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 260, python bench.py --steps 20 --warmup 5\n"
             "# (7-pt 256^3 Jacobi-CG; the window covers both CG flavours bench.py times: op::cg = spmv + p_cg_update + "
             "p_jacobi_dot_rz + p_lin2_zy,\n#  cg_device = spmv + p_cgdev_update_jacobi + p_lin2_zy<dev>)")
    f = full([os.path.join(G, "spmv7_r1_final.ncu-rep"), os.path.join(G, "ew_r1_final.ncu-rep")],
             os.path.join(P, "r1_ncu_full_summary.txt"),
             "# round 1: ncu --set full --clock-control none --import-source on (bench.py --steps 20 --warmup 5), Jacobi-CG 7-pt 256^3")
    if "spmv" in f:
        s = f["spmv"]
        s["algorithmic_bytes_per_launch"] = 12 * 117047296 + 4 * (16777216 + 1) + 16 * 16777216
        s["traffic_over_algorithmic"] = s["dram_bytes_per_launch"] / s["algorithmic_bytes_per_launch"]
        json.dump(s, open(os.path.join(P, "r1_spmv_ncu_summary.json"), "w"), indent=1)
        print(s)


if __name__ == "__main__":
    sys.exit(main())
