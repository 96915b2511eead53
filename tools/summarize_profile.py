#!/usr/bin/env python
"""Turn the ncu artefacts of one gpurun call into the committed summaries under profiles/ (development tool).

    python tools/summarize_profile.py <launches.csv|-> <full.ncu-rep> <tag> [workload [cells [kernel]]]
Writes profiles/<tag>_launches_ncu_gputime.csv (copy), profiles/<tag>_k_hru_step_nn3_ncu_raw_selected.csv,
profiles/<tag>_k_hru_step_nn3_by_function.txt and profiles/traffic.json (DRAM bytes and FP64 flops per launch, read by bench.py)."""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
workload = sys.argv[4] if len(sys.argv) > 4 else "fe_hourly"
cells = int(sys.argv[5]) if len(sys.argv) > 5 else 10000
kernel = sys.argv[6] if len(sys.argv) > 6 else "k_hru_step_nn3"
P = os.environ.get("PROFILE_OUT") or os.path.join(ROOT, "profiles")  # PROFILE_OUT: summarise on the GPU box into gpurun_out/
os.makedirs(P, exist_ok=True)
if launches != "-":
    shutil.copy(launches, os.path.join(P, f"{tag}_launches_ncu_gputime.csv"))
    rows = list(csv.reader(l for l in open(launches) if not l.startswith("==")))
    hdr, data = rows[0], rows[1:]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in data:
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu].startswith("ns") else v * 1e3 if r[iu].startswith("ms") else v
        agg[r[ik].split("(")[0].replace("void ", "").split("<")[0]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print("launch list (serialised): kernel, launches, avg us, share %")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {k:50s} {len(v):4d} {sum(v) / len(v):9.1f} {100 * sum(v) / tot:6.2f}")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
alldata = data
data = [r for r in data if kernel.split("_nn")[0] in r[hdr.index("Kernel Name")]]  # the step kernel's launches (the report may hold k_cell_output too)
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.per_cycle_active", "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
        "l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "memory_l2_theoretical_sectors_local", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "launch__shared_mem_config_size", "sm__cycles_elapsed.max",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"]
sel = [i for i, h in enumerate(hdr) if h in want or ("issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h)]
with open(os.path.join(P, f"{tag}_{kernel}_ncu_raw_selected.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "") for r in alldata])
    for i in sel:
        w.writerow([hdr[i], units[i]] + [r[i] for r in alldata])
        print(f"  {hdr[i]:75s} {units[i]:10s} {[r[i] for r in alldata]}")


def col(n):
    return [float(r[hdr.index(n)]) for r in data]


mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
traffic = [(a + b) * mul for a, b in zip(col("dram__bytes_read.sum"), col("dram__bytes_write.sum"))]
flops = [(2 * a + b + c) * cy for a, b, c, cy in zip(col("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"),
                                                     col("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"),
                                                     col("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"), col("sm__cycles_elapsed.max"))]
tj = os.path.join(P, "traffic.json")
try:
    allt = json.load(open(tj))
except Exception:
    allt = {}
allt[f"{kernel}:{workload}"] = {"dram_bytes_per_launch": sum(traffic) / len(traffic), "fp64_flop_per_launch": sum(flops) / len(flops), "cells": cells,
                                "source": f"profiles/{tag}_{kernel}_ncu_raw_selected.csv (ncu --set full, {len(data)} launch(es) of bench.py --workload {workload} "
                                          f"--no-e2e --no-cpu-baseline, {cells} cells; flops = 2 DFMA + DADD + DMUL, thread level)"}
json.dump(allt, open(tj, "w"), indent=1)
print("traffic per launch [MB]", [t / 1e6 for t in traffic], "fp64 flop per launch", flops)
nn = kernel.split("_nn")[1]
env = dict(os.environ, NCU_KERNEL=f"k_hru_step_nn{nn}ILb1")
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_function.py"), rep, os.path.join(ROOT, "vic_b200", "lib", "obj", f"vicgpu_step_nn{nn}.o"), "30"],
                     capture_output=True, text=True, env=env).stdout
open(os.path.join(P, f"{tag}_{kernel}_by_function.txt"), "w").write(out)
print(out)
