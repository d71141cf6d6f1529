#!/usr/bin/env python3
"""Turn one GPU session (tools/gpu_session.sh <tag>, files gpurun_out/<tag>_*) into the tracked evidence under profiles/ (round prefix r02):

  r02_bench.json            the bench line of that session
  r02_launches.csv          ncu launch list of `bench.py --steps 2 --warmup 3` (per launch: duration, DRAM bytes, FP64 instruction counts, warp instructions)
  r02_ncu_full_summary.csv  ncu --set full of the heavy kernels of one warm tick (selected metrics, one row per kernel and metric)
  r02_kernel_traffic.json   DRAM bytes per launch from the launch list (warm launches), per kernel and per tick
  r02_flop_count.json       executed FP64 flops of k_lq_pack per stage (2 DFMA + DMUL + DADD thread-level counts / stages); read by bench.py (roofline.fp64)
  r02_hot_lines.txt         hottest source lines + stall reasons per kernel (ncu source page joined with nvdisasm -gi of the SAME libbmpc.so)

usage: python tools/make_profiles.py <tag> [stages_per_launch]   (run in the build container after the session; needs the libbmpc.so that was profiled)
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
stages = int(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 103
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

shutil.copy(os.path.join(G, f"{tag}_bench.json"), os.path.join(P, "r02_bench.json"))

# ---- launch list
rows = list(csv.reader(l for l in open(os.path.join(G, f"{tag}_launches.csv")) if l.startswith('"')))
hdr, rows = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
launches = collections.OrderedDict()
for r in rows:
    key = int(r[ix["ID"]])
    d = launches.setdefault(key, {"kernel": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
metrics = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
           "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__inst_executed.sum"]
with open(os.path.join(P, "r02_launches.csv"), "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["id", "kernel", "grid", "block", "duration_ns", "dram_read_bytes", "dram_write_bytes", "dfma_thread_inst", "dmul_thread_inst", "dadd_thread_inst", "warp_inst"])
    for k, d in launches.items():
        w.writerow([k, d["kernel"], d["grid"], d["block"]] + [int(d.get(m, 0)) for m in metrics])


def short(name):
    n = name.replace("void ", "")
    return n.split("<")[0].split("(")[0]


# the last complete tick of the run = the last launch of every kernel
last = collections.OrderedDict()
for k, d in launches.items():
    last[short(d["kernel"])] = d
tick = {k: d for k, d in last.items() if k.startswith("k_") and k not in ("k_stage_static", "k_gait_init")}
traffic = {k: int(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) for k, d in tick.items()}
dur = {k: d.get("gpu__time_duration.sum", 0.0) / 1e6 for k, d in tick.items()}
json.dump({"_comment": "per launch of the last tick in profiles/r02_launches.csv (ncu --metrics pass, cold-cache and serialised durations: compare shares); bytes = dram__bytes_read.sum + dram__bytes_write.sum",
           "h1": traffic, "tick_total_bytes": sum(traffic.values()), "duration_ms_under_ncu": dur, "tick_ms_under_ncu": sum(dur.values()),
           "share_of_tick": {k: round(v / max(sum(dur.values()), 1e-9), 4) for k, v in dur.items()}},
          open(os.path.join(P, "r02_kernel_traffic.json"), "w"), indent=1)
lq = tick.get("k_lq_pack")
if lq:
    fl = 2 * lq.get(metrics[3], 0) + lq.get(metrics[4], 0) + lq.get(metrics[5], 0)
    prev = {}
    fp = os.path.join(P, "r02_flop_count.json")
    if os.path.exists(fp):
        prev = json.load(open(fp))
    prev.update({"_comment": "executed FP64 flops of k_lq_pack per stage: (2 x DFMA + DMUL + DADD thread-level instruction counts of one launch, ncu) / stages of the launch; "
                             "dfma_peak_tflops = DFMA microbenchmark on this pool (profiles/r01_fp64_peak_microbench.txt); the oracle's counting scalar gives the model-level "
                             "counts quoted in DESIGN.md section 4",
                 "h1": {"lq_flops_per_stage": fl / stages, "stages": stages, "warp_inst_per_stage": lq.get(metrics[6], 0) / stages}, "dfma_peak_tflops": 35.9})
    json.dump(prev, open(fp, "w"), indent=1)

# ---- ncu --set full summary
want = {
    "gpu__time_duration.sum": "duration", "launch__registers_per_thread": "registers / thread", "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occupancy limit (registers), CTAs", "launch__occupancy_limit_shared_mem": "occupancy limit (shared memory), CTAs",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved occupancy", "smsp__inst_executed.sum": "warp instructions",
    "sm__inst_executed.avg.per_cycle_elapsed": "IPC (elapsed)", "smsp__thread_inst_executed_per_inst_executed.ratio": "active lanes / instruction",
    "dram__bytes_read.sum": "DRAM read", "dram__bytes_write.sum": "DRAM written", "dram__throughput.avg.pct_of_peak_sustained_elapsed": "DRAM throughput",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "FP64 pipe active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "FP64 pipe instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue slots busy", "smsp__warps_eligible.avg.per_cycle_active": "eligible warps / scheduler",
    "smsp__warps_active.avg.per_cycle_active": "active warps / scheduler", "l1tex__t_sector_hit_rate.pct": "L1 hit rate", "lts__t_sector_hit_rate.pct": "L2 hit rate",
    "launch__shared_mem_per_block_dynamic": "dynamic shared memory / CTA", "launch__shared_mem_per_block_static": "static shared memory / CTA",
    "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active": "DMMA pipe instructions",
}
raw = list(csv.reader(open(os.path.join(G, f"{tag}_full_raw.csv"))))
rh, ru, rr = raw[0], raw[1], raw[2:]
rix = {h: i for i, h in enumerate(rh)}
with open(os.path.join(P, "r02_ncu_full_summary.csv"), "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["kernel", "metric", "what", "value", "unit"])
    for r in rr:
        for m, what in want.items():
            if m in rix:
                w.writerow([r[rix["Kernel Name"]], m, what, r[rix[m]], ru[rix[m]]])

# ---- hot lines (needs the profiled libbmpc.so in place)
subprocess.check_call(["bash", os.path.join(ROOT, "tools", "prof_split.sh"), tag], stdout=subprocess.DEVNULL)
mang = {"k_lq_pack": "k_lq_packILi10ELb1E", "k_project": "k_projectILi10ELb1E", "k_riccati_warp": "k_riccati_warpILi10E", "k_policy_expand": "k_policy_expandILi10E",
        "k_forward": "k_forwardILi10E", "k_linesearch": "k_linesearchILi10E"}
with open(os.path.join(P, "r02_hot_lines.txt"), "w") as fh:
    for k, mg in mang.items():
        c = f"/tmp/prof/{k}.csv"
        if not os.path.exists(c):
            continue
        fh.write(f"=== {k}: hottest source lines of the kernel body (ncu --set full source page joined with nvdisasm -gi line info; tools/ncu_line_profile.py --outer, then --depth 1)\n")
        for extra in (["14", "--outer"], ["24", "--depth", "1"]):
            fh.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_line_profile.py"), c, "/tmp/prof/disgi.txt", mg] + extra, capture_output=True, text=True).stdout)
        fh.write("--- stall reasons\n")
        fh.write("\n".join(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_src_summary.py"), c, "0"], capture_output=True, text=True).stdout.splitlines()[:9]) + "\n")
print("tick DRAM bytes (GB):", sum(traffic.values()) / 1e9, {k: round(v / 1e9, 2) for k, v in traffic.items()})
print("k_lq_pack flops/stage:", fl / stages if lq else None)
