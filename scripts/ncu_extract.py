#!/usr/bin/env python
"""Turns an `ncu --set full` report into the committed evidence under profiles/:

    python scripts/ncu_extract.py gpurun_out/r2_full.ncu-rep r2

writes profiles/<tag>_ncu_table.md (one row per captured launch: duration, DRAM bytes, issue slots, top pipe, occupancy,
registers), profiles/<tag>_ncu_details_<kernel>.txt (`--page details` of the first launch of each kernel) and updates
profiles/ncu_summary.json, the file bench.py reads `roofline.traffic` and the ncu figures of its timed kernel from."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, k, default=None):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return default


def to_unit(v, unit, want):
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return None if v is None else v * scale[unit] / scale[want]


PIPES = [k for k in hdr if re.match(r"sm__inst_executed_pipe_[a-z0-9_]+\.avg\.pct_of_peak_sustained_active$", k)]
table = []
for r in data:
    name = re.sub(r"^void ", "", r[ix["Kernel Name"]])
    short = re.sub(r"\(.*", "", name).replace("unnamed>::", "").replace("pc::<", "")
    dur = to_unit(num(r, "gpu__time_duration.sum"), units[ix["gpu__time_duration.sum"]], "us")
    rd = to_unit(num(r, "dram__bytes_read.sum"), units[ix["dram__bytes_read.sum"]], "byte")
    wr = to_unit(num(r, "dram__bytes_write.sum"), units[ix["dram__bytes_write.sum"]], "byte")
    pipes = sorted(((num(r, k, 0.0), k.split("pipe_")[1].split(".")[0]) for k in PIPES), reverse=True)
    row = {"id": int(r[ix["ID"]]), "kernel": short, "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]],
           "duration_us": dur, "dram_bytes": None if rd is None else int(rd + (wr or 0)),
           "dram_gbs": None if not dur else (rd + (wr or 0)) / dur / 1e3,
           "issue_slots_busy_pct": num(r, "sm__inst_issued.avg.pct_of_peak_sustained_active"),
           "top_pipe": pipes[0][1] if pipes else None, "top_pipe_pct": pipes[0][0] if pipes else None,
           "achieved_occupancy_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
           "registers": num(r, "launch__registers_per_thread"),
           "sm_cycles": num(r, "sm__cycles_elapsed.max")}
    table.append(row)

out = ["# `ncu --set full --clock-control none` captures, %s (scripts/ncu_targets.py, B = 16 bench shapes)" % tag, "",
       "Durations are cold-cache single launches under the profiler: compare shares and counters, not absolutes.", "",
       "| id | kernel | grid | block | us | DRAM MB | DRAM GB/s | issue slots % | top pipe | occupancy % | regs |",
       "|---|---|---|---|---|---|---|---|---|---|---|"]
for t in table:
    out.append("| %d | %s | %s | %s | %.1f | %.2f | %.0f | %.1f | %s %.1f %% | %.1f | %d |" % (
        t["id"], t["kernel"], t["grid"], t["block"], t["duration_us"], (t["dram_bytes"] or 0) / 1e6, t["dram_gbs"] or 0,
        t["issue_slots_busy_pct"] or 0, t["top_pipe"], t["top_pipe_pct"] or 0, t["achieved_occupancy_pct"] or 0,
        t["registers"] or 0))
open(os.path.join(ROOT, "profiles", "%s_ncu_table.md" % tag), "w").write("\n".join(out) + "\n")

# details pages, one per distinct kernel
seen = set()
for t in table:
    base = re.sub(r"<.*", "", t["kernel"])
    if base in seen:
        continue
    seen.add(base)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--kernel-name", base, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    open(os.path.join(ROOT, "profiles", "%s_ncu_details_%s.txt" % (tag, base)), "w").write(det)

# bench.py's evidence file: op name -> figures of the capture of the kernel that op launches at B = 16
path = os.path.join(ROOT, "profiles", "ncu_summary.json")
summary = json.load(open(path)) if os.path.exists(path) else {}
want = {"fps_sa1": ("fps_pruned_kernel", "(16, 1, 1)"), "query_ball_sa1": ("ball_query_tile_kernel", "(32, 16, 1)"),
        "three_nn_fp4": ("three_nn_tile_kernel", "(64, 16, 1)"), "grid_build": ("grid_build_kernel", "(16, 2, 1)"),
        "three_interpolate_fp4": ("interp_vec4_kernel", None), "csr_build_fp4": ("csr_build_kernel", None),
        "csr_reduce_fp4": ("csr_reduce_stream_kernel<4, 1>", None), "csr_reduce_sa2": ("csr_reduce_stream_kernel<2, 0>", None),
        "knn_sa1": ("knn_kernel", None), "fps_cluster_65536": ("fps_cluster_kernel<8", None),
        "fps_coop_1m": ("fps_cluster_kernel<0", None)}
for op, (needle, grid) in want.items():
    for t in table:
        if needle in t["kernel"] and (grid is None or t["grid"] == grid):
            e = {k: t[k] for k in ("kernel", "grid", "block", "duration_us", "dram_bytes", "issue_slots_busy_pct", "top_pipe",
                                   "top_pipe_pct", "achieved_occupancy_pct", "registers", "sm_cycles")}
            e["capture"] = "profiles/%s_ncu_table.md id %d (ncu --set full, cold cache)" % (tag, t["id"])
            if op == "fps_sa1" and t["sm_cycles"]:
                e["cycles_per_round"] = t["sm_cycles"] / 1023.0
            summary[op] = e
            break
json.dump(summary, open(path, "w"), indent=1)
print("wrote", len(table), "rows;", sorted(summary))
