#!/usr/bin/env python
"""Human-readable digest of a bench.py JSON line: python scripts/show_bench.py gpurun_out/bench.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.0f %s  ms/step %.4f  (x%d repeats)  e2e %.0f (%.4f ms, h2d %.1f GB/s d2h %.1f GB/s)" % (
    d["value"], d["unit"], d["ms_per_step"], d.get("timed_repeats", 1), d["e2e"]["value"], d["e2e"]["ms_per_step"],
    d["e2e"].get("h2d_gbs") or 0, d["e2e"].get("d2h_gbs") or 0))
print("clocks", d.get("clocks"))
print("cpu_baseline", d.get("cpu_baseline"))
r = d["roofline"]
print("roofline", {k: r[k] for k in ("kernel", "bound", "achieved", "peak", "frac", "traffic", "ms") if k in r})
print("  step", r.get("step"))
for k, v in (d.get("rooflines") or {}).items():
    print("  %-30s %8.1f us %-5s frac %.3f" % (k, v["ms"] * 1e3, v["bound"], v["frac"]))
print("grid variants", {k: round(v * 1e3, 1) for k, v in (d.get("grid_variants_ms") or {}).items()})
for k in ("gathers_steady_state", "attention_layer_tcgen05", "reference_gpu_kernels", "config1_single_scene_sa1",
          "with_attention_layers", "full_model_inference", "config3_training_step", "config4_whole_scene"):
    print(k, json.dumps(d.get(k))[:900])
c5 = d.get("config5_sweep") or {}
print("config5 total_ms", c5.get("total_ms"), "failed", c5.get("failed"), c5.get("error"))
for row in c5.get("rows", []):
    if row.get("ms"):
        print("  %-5s n=%-8d m=%-6d %10.2f ms  %8.1f Gpairs/s  frac %.3f" % (
            row["op"], row["n"], row["npoint"], row["ms"], row["gpairs_per_s"], row["frac_fp32"]))
    else:
        print("  %-5s n=%-8d m=%-6d failed" % (row["op"], row["n"], row["npoint"]))
print("summary", d.get("summary"))
