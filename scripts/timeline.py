"""Kernel timeline of the steady-state pipelined step (D batches in flight, graph replay), from CUPTI activity records
via torch.profiler: per kernel the duration INSIDE the mix (vs. its lone-launch time in the ncu lists), the number of
kernels in flight over time, and how long each stream's chain takes.  Writes gpurun_out/timeline_kernels.csv and prints
a summary (copied to profiles/ by hand).  usage: python scripts/timeline.py [depth] [steps]"""
import collections
import json
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcops_b200 import _lib, synth               # noqa: E402
from pcops_b200.pipeline import ScanNetGeometry   # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 8
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 48
B, N = 16, 8192
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
_lib.set_concurrency_hint(D)     # the streaming kernels size themselves for D batches in flight
x, f = synth.scannet_batch(0, B, N)
dx, df = torch.from_numpy(x).to(dev), torch.from_numpy(f).to(dev)
pipes = [ScanNetGeometry(B, N, 6, dev, attention=True, seed=d, own_streams=True, grid=True) for d in range(D)]
cur = torch.cuda.current_stream(dev)
for pl in pipes:
    pl.set_inputs(dx, df)
    pl.forward(True)
torch.cuda.synchronize()
for pl in pipes:
    pl.capture(True)


def run(steps):
    for pl in pipes:
        pl.main.wait_stream(cur)
    for i in range(steps):
        pl = pipes[i % D]
        pl.set_inputs(dx, df)
        pl.replay()
    for pl in pipes:
        cur.wait_stream(pl.main)


run(4 * D)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile   # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(STEPS)
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/timeline_trace.json")
tr = json.load(open("gpurun_out/timeline_trace.json"))
ev = [e for e in tr["traceEvents"] if e.get("cat") == "kernel"]
os.remove("gpurun_out/timeline_trace.json")
if not ev:
    print("no kernel records (CUPTI unavailable?)")
    sys.exit(0)


def short(n):
    n = n.replace("void ", "").replace("pc::(anonymous namespace)::", "").replace("pc::<unnamed>::", "")
    n = re.sub(r"\(.*", "", n)
    return re.sub(r"<.*", lambda m: m.group(0)[:24], n)


t0 = min(e["ts"] for e in ev)
t1 = max(e["ts"] + e["dur"] for e in ev)
span = t1 - t0
with open("gpurun_out/timeline_kernels.csv", "w") as fh:
    fh.write("start_us,dur_us,stream,grid,block,regs,smem,kernel\n")
    for e in sorted(ev, key=lambda e: e["ts"]):
        a = e.get("args", {})
        fh.write("%.2f,%.2f,%s,%s,%s,%s,%s,%s\n" % (e["ts"] - t0, e["dur"], a.get("stream"), "x".join(map(str, a.get("grid", []))),
                                                  "x".join(map(str, a.get("block", []))), a.get("registers per thread"),
                                                  a.get("shared memory"), short(e["name"])))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for e in ev:
    a = e.get("args", {})
    k = (short(e["name"]), "x".join(map(str, a.get("grid", []))))
    agg[k][0] += 1
    agg[k][1] += e["dur"]
    g = a.get("grid", [1, 1, 1])
    agg[k][2] = g[0] * g[1] * g[2]
print("depth %d, %d steps: %d kernels in %.1f us -> %.1f us per step, %.0f scenes/s" % (D, STEPS, len(ev), span, span / STEPS, B * STEPS / span * 1e6))
print("%-58s %-12s %6s %9s %11s" % ("kernel", "grid", "n/step", "mean us", "us per step"))
for k, (n, tot, ctas) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-58s %-12s %6.2f %9.1f %11.1f" % (k[0][:58], k[1], n / STEPS, tot / n, tot / STEPS))
# kernels in flight, sampled every microsecond over the middle half of the window
edges = []
for e in ev:
    edges.append((e["ts"], 1))
    edges.append((e["ts"] + e["dur"], -1))
edges.sort()
hist, live, last = collections.Counter(), 0, None
lo, hi = t0 + span * 0.25, t0 + span * 0.75
for t, d in edges:
    if last is not None and t > lo and last < hi:
        hist[live] += min(t, hi) - max(last, lo)
    live += d
    last = t
tot = sum(hist.values())
print("kernels in flight (share of the middle half of the window):", {k: round(v / tot, 3) for k, v in sorted(hist.items())})
# per pipeline instance: time from the first kernel of a replay to its last (chain latency under load)
by_stream = collections.defaultdict(list)
for e in ev:
    by_stream[e.get("args", {}).get("stream")].append(e)
# chain structure: per FPS SA1 launch, the time to the end of the last kernel of that replay is not recoverable from
# stream ids alone (two streams per instance), so report the gaps on the FPS streams instead
fps_streams = {e.get("args", {}).get("stream") for e in ev if "fps_pruned" in e["name"]}
for sid in sorted(fps_streams, key=str):
    es = sorted(by_stream[sid], key=lambda e: e["ts"])
    busy = sum(e["dur"] for e in es)
    print("stream %s: %d kernels, busy %.0f us of %.0f (%.2f)" % (sid, len(es), busy, span, busy / span))
fps = [e for e in ev if "fps_pruned" in e["name"]]
print("fps_pruned launches %d, mean %.1f us" % (len(fps), sum(e["dur"] for e in fps) / max(1, len(fps))))
