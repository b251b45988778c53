#!/usr/bin/env python
"""bench.py -- scenes/s of the PointNet++ ScanNet geometry hot path (4 SA + 4 FP + attention contraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of B=16 synthetic 8192-point ScanNet-shaped chunks (xyz + 6
feature channels), BASELINE.json configs[1] plus the attention contraction the metric names: per SA level
FPS -> gather_point -> query_ball_point -> group_point(xyz) -> group_point(features) -> attention contraction, per FP
level three_nn -> weights -> three_interpolate (36 reference-signature op calls; 46 kernel launches with the cell-grid
neighbour search and the fused FPS+gather; pointcloud-segmentation-attention_b200/pipeline.py).

Own arm (default).  One process per GPU, scenes sharded by rank, no data-path collective (weak scaling).  Prints ONE
JSON line on rank 0:
  value         scenes/s, inputs resident in HBM, K steps timed with CUDA events, max over ranks.  Every step reads a
                different input batch (ring of R batches); one step touches > 500 MB (> 126 MB L2).  Steps are
                independent batches, so --depth of them are in flight at once, each on its own streams and buffers
                (FPS is a latency-bound chain on B SMs; the other SMs work on neighbouring batches meanwhile).
  e2e           same metric through host buffers: per step H2D of the batch from pinned memory, the forward, ONE D2H
                of the integer geometry results (FPS / ball / three_nn indices, counts) to pinned memory.
  roofline      the kernel with the largest share of a step, duration from CUDA events on its own stream, against
                its own bound; `rooflines` lists every op (probed eager pass of one pipeline instance).
  cpu_baseline  the CPU oracle (C port of the reference algorithms) on this box's host cores, bounded sample.
Reference arm (--impl reference): the reference's own CPU code (oracle/_ref, compiled from /root/reference sources)
where the reference has CPU code for an op, the C port elsewhere, on all host threads, same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "8192-pt scenes/sec (4 SA + 4 FP + attn)"
UNIT = "scenes/s"
WORKLOAD = "ScanNet semseg geometry forward, B=16x8192 xyz+6ch, 4 SA (1024/256/64/16, r=0.1/0.2/0.4/0.8, k=32) + 4 FP + attention contraction"
NPOINTS = 8192


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)  # ~0.55 s timed region: several 100 ms clock samples fall inside it
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="scenes per GPU per step (the config names 16)")
    ap.add_argument("--ring", type=int, default=4, help="distinct input batches cycled through")
    ap.add_argument("--graph", type=int, default=1, help="1: replay each forward as a CUDA graph (default); 0: eager")
    ap.add_argument("--depth", type=int, default=8, help="independent batches in flight (pipeline instances, own streams)")
    ap.add_argument("--attention", type=int, default=1, help="0: leave the attention contraction out (diagnostics only)")
    ap.add_argument("--fuse-layers", type=int, default=0, help="1: pc_sa_group / pc_fp_interpolate instead of the op pairs")
    ap.add_argument("--grid", type=int, default=1, help="1: cell-grid ball query / three_nn; 0: all-pairs kernels")
    ap.add_argument("--no-overlap", action="store_true", help="single stream")
    ap.add_argument("--cpu-scenes", type=int, default=0, help="scenes in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--attention-layers", type=int, default=1,
                    help="1: also time the forward with the whole AttentionLayer (Dense Q/K/V + contraction on tcgen05) per level")
    ap.add_argument("--scenes", type=int, default=6, help="whole scans per GPU in the config-4 region (0: skip)")
    ap.add_argument("--train", type=int, default=1, help="1: also time config 3 (forward + the registered gradients)")
    ap.add_argument("--train-depth", type=int, default=4, help="batches in flight for the config-3 region")
    ap.add_argument("--skip-probe", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks sampling
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [ln for (t, ln) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [ln for (_, ln) in self.lines]
        sm, mx, reasons = [], [], set()
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arms
def _cpu_standins(rng_seed=7):
    """Per-scene stand-ins for the dense-layer outputs the geometry ops consume (same role as in pipeline.py)."""
    import numpy as np
    from pcops_b200.pipeline import SA_LEVELS
    rng = np.random.Generator(np.random.PCG64(rng_seed))
    st, cin = {}, 6
    for li, (m, r, ns, cout) in enumerate(SA_LEVELS):
        n = NPOINTS if li == 0 else SA_LEVELS[li - 1][0]
        st["feat%d" % li] = None if li == 0 else rng.standard_normal((1, n, cin), dtype=np.float32)
        st["Q%d" % li] = rng.standard_normal((m, cout), dtype=np.float32)
        st["K%d" % li] = rng.standard_normal((m, ns, cout), dtype=np.float32)
        st["V%d" % li] = rng.standard_normal((m, ns, cout), dtype=np.float32)
        cin = cout
    for li, c in ((3, 512), (2, 256), (1, 256), (0, 128)):
        st["p2_%d" % li] = rng.standard_normal((1, SA_LEVELS[li][0], c), dtype=np.float32)
    return st


def cpu_scene_forward(xyz, feat, st, use_ref):
    """The same 36 op calls for ONE scene on the host.  use_ref: call the reference's own compiled CPU functions
    (oracle/_ref) for the ops the reference implements on the CPU (ball query, group_point:
    grouping/test/query_ball_point.cpp:19-84; three_nn, three_interpolate: tf_interpolate.cpp:60-127); FPS, gather and
    the attention contraction have no CPU implementation in the reference -> C port (oracle/pcops_oracle.c)."""
    from oracle import cpu, ref
    from pcops_b200.pipeline import KEY_DIM, SA_LEVELS
    cur, chk = xyz, 0
    fp_in = {}
    for li, (m, r, ns, cout) in enumerate(SA_LEVELS):
        fi = cpu.farthest_point_sample(m, cur)
        new_xyz = cpu.gather_point(cur, fi)
        f = feat if li == 0 else st["feat%d" % li]
        if use_ref:
            idx = ref.cpu_query_ball_point(r, ns, cur, new_xyz)
            ref.cpu_group_point(cur, idx)
            ref.cpu_group_point(f, idx)
        else:
            idx, _ = cpu.query_ball_point(r, ns, cur, new_xyz)
            cpu.group_point(cur, idx)
            cpu.group_point(f, idx)
        cpu.attention_fwd(st["Q%d" % li], st["K%d" % li], st["V%d" % li], cout // KEY_DIM, KEY_DIM)
        fp_in[li] = (cur, new_xyz)
        chk += int(fi.sum()) + int(idx.sum())
        cur = new_xyz
    for li in (3, 2, 1, 0):
        x1, x2 = fp_in[li]
        if use_ref:
            d, i3 = ref.cpu_three_nn(x1, x2)
            w = cpu.three_weights(d)
            ref.cpu_three_interpolate(st["p2_%d" % li], i3, w)
        else:
            d, i3 = cpu.three_nn(x1, x2)
            w = cpu.three_weights(d)
            cpu.three_interpolate(st["p2_%d" % li], i3, w)
        chk += int(i3.sum())
    return chk


def cpu_scenes_per_s(scenes_xyz, scenes_feat, threads, use_ref):
    """Runs the per-scene host forward over the sample on `threads` host threads (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    st = _cpu_standins()
    n = scenes_xyz.shape[0]

    def one(i):
        return cpu_scene_forward(scenes_xyz[i:i + 1], scenes_feat[i:i + 1], st, use_ref)
    t0 = time.perf_counter()
    if threads <= 1:
        for i in range(n):
            one(i)
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(one, range(n)))
    dt = time.perf_counter() - t0
    return n / dt, dt


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def time_fused_attention_layer(torch, ops, B):
    """The fused AttentionLayer (Dense Q/K/V + contraction on tcgen05: csrc/attention_layer.cu for C = 64,
    csrc/attention_layer_wide.cu for C = 128 / 256 / 512) at the four ScanNet attention levels of a B-scene batch, next to
    the composition it replaces (three fp32 cuBLAS GEMMs + pc_attention_fwd); GPU time of a CUDA-graph replay."""
    from pcops_b200.attention_layer import attention_contract, attention_layer_fused
    from pcops_b200.pipeline import SA_LEVELS
    g = torch.Generator(device="cuda").manual_seed(5)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False

    def graph_ms(fn):
        fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        ts = []
        for _ in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]
    out = {"tensor_peak_note": "dense tf32 nominal 1.1 PFLOP/s; the 3xTF32 split triples the issued flops", "levels": {}}
    try:
        for li, (m, _r, S, C) in enumerate(SA_LEVELS):
            G = B * m
            x = torch.randn(G, S, C, generator=g, device="cuda")
            xq = x[:, 0, :].contiguous()
            W = [torch.randn(C, C, generator=g, device="cuda") / C ** 0.5 for _ in range(3)]
            b = [torch.randn(C, generator=g, device="cuda") * 0.1 for _ in range(3)]

            def comp():
                return attention_contract(xq @ W[0] + b[0], x @ W[1] + b[1], x @ W[2] + b[2], C // 4, 4)

            def fused():
                return attention_layer_fused(xq, x, W[0], b[0], W[1], b[1], W[2], b[2])
            a, c = fused(), comp()
            err = float(((a - c).abs().max() / c.abs().max()).item())
            tf, tc = graph_ms(fused), graph_ms(comp)
            flops = 3 * 2.0 * G * S * C * 2 * C  # three UMMA passes over the K|V projection
            out["levels"]["sa%d" % (li + 1)] = {
                "shape": "G=%d S=%d C=%d heads=%d key_dim=4" % (G, S, C, C // 4), "fused_ms": tf,
                "fp32_cublas_composition_ms": tc, "speedup": tc / tf, "max_err_over_max_abs_vs_fp32": err,
                "issued_tf32_tflops": flops / (tf * 1e-3) / 1e12}
            del x, xq, W, b
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    lv = out["levels"]
    out["fused_ms_all_levels"] = sum(v["fused_ms"] for v in lv.values())
    out["composition_ms_all_levels"] = sum(v["fp32_cublas_composition_ms"] for v in lv.values())
    return out


def time_config1(torch, ops, xyz1, feat1, with_cpu):
    """BASELINE config 1: a single synthetic 8192-point scene, B = 1, SA1 only -- FPS npoint=1024, gather, ball query
    r=0.1 nsample=32, group_point of xyz and features (sample_and_group, pointnet_util.py:16-58).  GPU: latency of one
    call through the wrapper (5 launches, eager, CUDA events, median of 20).  CPU: the same ops of the C port, 1 thread."""
    x, f = xyz1.cuda(), feat1.cuda()
    for _ in range(3):
        ops.sample_and_group(1024, 0.1, 32, x, f)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.sample_and_group(1024, 0.1, 32, x, f)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out = {"workload": "config 1: one 8192-point scene, SA1 only (FPS 1024, ball r=0.1 k=32, group xyz+6ch)",
           "gpu_ms": sorted(ts)[len(ts) // 2], "gpu_ms_min": min(ts)}
    if with_cpu:
        from oracle import cpu  # test infrastructure, used here only as the timed CPU baseline
        xn, fn = xyz1.numpy(), feat1.numpy()
        t0 = time.perf_counter()
        fi = cpu.farthest_point_sample(1024, xn)
        nx = cpu.gather_point(xn, fi)
        idx, _ = cpu.query_ball_point(0.1, 32, xn, nx)
        cpu.group_point(xn, idx)
        cpu.group_point(fn, idx)
        out["cpu_ms"] = 1e3 * (time.perf_counter() - t0)
        out["cpu_kind"] = "port (C restatement, 1 thread)"
        out["speedup_vs_cpu"] = out["cpu_ms"] / out["gpu_ms"]
    return out


def time_steady_state_gathers(torch, ops, hbm_peak):
    """The HBM-bound gathers at a size where launch ramp and tail are amortised: B=64 scenes per launch (config 5's batch),
    8 launches back to back on rotating buffer sets (> 126 MB L2 between reuses), one CUDA-event pair around all of them.
    Fraction = algorithmic bytes (SURVEY.md 8d formulas) / time / measured HBM copy bandwidth."""
    import ctypes
    L, p = ops._lib.lib(), ops._lib.ptr
    dev, B, reps = torch.device("cuda"), 64, 8
    g = torch.Generator(device=dev).manual_seed(11)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}

    def timed(name, nbytes, launch, nsets):
        for r in range(nsets):
            launch(r)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            launch(r % nsets)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name] = {"ms": ms, "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak, "bytes": nbytes, "batch": B}

    # group_point, SA2 features: points (B,1024,64), idx (B,256,32) -> (B,256,32,64)
    for tag, n, m, ns, c in (("group_point_sa2_c64", 1024, 256, 32, 64), ("group_point_sa3_c128", 256, 64, 32, 128)):
        nsets = 4
        pts = [torch.randn(B, n, c, generator=g, device=dev) for _ in range(nsets)]
        idx = [torch.randint(0, n, (B, m, ns), generator=g, device=dev, dtype=torch.int32) for _ in range(nsets)]
        dst = [torch.empty(B, m, ns, c, device=dev) for _ in range(nsets)]
        nbytes = B * (4 * m * ns + 4 * min(n, m * ns) * c + 4 * m * ns * c)
        timed(tag, nbytes, lambda r: L.pc_group_point(B, n, c, m, ns, p(pts[r]), p(idx[r]), p(dst[r]), st), nsets)
        del pts, idx, dst
    # three_interpolate, FP4: points (B,1024,128), idx/weight (B,8192,3) -> (B,8192,128)
    for tag, n, m, c in (("three_interpolate_fp4_c128", 8192, 1024, 128), ("three_interpolate_fp3_c256", 1024, 256, 256)):
        nsets = 3
        pts = [torch.randn(B, m, c, generator=g, device=dev) for _ in range(nsets)]
        idx = [torch.randint(0, m, (B, n, 3), generator=g, device=dev, dtype=torch.int32) for _ in range(nsets)]
        w = [torch.rand(B, n, 3, generator=g, device=dev) for _ in range(nsets)]
        dst = [torch.empty(B, n, c, device=dev) for _ in range(nsets)]
        nbytes = B * (24 * n + 4 * m * c + 4 * n * c)
        timed(tag, nbytes, lambda r: L.pc_three_interpolate(B, m, c, n, p(pts[r]), p(idx[r]), p(w[r]), p(dst[r]), st), nsets)
        del pts, idx, w, dst
    torch.cuda.empty_cache()
    return out


def run_config4(torch, np, args, pipes, rank, world, dev, sharding, with_cpu):
    """Config 4: whole-scan inference data path.  Per scan (100-200 k synthetic points, resident in HBM): the GPU chunker
    (complete_scene_loader mirror: cells, shuffle, 8192-point chunks, fill-up, masks, original indices, feature gathers),
    the geometry forward over its chunks in batches of B, and map_back of per-point values (coordinates and labels, as
    generate_predictions.py:162-166 does).  Scans are sharded over ranks; no collective."""
    from pcops_b200 import complete_scene_loader as csl
    from pcops_b200 import synth
    S, B = args.scenes, pipes[0].B
    # This function contains NO collective: a rank-local failure must not leave the other ranks in a barrier.  The caller
    # agrees on the timing (max over ranks) outside its try block.
    # Scans: seeds 1000 + 1000 * rank + k, k = 0, 1, ...  A scan in which some cell holds an exact multiple of 8192
    # points makes the reference raise (complete_scene_loader.py:89-90 concatenates an empty list with a 2-D array) and
    # so does the mirror; such scans (about 1 in 50 here) are not part of the workload: skipped, and counted.
    scans, skipped, k = [], 0, 0
    while len(scans) < S and k < 8 * S + 8:
        p, l, c, n = synth.whole_scene(1000 + 1000 * rank + k)
        k += 1
        f6 = np.concatenate([c.astype(np.float32) / 255.0, n], 1)      # train.py:95-98 (stock TF cast, outside the op path)
        dev_scan = tuple(torch.from_numpy(a).to(dev) for a in (p, l, f6))
        try:
            csl.chunk_scene(dev_scan[0])
        except ValueError:
            skipped += 1
            continue
        scans.append((p, dev_scan))
    if len(scans) < S:
        raise RuntimeError("could not find %d chunkable scans" % S)
    cur = torch.cuda.current_stream(dev)
    use_graph = bool(args.graph)
    stats = {"chunks": 0, "points": 0}

    def one(i):
        p, l, f6 = scans[i % S][1]
        chunks = csl.chunk_scene(p)
        feats = chunks.gather(f6)
        labels = chunks.gather(l)
        C = chunks.nchunks
        nb = (C + B - 1) // B
        keep = []   # batch tensors are produced on `cur` and read on the pipeline streams: hold them until the join
        for b in range(nb):
            pl = pipes[b % len(pipes)]
            sel = torch.arange(b * B, b * B + B, device=dev) % C    # the last batch wraps around (fixed-size pipeline)
            bx, bf = chunks.point_sets[sel], feats[sel]
            keep.append((bx, bf))
            pl.main.wait_stream(cur)
            pl.set_inputs(bx, bf)
            if use_graph:
                pl.replay()
            else:
                pl.forward(True)
        orig, masks = chunks.orig_idx.reshape(-1), chunks.masks.reshape(-1)
        back_p = csl.map_back(chunks.point_sets.reshape(-1, 3), orig, masks, (p.shape[0], 3))
        back_l = csl.map_back(labels.reshape(-1), orig, masks, (p.shape[0],))
        for pl in pipes:
            cur.wait_stream(pl.main)
        del keep
        stats["chunks"] += C
        stats["points"] += int(p.shape[0])
        return back_p, back_l

    np.random.seed(99 + rank)
    bp, bl = one(0)
    torch.cuda.synchronize(dev)
    # fraction of the scan's points whose coordinates come back bit-exact (the reference's float32 height bound can
    # leave the top-most point outside every un-padded cell, complete_scene_loader.py:34,41 -- reproduced, not fixed)
    ok = float((bp == scans[0][1][0]).all(dim=1).float().mean().item())
    stats["chunks"] = stats["points"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    for i in range(S):
        one(i)
    e1.record(cur)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    # the chunker alone (device tensors in, device tensors out)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(cur)
    for i in range(S):
        ch = csl.chunk_scene(scans[i][1][0])
        ch.gather(scans[i][1][2])
    c1.record(cur)
    torch.cuda.synchronize(dev)
    out = {"workload": "config 4: whole-scan inference data path, %d synthetic scans per GPU (100-200 k points), chunker + "
                       "geometry forward over the chunks (B=%d) + map_back" % (S, B),
           "ms_local": ms, "scans_per_gpu": S, "unit": "scans/s", "scans_rejected_like_the_reference": skipped,
           "chunks_per_scan": stats["chunks"] / S, "points_per_scan": stats["points"] / S,
           "chunker_ms_per_scan": c0.elapsed_time(c1) / S, "map_back_restored_fraction": ok}
    if with_cpu:
        from oracle import scene_chunks as osc   # test infrastructure, used here only as the timed CPU baseline
        import time as _t
        p = scans[0][0]
        l, f6 = (t.cpu().numpy() for t in scans[0][1][1:])
        np.random.seed(1)
        t0 = _t.perf_counter()
        osc.chunk_scene(p, [l, f6], True)
        out["cpu_chunker_ms_per_scan"] = 1e3 * (_t.perf_counter() - t0)
        out["cpu_chunker_kind"] = "port (numpy restatement of complete_scene_loader.py, 1 thread, 1 scan)"
    return out


def bind_to_gpu_cpus(index):
    """Pin this rank to the host cores NVML reports as local to GPU `index`, BEFORE any pinned buffer is allocated, so
    first-touch puts the staging memory on the GPU's NUMA node (with 8 ranks the host links are the e2e bottleneck)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + bit for w, mask in enumerate(words) for bit in range(64) if (mask >> bit) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    if rank != 0:
        return 0
    from oracle import cpu, ref
    from pcops_b200 import synth
    cpu.lib()
    use_ref = ref.available_cpu()
    threads = host_threads()
    per_step = max(threads, 8)              # scenes per step: a bounded sample of the B=16 x N-GPU batch
    xyz, feat = synth.scannet_batch(0, per_step, NPOINTS)
    for _ in range(min(args.warmup, 1)):
        cpu_scenes_per_s(xyz[:threads], feat[:threads], threads, use_ref)
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_scenes_per_s(xyz, feat, threads, use_ref)
    dt = time.perf_counter() - t0
    value = per_step * steps / dt
    sample = "%d steps x %d scenes of the B=16 workload, %d host threads, one scene per thread" % (steps, per_step, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "scenes_per_step": per_step, "npoints": NPOINTS,
                   "note": "host-only arm: the reference's geometry ops on the CPU (its interpolation ops are CPU-only; "
                           "ball query / group_point from its standalone CPU programs; FPS, gather, attention contraction: "
                           "C port of the reference CUDA / TF algorithm, the reference has no CPU code for them)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference" if use_ref else "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ own arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist

    import pcops_b200  # noqa: F401  (raises if libpcops.so is missing -- there is no fallback)
    from pcops_b200 import sharding, synth
    from pcops_b200.pipeline import ScanNetGeometry

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the geometry ops have no CPU implementation in the product")
    numa = bind_to_gpu_cpus(local) if world > 1 else None   # pinned staging buffers land on the GPU's own NUMA node
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, R, K, W = args.batch, max(1, args.ring), args.steps, max(args.warmup, 3)

    # scene shards: rank r owns scenes [r*B*R, (r+1)*B*R) of the synthetic scene list
    lo, _hi = sharding.shard_bounds(world * B * R, rank, world)
    host_xyz, host_feat, dev_xyz, dev_feat = [], [], [], []
    for i in range(R):
        x, f = synth.scannet_batch(lo + i * B, B, NPOINTS)
        hx, hf = torch.from_numpy(x).pin_memory(), torch.from_numpy(f).pin_memory()
        host_xyz.append(hx)
        host_feat.append(hf)
        dev_xyz.append(hx.to(dev))
        dev_feat.append(hf.to(dev))

    D = max(1, args.depth)
    pipes = [ScanNetGeometry(B, NPOINTS, 6, dev, attention=bool(args.attention), seed=rank * 64 + d, own_streams=True, grid=bool(args.grid),
                             fuse_layers=bool(args.fuse_layers))
             for d in range(D)]
    pipe = pipes[0]
    overlap = not args.no_overlap
    cur = torch.cuda.current_stream(dev)
    use_graph = bool(args.graph)

    def step_resident(i, probes=None, graph=False):
        pl = pipes[i % D]
        pl.set_inputs(dev_xyz[i % R], dev_feat[i % R])
        if graph:
            pl.replay()
        else:
            pl.forward(overlap, probes)

    def fork():
        for pl in pipes:
            pl.main.wait_stream(cur)

    def join():
        for pl in pipes:
            cur.wait_stream(pl.main)

    # warm-up (also sets per-device kernel attributes before any graph capture)
    for i in range(max(W, D)):
        step_resident(i)
    torch.cuda.synchronize(dev)
    if use_graph:
        for pl in pipes:
            pl.capture(overlap)
        for i in range(2 * D):
            step_resident(i, graph=True)
        torch.cuda.synchronize(dev)

    # ---- timed region 1: inputs resident in HBM ----------------------------------------------------------------
    work = pipe.algorithmic_work()
    top_guess = "fps_sa1"
    probes = {top_guess: []} if not use_graph else None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    sharding.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(cur)
    fork()
    for i in range(K):
        step_resident(i, probes, use_graph)
    join()
    e1.record(cur)
    t_enq = time.time()
    torch.cuda.synchronize(dev)
    t_wall1 = time.time()
    sharding.barrier()
    ms_local = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_total = sharding.max_over_ranks(ms_local)
    value = world * B * K / (ms_total * 1e-3)

    # ---- timed region 2: end to end through host buffers -------------------------------------------------------
    # per step: H2D of the batch from pinned memory, the forward, ONE D2H of the result arena into pinned memory
    host_out = [torch.empty(pl.result_arena().shape, dtype=torch.int32).pin_memory() for pl in pipes]
    h2d = pipe.input_bytes()
    d2h = pipe.result_arena().numel() * 4

    def step_e2e(i):
        pl = pipes[i % D]
        pl.set_inputs(host_xyz[i % R], host_feat[i % R], non_blocking=True)
        if use_graph:
            pl.replay()
        else:
            pl.forward(overlap)
        pl.read_results(host_out[i % D])

    for i in range(max(3, D)):
        step_e2e(i)
    torch.cuda.synchronize(dev)
    sharding.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(cur)
    fork()
    for i in range(K):
        step_e2e(i)
    join()
    f1.record(cur)
    torch.cuda.synchronize(dev)
    sharding.barrier()
    e2e_ms = sharding.max_over_ranks(f0.elapsed_time(f1))
    e2e_value = world * B * K / (e2e_ms * 1e-3)
    checksum = int(sum(int(h.to(torch.int64).sum()) for h in host_out))

    # ---- timed region 3: config 3, a training step's geometry (forward + GroupPointGrad / ThreeInterpolateGrad /
    # attention-contraction backward), inputs resident ----------------------------------------------------------
    train = None
    if args.train and args.attention:
        TD = max(1, min(D, args.train_depth))
        tp = pipes[:TD]
        for pl in tp:
            pl.allocate_backward(seed=4321 + rank)
            pl.set_inputs(dev_xyz[0], dev_feat[0])
            if use_graph:
                pl.capture(overlap, train=True)
            else:
                pl.forward(overlap, train=True)
        torch.cuda.synchronize(dev)

        def step_train(i):
            pl = tp[i % TD]
            pl.set_inputs(dev_xyz[i % R], dev_feat[i % R])
            if use_graph:
                pl.replay()
            else:
                pl.forward(overlap, train=True)
        for i in range(max(3, TD)):
            step_train(i)
        torch.cuda.synchronize(dev)
        sharding.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(cur)
        fork()
        for i in range(K):
            step_train(i)
        join()
        g1.record(cur)
        torch.cuda.synchronize(dev)
        sharding.barrier()
        tr_ms = sharding.max_over_ranks(g0.elapsed_time(g1))
        train = {"workload": "config 3: attention model with 6-ch features, forward + registered gradients "
                             "(GroupPointGrad SA2-4, ThreeInterpolateGrad FP1-4, attention contraction backward SA1-4), "
                             "B=%d x %d" % (B, NPOINTS),
                 "value": world * B * K / (tr_ms * 1e-3), "unit": UNIT, "ms_per_step": tr_ms / K,
                 "batches_in_flight": TD, "gpu_launches": tp[0].launches_per_train_step * K * world}
        if use_graph:   # back to the forward-only graphs for the probes below
            for pl in tp:
                pl.capture(overlap)

    # ---- timed region 3b: the forward with the WHOLE attention layer per level (Dense Q/K/V on the tensor cores +
    # contraction, from stand-in grouped activations) instead of the contraction alone on stand-in K / V -------------
    with_layers = None
    if args.attention_layers and args.attention:
        LD = max(1, min(D, 4))
        lp = [ScanNetGeometry(B, NPOINTS, 6, dev, attention=True, seed=rank * 64 + 32 + d, own_streams=True,
                              grid=bool(args.grid), attention_layers=True) for d in range(LD)]
        for pl in lp:
            pl.set_inputs(dev_xyz[0], dev_feat[0])
            pl.forward(overlap)
        torch.cuda.synchronize(dev)
        if use_graph:
            for pl in lp:
                pl.capture(overlap)

        def step_layers(i):
            pl = lp[i % LD]
            pl.set_inputs(dev_xyz[i % R], dev_feat[i % R])
            if use_graph:
                pl.replay()
            else:
                pl.forward(overlap)
        for i in range(max(3, LD)):
            step_layers(i)
        torch.cuda.synchronize(dev)
        sharding.barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record(cur)
        for pl in lp:
            pl.main.wait_stream(cur)
        for i in range(K):
            step_layers(i)
        for pl in lp:
            cur.wait_stream(pl.main)
        h1.record(cur)
        torch.cuda.synchronize(dev)
        sharding.barrier()
        lay_ms = sharding.max_over_ranks(h0.elapsed_time(h1))
        with_layers = {"workload": "the same forward with pc_attention_layer_fwd (Dense Q/K/V as 3xTF32 tcgen05 UMMA + "
                                   "contraction, K and V never stored) at all four levels instead of the contraction on "
                                   "stand-in K / V; 25.8 GFLOP of projections per step added",
                       "value": world * B * K / (lay_ms * 1e-3), "unit": UNIT, "ms_per_step": lay_ms / K,
                       "batches_in_flight": LD, "gpu_launches": lp[0].launches_per_step * K * world}
        del lp
        torch.cuda.empty_cache()

    # ---- timed region 4: config 4, whole scans through the GPU chunker + forward + map_back -------------------
    config4 = None
    if args.scenes > 0 and args.attention and not args.fuse_layers:
        c4_ms = -1.0
        try:
            config4 = run_config4(torch, np, args, pipes, rank, world, dev, sharding, world == 1 and not args.skip_cpu)
            c4_ms = float(config4.pop("ms_local"))
        except Exception as exc:   # reported, never fatal for the headline numbers
            config4 = {"error": repr(exc)[:300]}
        # every rank takes part in both reductions whether or not its own region succeeded
        c4_max, c4_all_ok = sharding.agree_on_region(c4_ms)
        if "error" not in config4:
            if not c4_all_ok:
                config4 = {"error": "the config-4 region failed on another rank"}
            else:
                S4 = config4["scans_per_gpu"]
                config4["value"] = world * S4 / (c4_max * 1e-3)
                config4["ms_per_scan"] = c4_max / S4

    # ---- probed pass: every op's duration (separate region; events perturb overlap) ----------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    nsm = pcops_b200._lib.lib().pc_num_sms()
    fp32_peak_tops = nsm * 128 * sm_max * 1e6 / 1e12   # un-fused fp32 instructions/s (one op per lane per clock)

    def roof(name, ms):
        wk = work[name]
        if wk["kind"] == "bytes":
            ach = wk["amount"] / (ms * 1e-3) / 1e9
            return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": None, "ms": ms, "peak_source": peak_src}
        ach = wk["amount"] / (ms * 1e-3) / 1e12
        return {"kernel": name, "bound": "fp32", "achieved": ach, "peak": fp32_peak_tops, "unit": "TFLOP/s",
                "frac": ach / fp32_peak_tops, "traffic": None, "ms": ms,
                "peak_source": "%d SMs x 128 fp32 lanes x %.0f MHz, un-fused (1 flop per lane-clock)" % (nsm, sm_max)}

    rooflines, op_ms, grid_ms, fused_layer, steady, config1 = {}, {}, {}, None, None, None
    if not args.skip_probe:
        # Every op's stand-alone duration: one pipeline instance alone, eager, ONE stream.  The table is taken with the
        # reference-signature ops (all-pairs ball query / three_nn, separate gather), whose algorithmic op counts the
        # fractions are computed from; the durations of the cell-grid variants the timed pipeline uses are listed
        # beside them in `grid_variants_ms`.
        def probe(pl):
            allp = {n: [] for n in pl.op_names()}
            for i in range(min(K, 10)):
                pl.set_inputs(dev_xyz[i % R], dev_feat[i % R])
                pl.forward(False, allp)
                torch.cuda.synchronize(dev)
            return {n: sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2] for n, evs in allp.items()}
        ref_pipe = ScanNetGeometry(B, NPOINTS, 6, dev, attention=bool(args.attention), seed=999, own_streams=True,
                                   grid=False, fuse_gather=False, fuse_layers=False)
        for i in range(3):
            ref_pipe.set_inputs(dev_xyz[i % R], dev_feat[i % R])
            ref_pipe.forward(False)
        torch.cuda.synchronize(dev)
        op_ms = probe(ref_pipe)
        for n, ms in op_ms.items():
            rooflines[n] = roof(n, ms)
        del ref_pipe
        if args.grid:
            g = probe(pipe)
            grid_ms = {n: ms for n, ms in g.items() if n.startswith(("query_ball", "three_nn", "fps"))}
        if train is not None:      # the gradient ops alone (single stream, eager), against the HBM roofline
            bw = pipe.backward_work()
            allp = {n: [] for n in pipe.backward_op_names()}
            for i in range(min(K, 10)):
                pipe.set_inputs(dev_xyz[i % R], dev_feat[i % R])
                pipe.forward(False, allp, train=True)
                torch.cuda.synchronize(dev)
            work.update(bw)
            for n, evs in allp.items():
                rooflines[n] = roof(n, sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2])
        try:
            steady = time_steady_state_gathers(torch, pcops_b200, hbm_peak)
        except Exception as exc:
            steady = {"error": str(exc)[:200]}
        try:   # BASELINE config 1: ONE 8192-point scene, SA1 only, through the reference-named wrapper (latency)
            config1 = time_config1(torch, pcops_b200, host_xyz[0][:1], host_feat[0][:1],
                                   rank == 0 and world == 1 and not args.skip_cpu)
        except Exception as exc:
            config1 = {"error": str(exc)[:200]}
        # The fused AttentionLayer (Dense Q/K/V + contraction on tcgen05, csrc/attention_layer.cu) at the SA1 shape, next
        # to the composition it replaces (three fp32 cuBLAS GEMMs + pc_attention_fwd); GPU time of a CUDA-graph replay.
        try:
            fused_layer = time_fused_attention_layer(torch, pcops_b200, B)
        except Exception as exc:  # reported, never fatal for the headline numbers
            fused_layer = {"error": str(exc)[:200]}
    if probes:
        d = [a.elapsed_time(b) for a, b in probes[top_guess]]
        top_ms = sum(d) / len(d)
    elif top_guess in op_ms:
        top_ms = op_ms[top_guess]
    else:                                 # --skip-probe under graph replay: time the dominant kernel alone
        solo = {top_guess: []}
        for i in range(5):
            pipe.set_inputs(dev_xyz[i % R], dev_feat[i % R])
            pipe.forward(False, solo)
        torch.cuda.synchronize(dev)
        d = sorted(a.elapsed_time(b) for a, b in solo[top_guess]) or [float("inf")]
        top_ms = d[len(d) // 2]
    roofline = roof(top_guess, top_ms)
    try:  # DRAM bytes of one launch of this kernel from the committed `ncu --set full` capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(top_guess, {})
        if B == 16 and tr.get("bytes") is not None:
            roofline["traffic"] = tr["bytes"]
            roofline["traffic_source"] = tr.get("capture")
    except Exception:
        pass
    roofline["algorithmic_bytes"] = work[top_guess].get("bytes")
    roofline["occupied_sms"] = min(B, nsm)
    roofline["frac_of_occupied_sms"] = roofline["frac"] * nsm / min(B, nsm)   # FPS runs one scene per SM
    roofline["share_of_step"] = top_ms / sum(op_ms.values()) if op_ms else None
    roofline["timed"] = "CUDA events around each launch on its stream, inside the value region (mean of %d)" % K \
        if probes else "CUDA events around each launch in a probed single-stream eager pass right after the value " \
                       "region (graph replay hides single launches); share = its time / sum of all op times"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline on this box (bounded sample) ------------------------------------------------------------
    cpu_baseline = None
    if world == 1 and not args.skip_cpu:
        from oracle import cpu  # test infrastructure, used here only as the timed CPU baseline
        cpu.lib()
        threads = host_threads()
        x = np.concatenate([t.numpy() for t in host_xyz])
        f = np.concatenate([t.numpy() for t in host_feat])
        v1, dt1 = cpu_scenes_per_s(x[:2], f[:2], 1, False)
        # bounded sample of about 10 s of wall time: the ring's scenes, cycled, sized from the single-thread rate
        ns = args.cpu_scenes or int(min(4096, max(threads, 10.0 * v1 * threads * 0.6)))
        sel = np.arange(ns) % x.shape[0]
        x, f = x[sel], f[sel]
        vN, dtN = cpu_scenes_per_s(x, f, threads, False)
        cpu_baseline = {"value": vN, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "%d scenes drawn from the bench batches, %d host threads (one scene per thread), %.1f s"
                                  % (x.shape[0], threads, dtN),
                        "single_thread": {"value": v1, "cores": 1, "sample": "2 scenes, %.1f s" % dt1}}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "host_enqueue_ms_per_step": 1e3 * (t_enq - t_wall0) / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "npoints": NPOINTS, "feature_channels": 6,
                   "parallelism": "scene-sharded x%d, no collective" % world, "rank_cpu_affinity": numa, "streams_per_batch": 5 if overlap else 1,
                   "cuda_graph": use_graph, "input_ring": R, "batches_in_flight": D,
                   "neighbour_search": "cell grid" if args.grid else "all pairs",
                   "l2": "inputs larger than L2: one step streams >500 MB (K/V/out tensors) through a 126 MB L2; "
                         "each step reads a different batch of scenes"},
        "fps_us_per_scene": {"sa1_batch_latency_us": top_ms * 1e3, "sa1_us_per_scene_throughput": top_ms * 1e3 / B},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / K, "result_checksum": checksum},
        "gpu_launches": pipe.launches_per_step * K * world,
        "clocks": clocks,
        "roofline": roofline,
        "rooflines": rooflines,
        "grid_variants_ms": grid_ms,
        "with_attention_layers": with_layers,
        "config3_training_step": train,
        "gathers_steady_state": steady,
        "config4_whole_scene": config4,
        "config1_single_scene_sa1": config1,
        "attention_layer_tcgen05": fused_layer,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
