#!/usr/bin/env python
"""bench.py -- scenes/s of the PointNet++ ScanNet geometry hot path (4 SA + 4 FP + attention contraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of B=16 synthetic 8192-point ScanNet-shaped chunks (xyz + 6
feature channels), BASELINE.json configs[1] plus the attention contraction the metric names: per SA level
FPS -> gather_point -> query_ball_point -> group_point(xyz) -> group_point(features) -> attention contraction, per FP
level three_nn -> weights -> three_interpolate (36 reference-signature op calls; 39 kernel launches with the cell-grid
neighbour search and the fused FPS+gather; pointcloud-segmentation-attention_b200/pipeline.py).

Own arm (default).  One process per GPU, scenes sharded by rank, no data-path collective (weak scaling).  Prints ONE
JSON line on rank 0:
  value         scenes/s, inputs resident in HBM, K steps timed with CUDA events, max over ranks.  Every step reads a
                different input batch (ring of R batches); one step touches > 500 MB (> 126 MB L2).  Steps are
                independent batches, so --depth of them are in flight at once, each on its own streams and buffers
                (FPS is a latency-bound chain on B SMs; the other SMs work on neighbouring batches meanwhile).
  e2e           same metric through host buffers, one CUDA graph per step: ONE H2D of the packed batch from pinned
                memory (xyz | normals | colours as the uint8 they are stored as, 27 B/point; the /255 of train.py:95
                runs on the device), the forward, the integer geometry results (FPS / ball / three_nn indices,
                counts -- everything a host consumer needs to rebuild any gathered tensor) narrowed to uint16 on
                the device (lossless: n <= 8192) and ONE D2H into pinned memory.
  roofline      the kernel with the largest share of a step, duration from CUDA events on its own stream, against
                its own bound; `rooflines` lists every op (probed eager pass of one pipeline instance).
  cpu_baseline  the CPU oracle (C port of the reference algorithms) on this box's host cores, bounded sample.
Every timed region runs its K steps `repeats` times back to back inside ONE event pair, repeats chosen so the region
lasts >= --min-seconds (0.5 s) whatever --steps is; ms_per_step = region / (K * repeats).
Reference arm (--impl reference): the reference's own CPU code (oracle/_ref, compiled from /root/reference sources)
where the reference has CPU code for an op, the C port elsewhere, on all host threads, same metric, same config keys,
honouring --steps / --warmup (a step is a bounded sample of the batch, sized so that the run ends within minutes).
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
    ap.add_argument("--scenes", type=int, default=39,
                    help="whole scans per GPU in the config-4 region (0: skip); 39 x 8 GPUs = the 312 scans of the validation split")
    ap.add_argument("--full-model", type=int, default=1,
                    help="1: also time the WHOLE attention-model inference forward (4 SA-attention + 4 FP levels + fc, every "
                         "dense layer on the tcgen05 engine, no stand-in tensors)")
    ap.add_argument("--config5", type=int, default=1, help="1: the config-5 sweep (FPS / ball / kNN, N = 16k..1M, npoint 1k..16k, B = 64 sharded)")
    ap.add_argument("--sweep-batch", type=int, default=64, help="scenes of the config-5 sweep, sharded over the ranks")
    ap.add_argument("--sweep-max-n", type=int, default=1 << 20)
    ap.add_argument("--min-seconds", type=float, default=0.5, help="minimum duration of every headline timed region")
    ap.add_argument("--lib", default="", help="path of an alternative libpcops build (A/B runs of kernel variants)")
    ap.add_argument("--e2e-legacy", type=int, default=0, help="1: the round-1 e2e path (two fp32 H2D copies, int32 D2H, no graph)")
    ap.add_argument("--train", type=int, default=1, help="1: also time config 3 (forward + the registered gradients)")
    ap.add_argument("--train-depth", type=int, default=4, help="batches in flight for the config-3 region")
    ap.add_argument("--skip-probe", action="store_true")
    ap.add_argument("--hint", type=int, default=0, help="pc_set_concurrency_hint for the pipelined regions (0: the number of "
                    "batches in flight of each region); lone-launch probes always run with 1")
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

    def consume(i, chunks):
        p, l, f6 = scans[i % S][1]
        feats = chunks.gather(f6)
        labels = chunks.gather(l)
        C = chunks.nchunks
        nb = (C + B - 1) // B
        keep = [chunks]   # tensors produced on `cur` / the chunker stream and read on the pipeline streams: hold them until the join
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
        stats["chunks"] += C
        stats["points"] += int(p.shape[0])
        return back_p, back_l, keep

    def run(n_scans):
        """n_scans scans through chunker -> forward -> map_back.  The chunker is pipelined over scans
        (complete_scene_loader.chunk_scenes: the host round trips of scan i+1, i+2 overlap the device work of scan i);
        numpy's RNG is consumed in scan order, so the chunks are those of the sequential reference."""
        held, out = [], None
        for i, (chunks, ev) in enumerate(csl.chunk_scenes((scans[k % S][1][0] for k in range(n_scans)), lookahead=2, background=True)):
            cur.wait_event(ev)
            bp_, bl_, keep = consume(i, chunks)
            held.append(keep)
            if len(held) > 3:       # the forward of scan i-3 has long been enqueued; its inputs may go once it ran
                for pl in pipes:
                    cur.wait_stream(pl.main)
                held.pop(0)
            if out is None:
                out = (bp_, bl_)
        for pl in pipes:
            cur.wait_stream(pl.main)
        return out

    np.random.seed(99 + rank)
    bp, bl = run(1)
    torch.cuda.synchronize(dev)
    # fraction of the scan's points whose coordinates come back bit-exact (the reference's float32 height bound can
    # leave the top-most point outside every un-padded cell, complete_scene_loader.py:34,41 -- reproduced, not fixed)
    ok = float((bp == scans[0][1][0]).all(dim=1).float().mean().item())
    stats["chunks"] = stats["points"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    PASSES = 3    # the S scans three times inside ONE event pair: ~0.5 s, and one host hiccup no longer decides the figure
    e0.record(cur)
    run(S * PASSES)
    e1.record(cur)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / PASSES
    stats["chunks"] /= PASSES
    stats["points"] /= PASSES
    # the chunker alone (device tensors in, device tensors out)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(cur)
    for i in range(S):
        ch = csl.chunk_scene(scans[i][1][0])
        ch.gather(scans[i][1][2])
    c1.record(cur)
    torch.cuda.synchronize(dev)
    out = {"workload": "config 4: whole-scan inference data path, %d synthetic scans per GPU (100-200 k points), chunker "
                       "(pipelined over scans, host planning on a worker thread) + geometry forward over the chunks (B=%d) + map_back; "
                       "timed over 3 passes of the scans" % (S, B),
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


def base_config(B, world):
    """The config keys both arms print (the driver compares them)."""
    return {"workload": WORKLOAD, "batch_per_gpu": B, "npoints": NPOINTS, "feature_channels": 6,
            "parallelism": "scene-sharded x%d, no collective" % world,
            "l2": "inputs larger than L2: one step streams > 500 MB (K / V / output tensors) through the 126 MB L2 and "
                  "every step reads a different batch of scenes (ring of batches)"}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores, all host threads,
    same metric and config as the own arm.  A step = a bounded sample of the B-scene batch (S scenes, sized from a
    probe so that warmup + K steps stay within about two minutes); the K*S scenes of the timed region are one stream
    over the thread pool (one scene per thread), so every thread is busy whatever S is."""
    if rank != 0:
        return 0
    from oracle import cpu, ref
    from pcops_b200 import synth
    cpu.lib()
    use_ref = ref.available_cpu()
    threads = host_threads()
    B, K, W = args.batch, max(1, args.steps), max(0, args.warmup)
    xyz, feat = synth.scannet_batch(0, B, NPOINTS)
    # probe: one scene per thread -> whole-pool rate
    probe_n = min(B, threads)
    rate, _ = cpu_scenes_per_s(xyz[:probe_n], feat[:probe_n], threads, use_ref)
    budget_s = 100.0
    S = int(max(1, min(B, budget_s * rate / (K + W))))
    sel = lambda steps: np_arange_mod(steps * S, B)   # noqa: E731
    if W:
        idx = sel(W)
        cpu_scenes_per_s(xyz[idx], feat[idx], threads, use_ref)
    idx = sel(K)
    value, dt = cpu_scenes_per_s(xyz[idx], feat[idx], threads, use_ref)
    sample = "%d steps x %d of the %d scenes of a batch (%d scenes), %d host threads, one scene per thread, %.1f s" \
        % (K, S, B, K * S, threads, dt)
    cfg = base_config(B, world)
    run_cfg = {"scenes_per_step": S,
               "note": "host-only arm: ball query / group_point / three_nn / three_interpolate = the reference's own CPU "
                       "functions compiled from its sources; FPS, gather, attention contraction = C port of the reference "
                       "CUDA / TF algorithm (the reference has no CPU code for them)"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference" if use_ref else "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "run_config": run_cfg,
    }
    print(json.dumps(line))
    return 0


def np_arange_mod(n, b):
    import numpy as np
    return np.arange(n) % b


# ------------------------------------------------------------------------------------------------ own arm
def run_config5(torch, args, rank, world, dev, sharding, fp32_peak_tops):
    """BASELINE config 5: geometry-op scaling sweep -- FPS, ball query, kNN at N = 16k ... 1M points, npoint 1k ... 16k,
    B = --sweep-batch (64) scenes sharded over the ranks (uniform clouds, radius chosen for ~32 expected neighbours,
    nsample = k = 32).  Every entry is ONE call through the C ABI with pre-allocated outputs and workspaces (no
    allocation inside the timed region), after a cheap warm-up call of the same kernel (2 samples / 32 queries), timed
    with CUDA events on the launching stream; a 512 MB write between entries flushes L2.  Contains NO collective: the
    per-entry times of all ranks are reduced in one call afterwards (max over ranks)."""
    import ctypes
    import pcops_b200 as ops
    L, p, ws_of = ops._lib.lib(), ops._lib.ptr, ops._lib.workspace
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    Bt = args.sweep_batch
    lo, hi = sharding.shard_bounds(Bt, rank, world)
    b = hi - lo
    entries, times = [], []
    flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev).manual_seed(1000 + lo)
    f32, i32 = torch.float32, torch.int32

    def timed(fn):
        flush.fill_(0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn()
        e1.record()
        torch.cuda.synchronize(dev)
        if rc != 0:
            raise RuntimeError("pcops rc %d" % rc)
        return e0.elapsed_time(e1)

    for N in (16384, 65536, 262144, 1048576):
        if N > args.sweep_max_n:
            continue
        x = torch.rand((b, N, 3), generator=g, device=dev) if b > 0 else None
        r = (32.0 / N * 3.0 / (4.0 * 3.14159265)) ** (1.0 / 3.0)
        for m in (1024, 4096, 16384):
            for op in ("fps", "ball", "knn"):
                entries.append((op, N, m))
            if b == 0:
                times += [0.0, 0.0, 0.0]
                continue
            try:
                fi = torch.empty((b, m), dtype=i32, device=dev)
                q = torch.empty((b, m, 3), dtype=f32, device=dev)
                fws = ws_of(L.pc_fps_workspace_bytes(b, N, m), dev)
                idx = torch.empty((b, m, 32), dtype=i32, device=dev)
                cnt = torch.empty((b, m), dtype=i32, device=dev)
                bws = ws_of(L.pc_query_ball_grid_workspace_bytes(b, N, m), dev)
                val = torch.empty((b, m, 32), dtype=f32, device=dev)
                L.pc_fps_gather(b, N, 2, p(x), p(fws), p(fi), p(q), st)              # warm-up (attributes, code load)
                t_fps = timed(lambda: L.pc_fps_gather(b, N, m, p(x), p(fws), p(fi), p(q), st))
            except Exception:
                times += [-1.0, -1.0, -1.0]
                continue
            times.append(t_fps)
            try:
                L.pc_query_ball_grid(b, N, 32, r, 32, p(x), p(q), p(idx), p(cnt), p(bws), st)
                t_ball = timed(lambda: L.pc_query_ball_grid(b, N, m, r, 32, p(x), p(q), p(idx), p(cnt), p(bws), st))
            except Exception:
                t_ball = -1.0
            times.append(t_ball)
            try:
                L.pc_knn(b, N, 32, 32, 3, p(x), p(q), p(val), p(idx), st)
                t_knn = timed(lambda: L.pc_knn(b, N, m, 32, 3, p(x), p(q), p(val), p(idx), st))
            except Exception:
                t_knn = -1.0
            times.append(t_knn)
            del fi, q, fws, idx, cnt, bws, val
        del x
        torch.cuda.empty_cache()
    del flush
    torch.cuda.empty_cache()
    return entries, times, Bt


def finish_config5(entries, times, Bt, world, fp32_peak_tops):
    rows = []
    for (op, N, m), ms in zip(entries, times):
        row = {"op": op, "n": N, "npoint": m, "ms": ms if ms >= 0 else None}
        if ms > 0:
            row["scenes_per_s"] = Bt / (ms * 1e-3)
            pairs = float(Bt) * N * (m - 1 if op == "fps" else m)
            row["gpairs_per_s"] = pairs / (ms * 1e-3) / 1e9          # distance evaluations (kNN makes two passes)
            flops = pairs * (10 if op == "fps" else 11)
            row["frac_fp32"] = flops / (ms * 1e-3) / 1e12 / (fp32_peak_tops * world)
        rows.append(row)
    ok = [r for r in rows if r["ms"]]
    return {"workload": "config 5: FPS / ball query (r for ~32 neighbours, nsample 32) / kNN (k 32) on uniform clouds, "
                        "B=%d scenes sharded over %d GPU(s); one launch each, max over ranks; frac_fp32 = algorithmic "
                        "pair tests x 10 (FPS) or 11 flops / whole-job un-fused fp32 peak" % (Bt, world),
            "paths": "FPS: one 2..16-CTA cluster per scene up to 262144 points (state on chip, DSMEM exchange), beyond that "
                     "a cooperative grid of 16384-point CTAs (state on chip, winners exchanged through global memory); "
                     "ball query: cell grid up to 21088 points (frac_fp32 may exceed 1: pruned pairs), all-pairs kernel "
                     "above (the per-query bitmap over original indices no longer fits shared memory); "
                     "kNN: fused two-pass kernel, no (b,m,n) matrix",
            "rows": rows, "total_ms": sum(r["ms"] for r in ok), "failed": len(rows) - len(ok)}


def time_reference_gpu_kernels(torch, dev, dev_xyz, dev_feat):
    """The reference's own CUDA kernels (tf_sampling_g.cu, tf_grouping_g.cu compiled UNMODIFIED for sm_100a with the
    reference's flags into oracle/_ref/libref_gpu.so; original launch shapes, legacy default stream) timed per op on
    this GPU at the bench shapes: the GPU-vs-GPU baseline of SURVEY.md 8(d).  Bench-leg use of oracle/ only."""
    from oracle import ref
    import pcops_b200 as ops
    if not ref.available_gpu(nofma=False):
        return {"unavailable": "oracle/_ref/libref_gpu.so not present"}
    G = ref.Gpu(nofma=False)
    x, f = dev_xyz, dev_feat
    B, n = x.shape[0], x.shape[1]
    torch.cuda.synchronize(dev)

    def med(fn, iters=5):
        fn()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]
    out = {"source": "tf_sampling_g.cu:203-205 <<<32,512>>>, tf_grouping_g.cu:125-141 <<<b,256>>>, nvcc -O2 (default fmad), "
                     "legacy default stream, median of 5, B=%d" % B}
    with torch.cuda.stream(torch.cuda.default_stream(dev)):
        temp = torch.empty((32, n), dtype=torch.float32, device=dev)
        fi = torch.zeros((B, 1024), dtype=torch.int32, device=dev)
        ref_fps = med(lambda: G.fps_launch(1024, x, temp, fi))
        ours_fps = med(lambda: ops.farthest_point_sample(1024, x))
        nx = ops.gather_point(x, ops.farthest_point_sample(1024, x))
        idx = torch.zeros((B, 1024, 32), dtype=torch.int32, device=dev)
        cnt = torch.zeros((B, 1024), dtype=torch.int32, device=dev)
        ref_ball = med(lambda: G.ball_launch(0.1, 32, x, nx, idx, cnt))
        ours_ball = med(lambda: ops.query_ball_point(0.1, 32, x, nx))
        bi, _ = ops.query_ball_point(0.1, 32, x, nx)
        gx = torch.empty((B, 1024, 32, 3), dtype=torch.float32, device=dev)
        ref_gxyz = med(lambda: G.group_launch(x, bi, gx))
        ours_gxyz = med(lambda: ops.group_point(x, bi))
        p64 = torch.randn((B, 1024, 64), device=dev)
        i2 = torch.randint(0, 1024, (B, 256, 32), dtype=torch.int32, device=dev)
        g64 = torch.empty((B, 256, 32, 64), dtype=torch.float32, device=dev)
        ref_g64 = med(lambda: G.group_launch(p64, i2, g64))
        ours_g64 = med(lambda: ops.group_point(p64, i2))
    out["ops"] = {
        "fps_sa1": {"reference_ms": ref_fps, "ours_ms": ours_fps, "speedup": ref_fps / ours_fps},
        "query_ball_sa1": {"reference_ms": ref_ball, "ours_ms": ours_ball, "speedup": ref_ball / ours_ball},
        "group_point_xyz_sa1": {"reference_ms": ref_gxyz, "ours_ms": ours_gxyz, "speedup": ref_gxyz / ours_gxyz},
        "group_point_c64_sa2": {"reference_ms": ref_g64, "ours_ms": ours_g64, "speedup": ref_g64 / ours_g64},
    }
    return out


def probe_host_links(torch, dev, sharding):
    """Pinned-host <-> device copy bandwidth of this rank while ALL ranks copy at the same time (64 MB each way): the
    ceiling of the e2e region at N GPUs.  Returns (h2d GB/s, d2h GB/s) as the MIN over ranks."""
    n = 64 * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    out = []
    for direction in (0, 1):
        (d.copy_(h, non_blocking=True) if direction == 0 else h.copy_(d, non_blocking=True))
        torch.cuda.synchronize(dev)
        sharding.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            (d.copy_(h, non_blocking=True) if direction == 0 else h.copy_(d, non_blocking=True))
        e1.record()
        torch.cuda.synchronize(dev)
        gbs = 4 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
        out.append(-sharding.max_over_ranks(-gbs))
    del h, d
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)

    import math
    import numpy as np
    import torch
    import torch.distributed as dist

    import pcops_b200  # noqa: F401  (raises if libpcops.so is missing -- there is no fallback)
    if args.lib:
        pcops_b200._lib.LIB_PATH = os.path.abspath(args.lib)
    pcops_b200._lib.lib()
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
    B, R, K, W = args.batch, max(1, args.ring), max(1, args.steps), max(args.warmup, 3)

    # scene shards: rank r owns scenes [r*B*R, (r+1)*B*R) of the synthetic scene list
    lo, _hi = sharding.shard_bounds(world * B * R, rank, world)
    host_xyz, host_feat, host_packed, dev_xyz, dev_feat = [], [], [], [], []
    for i in range(R):
        x, f = synth.scannet_batch(lo + i * B, B, NPOINTS)
        hx, hf = torch.from_numpy(x).pin_memory(), torch.from_numpy(f).pin_memory()
        host_xyz.append(hx)
        host_feat.append(hf)
        col, nrm = synth.split_features(f)      # storage form of the data set: colours are bytes (train.py:95)
        host_packed.append(ScanNetGeometry.pack_host_batch(x, col, nrm))
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

    def fork(ps):
        for pl in ps:
            pl.main.wait_stream(cur)

    def join(ps):
        for pl in ps:
            cur.wait_stream(pl.main)

    rank_spread = []

    def timed_region(step, ps, sampler=None):
        """K steps x `repeats` inside ONE event pair on `cur` (forked to / joined from the pipelines' streams), repeats
        agreed over the ranks so that the region lasts >= --min-seconds.  Returns (ms per K-step block, repeats, host
        enqueue seconds, wall t0, wall t1)."""
        for i in range(max(3, len(ps))):
            step(i)
        torch.cuda.synchronize(dev)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(cur)
        fork(ps)
        for i in range(K):
            step(i)
        join(ps)
        p1.record(cur)
        torch.cuda.synchronize(dev)
        est = sharding.max_over_ranks(p0.elapsed_time(p1))
        reps = max(1, int(math.ceil(args.min_seconds * 1e3 / max(est, 1e-3))))
        if sampler is not None:
            sampler.start()
            time.sleep(0.25)
        sharding.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(cur)
        fork(ps)
        for i in range(reps * K):          # ONE running index: the round robin over the batches in flight must not
            step(i)                        # restart every K steps (K = 20, 8 batches: 4 of 20 steps would queue behind
        join(ps)                           # the replay just issued on the same instance -- 10 % of the N > 1 figures)
        e1.record(cur)
        t_enq = time.time() - t0
        torch.cuda.synchronize(dev)
        t1 = time.time()
        sharding.barrier()
        own = e0.elapsed_time(e1)
        ms = sharding.max_over_ranks(own)
        rank_spread.append(-sharding.max_over_ranks(-own) / ms)       # fastest rank / slowest rank, per region
        return ms / reps, reps, t_enq / reps, t0, t1

    def step_resident(i, probes=None, graph=False):
        pl = pipes[i % D]
        pl.set_inputs(dev_xyz[i % R], dev_feat[i % R])
        if graph:
            pl.replay()
        else:
            pl.forward(overlap, probes)

    # The library sizes its streaming kernels by the number of launches the caller overlaps (pc_set_concurrency_hint):
    # each pipelined region announces its batches in flight BEFORE its graphs are captured (grid sizes are baked into
    # a graph); every lone-launch measurement below (probes, steady-state gathers, configs 1 and 5) runs with 1.
    from pcops_b200 import _lib as pclib

    def hint(n):
        pclib.set_concurrency_hint(args.hint if args.hint > 0 and n > 1 else n)
    hint(D)
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
    sampler = ClockSampler(local) if rank == 0 else None
    ms_block, reps, enq_s, t_wall0, t_wall1 = timed_region(lambda i: step_resident(i, None, use_graph), pipes, sampler)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    value = world * B * K / (ms_block * 1e-3)

    # ---- timed region 2: end to end through host buffers -------------------------------------------------------
    h2d_gbs, d2h_gbs = probe_host_links(torch, dev, sharding)
    if args.e2e_legacy:
        host_out = [torch.empty(pl.result_arena().shape, dtype=torch.int32).pin_memory() for pl in pipes]
        h2d, d2h = pipe.input_bytes(), pipe.result_arena().numel() * 4

        def step_e2e(i):
            pl = pipes[i % D]
            pl.set_inputs(host_xyz[i % R], host_feat[i % R], non_blocking=True)
            if use_graph:
                pl.replay()
            else:
                pl.forward(overlap)
            pl.read_results(host_out[i % D])
        e2e_launches = pipe.launches_per_step
    else:
        host_out = [torch.empty(pl.result_arena().numel(), dtype=torch.int16).pin_memory() for pl in pipes]
        h2d, d2h = pipe.packed_input_bytes(), pipe.result_bytes_u16()
        e2e_graphs = {}
        if use_graph:   # one graph per (pipeline instance, input slot): H2D + prologue + forward + narrowing + D2H
            for d_ in range(D):
                for r_ in sorted({i % R for i in range(d_, D * R, D)}):
                    e2e_graphs[(d_, r_)] = pipes[d_].capture_e2e(host_packed[r_], host_out[d_], overlap)

        def step_e2e(i):
            pl = pipes[i % D]
            if use_graph:
                with torch.cuda.stream(pl.main):
                    e2e_graphs[(i % D, i % R)].replay()
            else:
                pl.set_inputs_packed(host_packed[i % R])
                pl.forward(overlap)
                pl.read_results_u16(host_out[i % D])
        e2e_launches = pipe.launches_per_step + 2
    e2e_block, e2e_reps, e2e_enq, _, _ = timed_region(step_e2e, pipes)
    e2e_value = world * B * K / (e2e_block * 1e-3)
    checksum = int(sum(int(h.view(torch.uint16).to(torch.int64).sum()) if h.dtype == torch.int16
                       else int(h.to(torch.int64).sum()) for h in host_out))
    if use_graph and not args.e2e_legacy:
        e2e_graphs.clear()

    # ---- timed region 3: config 3, a training step's geometry (forward + GroupPointGrad / ThreeInterpolateGrad /
    # attention-contraction backward), inputs resident ----------------------------------------------------------
    train = None
    if args.train and args.attention:
        TD = max(1, min(D, args.train_depth))
        hint(TD)
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
        tr_block, tr_reps, _, _, _ = timed_region(step_train, tp)
        train = {"workload": "config 3: forward + registered gradients (GroupPointGrad SA2-4, ThreeInterpolateGrad FP1-4, "
                             "attention contraction backward SA1-4), B=%d x %d" % (B, NPOINTS),
                 "value": world * B * K / (tr_block * 1e-3), "unit": UNIT, "ms_per_step": tr_block / K, "repeats": tr_reps,
                 "batches_in_flight": TD, "gpu_launches": tp[0].launches_per_train_step * K * tr_reps * world}
        if use_graph:   # back to the forward-only graphs for the regions below
            for pl in tp:
                pl.capture(overlap)

    # ---- timed region 3b: the forward with the WHOLE attention layer per level (Dense Q/K/V on the tensor cores +
    # contraction, from stand-in grouped activations) instead of the contraction alone on stand-in K / V -------------
    with_layers = None
    if args.attention_layers and args.attention:
        LD = max(1, min(D, 4))
        hint(LD)
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
        lay_block, lay_reps, _, _, _ = timed_region(step_layers, lp)
        with_layers = {"workload": "the same forward with pc_attention_layer_fwd (Dense Q/K/V as 3xTF32 tcgen05 UMMA + "
                                   "contraction, K and V never stored) at all four levels; +25.8 GFLOP per step",
                       "value": world * B * K / (lay_block * 1e-3), "unit": UNIT, "ms_per_step": lay_block / K,
                       "repeats": lay_reps, "batches_in_flight": LD,
                       "gpu_launches": lp[0].launches_per_step * K * lay_reps * world}
        del lp
        torch.cuda.empty_cache()

    # ---- timed region 3c: the WHOLE inference forward of the attention model (pointnet2_sem_seg_attention.py:28-62):
    # geometry + shared MLPs + attention layers + FP MLPs + fc, every tensor produced by the layer before it ----------
    full_model = None
    if args.full_model:
        try:
            from pcops_b200.model_pipeline import ScanNetAttentionModel
            MD = max(1, min(D, 4))
            hint(MD)
            models = [ScanNetAttentionModel(B, NPOINTS, 6, dev, seed=rank * 16 + d) for d in range(MD)]
            for md in models:
                md.set_inputs(dev_xyz[0], dev_feat[0])
                if use_graph:
                    md.capture()
                else:
                    md.forward()
            torch.cuda.synchronize(dev)

            def step_model(i):
                md = models[i % MD]
                md.set_inputs(dev_xyz[i % R], dev_feat[i % R])
                if use_graph:
                    md.replay()
                else:
                    md.forward()
            fm_ok, fm_flops, fm_launches = 1.0, models[0].dense_flops(), models[0].launches_per_step
        except Exception as exc:
            fm_ok, full_model = -1.0, {"error": repr(exc)[:300]}
        # the timed region itself contains collectives (agreeing on the repeats): only enter it if every rank is ready
        if sharding.max_vector_over_ranks([fm_ok])[0] > 0:
            fm_block, fm_reps, _, _, _ = timed_region(step_model, models)
            full_model = {"workload": "whole attention-model inference forward, B=%d x %d, xyz + 6 features -> %d-class logits: "
                                      "4 x (sample_and_group, shared MLP, AttentionLayer), 4 x (three_nn, interpolate, MLP), fc1, "
                                      "fc2; no stand-in tensors; dense layers = 3xTF32 tcgen05" % (B, NPOINTS, 21),
                          "value": world * B * K / (fm_block * 1e-3), "unit": UNIT, "ms_per_step": fm_block / K,
                          "repeats": fm_reps, "batches_in_flight": MD, "dense_gflop_per_step": fm_flops / 1e9,
                          "dense_tflops_fp32_equivalent": fm_flops / (fm_block / K * 1e-3) / 1e12,
                          "gpu_launches": fm_launches * K * fm_reps * world}
            del models
            torch.cuda.empty_cache()
        elif full_model is None:
            full_model = {"error": "the whole-model region failed on another rank"}

    # ---- timed region 4: config 4, whole scans through the GPU chunker + forward + map_back -------------------
    config4 = None
    hint(D)
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
                config4["scans_total"] = world * S4
                config4["value"] = world * S4 / (c4_max * 1e-3)
                config4["ms_per_scan"] = c4_max / S4

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

    # ---- region 5: config 5, the geometry-op scaling sweep (no collective inside; one vector reduce after) ------
    hint(1)
    config5 = None
    if args.config5:
        try:
            entries, times, Bt = run_config5(torch, args, rank, world, dev, sharding, fp32_peak_tops)
        except Exception as exc:
            entries, times, Bt = None, None, args.sweep_batch
            config5 = {"error": repr(exc)[:300]}
        n_entries = 3 * 3 * sum(1 for N in (16384, 65536, 262144, 1048576) if N <= args.sweep_max_n)
        vec = sharding.max_vector_over_ranks(times if times is not None and len(times) == n_entries else [-1.0] * n_entries)
        if entries is not None and len(entries) == n_entries:
            config5 = finish_config5(entries, vec, Bt, world, fp32_peak_tops)

    # ---- probed pass: every op's duration (separate region; events perturb overlap) ----------------------------
    def roof(name, ms):
        wk = work[name]
        if wk["kind"] == "bytes":
            ach = wk["amount"] / (ms * 1e-3) / 1e9
            return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": None, "ms": ms}
        ach = wk["amount"] / (ms * 1e-3) / 1e12
        return {"kernel": name, "bound": "fp32", "achieved": ach, "peak": fp32_peak_tops, "unit": "TFLOP/s",
                "frac": ach / fp32_peak_tops, "traffic": None, "ms": ms}

    rooflines, op_ms, grid_ms, fused_layer, steady, config1, ref_gpu = {}, {}, {}, None, None, None, None
    if not args.skip_probe:
        # Every op's stand-alone duration: one pipeline instance alone, eager, ONE stream.  The table is taken with the
        # reference-signature ops (all-pairs ball query / three_nn, separate gather), whose algorithmic op counts the
        # fractions are computed from; the durations of the cell-grid variants the timed pipeline uses are listed
        # beside them in `grid_variants_ms`.
        def probe(pl):
            allp = {n: [] for n in pl.op_names()}
            for i in range(10):
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
            for i in range(10):
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
        try:
            fused_layer = time_fused_attention_layer(torch, pcops_b200, B)
        except Exception as exc:  # reported, never fatal for the headline numbers
            fused_layer = {"error": str(exc)[:200]}
        if rank == 0:
            try:   # GPU-vs-GPU: the reference's own CUDA kernels, recompiled unmodified, on this GPU
                ref_gpu = time_reference_gpu_kernels(torch, dev, dev_xyz[0], dev_feat[0])
            except Exception as exc:
                ref_gpu = {"error": str(exc)[:200]}
    if top_guess in op_ms:
        top_ms = grid_ms.get(top_guess, op_ms[top_guess])
    else:                                 # --skip-probe: time the dominant kernel alone
        solo = {top_guess: []}
        for i in range(5):
            pipe.set_inputs(dev_xyz[i % R], dev_feat[i % R])
            pipe.forward(False, solo)
        torch.cuda.synchronize(dev)
        d = sorted(a.elapsed_time(b) for a, b in solo[top_guess]) or [float("inf")]
        top_ms = d[len(d) // 2]
    roofline = roof(top_guess, top_ms)
    roofline["peak_source"] = "%d SMs x 128 fp32 lanes x %.0f MHz, un-fused (1 flop per lane-clock); no fp32 peak in " \
        "MEASURED_PEAKS.json -- scripts/ubench/fp32_rate.cu measures 127 flop/clk/SM on this part" % (nsm, sm_max)
    roofline["algorithmic_bytes"] = work[top_guess].get("bytes")
    roofline["kernel_sms"] = min(B, nsm)   # one scene per SM: a B-scene launch cannot occupy more
    # ncu evidence of the TIMED kernel (committed capture of this round; static, labelled with its source)
    try:
        ev = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json"))).get(top_guess, {})
        if B == 16 and ev:
            roofline["traffic"] = ev.get("dram_bytes")
            roofline["ncu"] = {k: ev.get(k) for k in ("kernel", "capture", "duration_us", "issue_slots_busy_pct",
                                                      "top_pipe", "top_pipe_pct", "cycles_per_round", "registers")}
    except Exception:
        pass
    # the pipelined step as a whole: where its SM-time and its HBM bytes go
    step_ms = ms_block / K
    step_bytes = sum(w.get("bytes", w["amount"]) if w["kind"] != "bytes" else w["amount"]
                     for n_, w in work.items() if n_ in set(pipe.op_names()))
    fps_names = [n_ for n_ in pipe.op_names() if n_.startswith("fps")]
    src_ms = grid_ms if grid_ms else op_ms
    step_shape = {"ms_per_step": step_ms, "algorithmic_bytes_per_step": step_bytes,
                  "step_hbm_frac": step_bytes / (step_ms * 1e-3) / 1e9 / hbm_peak if step_ms > 0 else None}
    if src_ms:
        fps_sm_ms = sum(src_ms.get(n_, 0.0) for n_ in fps_names) * min(B, nsm)
        step_shape["fps_sm_time_share"] = fps_sm_ms / (nsm * step_ms)
        step_shape["note"] = "fps_sm_time_share = (SMs an FPS launch holds x its duration, 4 levels) / (all SMs x ms_per_step); " \
                             "HBM-byte shares per op = rooflines[op] bytes / algorithmic_bytes_per_step"
        att_bytes = sum(work[n_]["amount"] for n_ in pipe.op_names() if n_.startswith("attention"))
        step_shape["attention_kv_byte_share"] = att_bytes / step_bytes
    roofline["step"] = step_shape
    roofline["timed"] = "CUDA events around each launch on its stream in a probed single-stream eager pass right after " \
                        "the value region (graph replay hides single launches)"

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

    def brief(d, keys):
        return {k: d.get(k) for k in keys} if isinstance(d, dict) else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_block / K, "timed_repeats": reps, "steps_timed": K * reps,
        "fastest_over_slowest_rank": rank_spread[:2],      # [value region, e2e region]: 1.0 = all ranks equally fast
        "host_enqueue_ms_per_step": 1e3 * enq_s / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(B, world),
        "run_config": {"rank_cpu_affinity": numa, "streams_per_batch": 5 if overlap else 1, "cuda_graph": use_graph,
                       "input_ring": R, "batches_in_flight": D, "neighbour_search": "cell grid" if args.grid else "all pairs",
                       "min_seconds": args.min_seconds, "lib": os.path.basename(pcops_b200._lib.LIB_PATH),
                       "concurrency_hint": "pc_set_concurrency_hint(batches in flight) for the pipelined regions%s, 1 for "
                                           "every lone-launch figure (rooflines, gathers_steady_state, configs 1 and 5)"
                                           % (" (forced to %d)" % args.hint if args.hint > 0 else "")},
        "fps_us_per_scene": {"sa1_batch_latency_us": top_ms * 1e3, "sa1_us_per_scene_throughput": top_ms * 1e3 / B},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_block / K, "repeats": e2e_reps, "result_checksum": checksum,
                "host_enqueue_ms_per_step": 1e3 * e2e_enq / K,
                "h2d_gbs": h2d_gbs, "d2h_gbs": d2h_gbs,
                "link_floor_ms_per_step": 1e3 * (h2d / (h2d_gbs * 1e9) + d2h / (d2h_gbs * 1e9)) if not args.e2e_legacy else None,
                "what": ("round-1 path: two fp32 H2D copies, int32 result arena" if args.e2e_legacy else
                         "one CUDA graph per step: H2D of ONE packed arena (xyz | normals | colours uint8 = 27 B/point), "
                         "pc_unpack_features, forward, pc_narrow_indices_u16, ONE D2H of the integer decisions (FPS / ball / "
                         "three_nn indices + counts as uint16; grouped / interpolated tensors stay on the GPU for their "
                         "on-GPU consumer); h2d_gbs / d2h_gbs = pinned-copy bandwidth with all ranks copying at once, "
                         "min over ranks")},
        "gpu_launches": pipe.launches_per_step * K * reps * world,
        "gpu_launches_e2e": e2e_launches * K * e2e_reps * world,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "rooflines": rooflines,
        "grid_variants_ms": grid_ms,
        "gathers_steady_state": steady,
        "attention_layer_tcgen05": fused_layer,
        "reference_gpu_kernels": ref_gpu,
        "config1_single_scene_sa1": config1,
        "with_attention_layers": with_layers,
        "full_model_inference": full_model,
        "config3_training_step": train,
        "config4_whole_scene": config4,
        "config5_sweep": config5,
    }
    # the numbers of every config once more, last in the line, so that they survive a truncated log tail
    line["summary"] = {
        "config2_value": value, "config2_e2e": e2e_value, "n_gpus": world,
        "config3": brief(train, ("value", "ms_per_step")),
        "config4": brief(config4, ("value", "ms_per_scan", "scans_total", "chunker_ms_per_scan", "error")),
        "config5": brief(config5, ("total_ms", "failed", "error")),
        "with_attention_layers": brief(with_layers, ("value",)),
        "full_model_inference": brief(full_model, ("value", "ms_per_step", "dense_tflops_fp32_equivalent", "error")),
        "roofline_frac": roofline["frac"], "step_hbm_frac": step_shape["step_hbm_frac"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
