"""Scene sharding across the GPUs of one box.

Every op on the path is independent per batch element (reference kernels start from ``batch_index = blockIdx.x``,
tf_grouping_g.cu:4,41,62,84; tf_sampling_g.cu:113,173,184) and every 8192-point chunk of a scene is independent
(complete_scene_loader.py:57-109), so the multi-GPU story is a static partition of the scene / chunk list with NO
data-path collective: one process per GPU, each runs the same kernels on its slice and writes disjoint outputs.
torch.distributed is used only to agree on timings (max over ranks) and totals (sum over ranks).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_units, rank, world):
    """Contiguous balanced split of ``n_units`` into ``world`` parts; returns [lo, hi) of ``rank``.
    The first ``n_units % world`` ranks get one extra unit."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    base, extra = divmod(int(n_units), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_units, world):
    return [shard_bounds(n_units, r, world)[1] - shard_bounds(n_units, r, world)[0] for r in range(world)]


def _reduce(value, op, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(value, device=None):
    return _reduce(value, dist.ReduceOp.MAX, device)


def sum_over_ranks(value, device=None):
    return _reduce(value, dist.ReduceOp.SUM, device)


def max_vector_over_ranks(values, device=None):
    """Element-wise max over ranks of a fixed-length list of floats: ONE collective for a whole optional region (each
    rank passes -1 where its own measurement failed; see agree_on_region for why the region itself holds none)."""
    vals = [float(v) for v in values]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or not vals:
        return vals
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    hi = torch.tensor(vals, dtype=torch.float64, device=device)
    lo = -hi.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MAX)   # -min: an entry that failed anywhere (-1) is reported as failed
    lo = -lo
    return [float(h) if float(l) >= 0 else -1.0 for h, l in zip(hi.tolist(), lo.tolist())]


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def agree_on_region(local_ms):
    """Outcome of an OPTIONAL timed region that a rank may have failed on its own (it caught an exception and passes
    ``None`` or a negative time): returns (max over ranks of the times, True if every rank succeeded).  Every rank must
    call this exactly once per region, success or not -- the region itself must contain no collective, otherwise a
    rank-local failure leaves the others waiting in it (this happened: one rank's scan was rejected by the chunker and
    seven ranks sat in a barrier until the NCCL watchdog fired)."""
    ms = -1.0 if local_ms is None or local_ms < 0 else float(local_ms)
    slowest = max_over_ranks(ms)
    fastest = -max_over_ranks(-ms)
    return slowest, fastest >= 0
