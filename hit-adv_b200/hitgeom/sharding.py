"""Instance-sharded multi-GPU driver.

Every cloud is attacked independently (per-sample weights / binary-search state in CW/*.py and
ShapeAttack/HiT_ADV.py), so the batch is split by instance across ranks -- one process per GPU, launched by
torchrun -- and NOTHING crosses NVLink inside an attack iteration.  The only collectives are, once at the end,
an `all_gather` of the adversarial clouds and an `all_reduce(SUM)` of the metric counters that
util/other_utils.py:33-43,87-98 accumulates (ASR numerator/denominator, kNN / uniform / curvature sums).
The reference has no distributed code at all; this is the B200-native addition.

The helpers are backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def world():
    return dist.get_world_size() if dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


def shard_range(n_items, r=None, w=None):
    """Contiguous block of instance indices owned by rank r: sizes differ by at most one."""
    r = rank() if r is None else r
    w = world() if w is None else w
    base, rem = divmod(n_items, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def shard(t, r=None, w=None):
    lo, hi = shard_range(t.shape[0], r, w)
    return t[lo:hi]


def gather_clouds(local, n_total, timing=None):
    """all_gather of per-rank result blocks [n_r, ...] -> [n_total, ...] on every rank (ragged tail padded).

    `timing` (optional dict) receives the collective's record: op, payload bytes, and -- for CUDA tensors -- a pair of
    CUDA events bracketing it on the current stream (`timing["events"]`; read them after a synchronize)."""
    w = world()
    if timing is not None:
        timing.update(op="all_gather_into_tensor", world=w, bytes_per_rank=int(local.numel() * local.element_size()),
                      bytes_total=int(local.numel() * local.element_size()) * w, events=None)
    if w == 1:
        return local
    ev = None
    if timing is not None and local.is_cuda:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    per = (n_total + w - 1) // w
    if n_total % w == 0 and local.shape[0] == per and local.is_contiguous():  # equal blocks: no padding, no re-cut
        out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local)
    else:
        pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        full = torch.empty((w * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, pad)
        pieces = []
        for r in range(w):
            lo, hi = shard_range(n_total, r, w)
            pieces.append(full[r * per : r * per + (hi - lo)])
        out = torch.cat(pieces, dim=0)
    if ev is not None:
        ev[1].record()
        timing["events"] = ev
    return out


def reduce_counters(counters, device=None):
    """all_reduce(SUM) of a {name: number} dict (the eval_ASR accumulators); returns python floats."""
    names = sorted(counters)
    if world() == 1:
        return {k: float(counters[k]) for k in names}
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = torch.tensor([float(counters[k]) for k in names], dtype=torch.float64, device=device)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(names, buf.tolist())}


def max_over_ranks(value, device=None):
    """Max of a per-rank scalar (device-time in ms for the benchmark)."""
    if world() == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized():
        dist.barrier()


def run_sharded(fn, data, *extra, timing=None):
    """Run `fn(local_data, *local_extra) -> (local_result [n_r,...], counters dict)` on this rank's block of
    instances and return (all results [n,...], summed counters) on every rank.  `timing`: see gather_clouds."""
    n = data.shape[0]
    lo, hi = shard_range(n)
    result, counters = fn(data[lo:hi], *[e[lo:hi] for e in extra])
    return (gather_clouds(result, n, timing=timing),
            reduce_counters(counters, device=result.device if result.is_cuda else None))
