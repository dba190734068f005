"""Sharding of independent planning problems over the GPUs of one box.

Problems (worlds x replan states) are independent, so the data path needs no collective: rank r owns the
contiguous block shard_bounds(n, world, r) and runs it on its own device through its own C-ABI context.
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only after the timed region: to gather the
per-problem results (k_opt, verdict, first violated row) and to reduce timings (max over ranks).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of problem indices owned by `rank`; block sizes differ by at most one."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_results(local: dict[str, torch.Tensor], n_total: int) -> dict[str, torch.Tensor]:
    """All-gather per-problem result tensors (first dim = local problems) into global problem order.
    Works for uneven shards (pads to the largest shard).  Without an initialised process group returns `local`."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    out = {}
    for name, t in local.items():
        pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        out[name] = torch.cat([parts[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
    return out


def reduce_max(value: float, device="cpu") -> float:
    """Max over ranks of a scalar (device-side timings are reported as the max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
