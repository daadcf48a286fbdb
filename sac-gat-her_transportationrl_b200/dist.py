"""Multi-GPU plumbing: one process per GPU, scenarios sharded with no data-path collective.

The assignment path has no arithmetic that couples two scenarios, so each rank
owns a contiguous block of scenarios and solves it independently.  Collectives exist only for
the training loop around the env (BASELINE.json north_star): all-gather of per-step transitions
into every rank's replay buffer, and all-reduce of learner gradients.  They replace the
reference's pickled ``mp.Queue`` traffic (reference: src/train.py:197-209, :863-874).  Backend is
NCCL on GPUs; the same code runs on gloo for the CPU tests.
"""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import torch
import torch.distributed as dist


def scenario_shard(num_scenarios: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of rank `rank`; blocks differ in size by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(num_scenarios, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def all_gather_transitions(local: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Concatenate every rank's per-step transition tensors along dim 0, in rank order.
    All ranks must pass the same keys with equal per-rank shapes (equal shards)."""
    world = _world()
    if world == 1:
        return dict(local)
    out = {}
    for key in sorted(local):
        t = local[key].contiguous()
        buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t)
        out[key] = buf
    return out


def allreduce_mean_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20) -> None:
    """Average gradients across ranks.  Gradients are packed into flat buckets sized for launch
    latency (NVSwitch gives every GPU full bandwidth to every peer, so bucket size is about the
    number of launches, not about links)."""
    world = _world()
    if world == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0

    for g in grads:
        if bucket and (bucket[0].dtype != g.dtype or size + g.numel() * g.element_size() > bucket_bytes):
            flush()
        bucket.append(g)
        size += g.numel() * g.element_size()
    flush()
