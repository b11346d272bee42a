"""Data-parallel plumbing for the lift-splat path.

The path is per-sample (the reference loops ``for b in range(batch)``,
model/bev_model.py:79), so multi-GPU means: one process per GPU, every rank owns a
contiguous slice of the batch, no collective inside the path.  The only cross-rank
traffic is what the caller adds around it (DDP's gradient all-reduce in training, a
MAX-reduce of device timings in bench.py).  Backend-agnostic: ``nccl`` on GPUs,
``gloo`` in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple

import torch


def shard_bounds(global_batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the samples rank ``rank`` owns; remainders go to the lowest ranks."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(global_batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensor: torch.Tensor, cams: int, world_size: int, rank: int, per_camera: bool) -> torch.Tensor:
    """Slice of a batch-major tensor for one rank.  ``per_camera`` tensors are laid out
    [B*N, ...] (encoder outputs, model/cam_encoder.py:102-111), the others [B, ...]."""
    rows = tensor.shape[0] // cams if per_camera else tensor.shape[0]
    lo, hi = shard_bounds(rows, world_size, rank)
    return tensor[lo * cams:hi * cams] if per_camera else tensor[lo:hi]


def max_over_ranks(value: float, device="cpu") -> float:
    """MAX-reduce a timing over all ranks (the contract of bench.py); identity without a
    process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_batch(local: torch.Tensor, global_rows: int, cams: int, per_camera: bool) -> torch.Tensor:
    """All-gather the per-rank slices back into batch order (used by tests / evaluation)."""
    import torch.distributed as dist
    world = dist.get_world_size()
    mult = cams if per_camera else 1
    sizes = [(shard_bounds(global_rows, world, r)[1] - shard_bounds(global_rows, world, r)[0]) * mult
             for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)
