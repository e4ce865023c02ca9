"""Multi-GPU MSM: shard the (scalar, point) pairs across ranks, ONE collective, then add.

The MSM is a sum of independent terms, so it shards by pairs with no data-path collective until the end
(shape 2): rank r reduces pairs [lo_r, hi_r) to one affine partial sum (96 B), one
all-gather moves world x 96 bytes, and every rank adds the partials (an MSM with unit scalars).
NCCL cannot reduce curve points (its reductions are numeric), hence all-gather + EC add.

The functions are backend-agnostic (torch uint8 tensors + callables), so the same code runs under
NCCL on GPUs (bench.py) and under gloo on CPU in tests/test_multi_gloo.py.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist

UNIT_SCALAR = (1).to_bytes(32, "big")


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of range(n_total): sizes differ by at most one."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def msm_sharded(local_partial: Callable[[], torch.Tensor], combine: Callable[[torch.Tensor, int], torch.Tensor],
                point_bytes: int = 96, group=None) -> torch.Tensor:
    """local_partial() -> uint8[point_bytes] (this rank's partial sum, operand form, on the rank's device);
    combine(gathered uint8[world*point_bytes], world) -> result tensor.  Returns combine's result on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    part = local_partial()
    assert part.dtype == torch.uint8 and part.numel() == point_bytes
    if world == 1:
        return combine(part, 1)
    gathered = torch.empty(world * point_bytes, dtype=torch.uint8, device=part.device)
    dist.all_gather_into_tensor(gathered, part.contiguous(), group=group)
    return combine(gathered, world)
