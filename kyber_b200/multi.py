"""Multi-GPU MSM: shard the (scalar, point) pairs across ranks; two exchange shapes.

Shape 2 (`msm_sharded`, default of bench.py): every rank finishes its own MSM, ONE all-gather of 96-byte results, add.
Shape 1 (`msm_bucket_exchange`, the shape BASELINE.json's north_star names): the ranks exchange their PARTIAL
BUCKETS.  NCCL cannot reduce curve points, so the "all-reduce of partial bucket sums" is spelled out as its two
halves with the reduction done by our own kernel: an all-to-all of raw limb buffers (rank g receives windows
[g W/G, (g+1) W/G) of every rank: a reduce-scatter without the reduction), ONE kernel that reads the receive buffer
and fuses the bucket-wise EC additions into the running-sum reduction of those windows, then an all-gather of the
W window sums and the Horner recombination on every rank.  The bucket reduction is thereby split G ways.


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


def msm_bucket_exchange(local_buckets: Callable[[], torch.Tensor],
                        reduce_windows: Callable[[torch.Tensor, int, int], torch.Tensor],
                        finish: Callable[[torch.Tensor], torch.Tensor],
                        W: int, buckets_per_window: int, bucket_bytes: int, group=None) -> torch.Tensor:
    """Shape 1.  local_buckets() -> uint8[W * buckets_per_window * bucket_bytes], window-major partial buckets of
    this rank; reduce_windows(recv uint8[parts][w_cnt][buckets_per_window][bucket_bytes], parts, w_cnt) ->
    uint8[w_cnt * bucket_bytes] window sums of the windows this rank owns; finish(uint8[W * bucket_bytes]) -> result.
    Rank g owns windows [g * W/G, (g+1) * W/G); W must be a multiple of the world size."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if W % world:
        raise ValueError(f"bucket exchange needs the window count ({W}) to be a multiple of the world size ({world})")
    w_cnt = W // world
    buckets = local_buckets()
    assert buckets.dtype == torch.uint8 and buckets.numel() == W * buckets_per_window * bucket_bytes
    if world == 1:
        return finish(reduce_windows(buckets, 1, w_cnt))
    recv = torch.empty_like(buckets)
    dist.all_to_all_single(recv, buckets, group=group)      # equal splits: chunk j of `buckets` (windows of rank j) -> rank j
    mine = reduce_windows(recv, world, w_cnt)
    assert mine.dtype == torch.uint8 and mine.numel() == w_cnt * bucket_bytes
    wsums = torch.empty(W * bucket_bytes, dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(wsums, mine.contiguous(), group=group)
    return finish(wsums)
