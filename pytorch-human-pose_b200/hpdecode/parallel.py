"""Multi-GPU plumbing: images shard by contiguous ranges, one process per GPU, and the only
communication is a gather of the fixed-size packed pose lists to rank 0 (NCCL on CUDA tensors;
the same code runs over gloo on CPU tensors in the tests).  No collective touches the hot path
(SURVEY.md 8(e))."""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of rank's share of n_items; the first n_items % world ranks get one extra."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_packed(packed: torch.Tensor, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Gather every rank's packed result rows [b_r, F] on ``dst`` (ranks may hold different b_r).
    Returns the concatenation in rank order on dst, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return packed
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [torch.zeros(1, dtype=torch.int64, device=packed.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([packed.shape[0]], dtype=torch.int64, device=packed.device), group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(counts)
    padded = packed
    if packed.shape[0] < cap:
        padded = torch.cat([packed, packed.new_zeros((cap - packed.shape[0], packed.shape[1]))])
    if rank == dst:
        bufs = [torch.empty_like(padded) for _ in range(world)]
        dist.gather(padded.contiguous(), bufs, dst=dst, group=group)
        return torch.cat([b[:c] for b, c in zip(bufs, counts)])
    dist.gather(padded.contiguous(), None, dst=dst, group=group)
    return None


def gather_packed_equal(packed: torch.Tensor, out: Optional[torch.Tensor], dst: int = 0, group=None):
    """Fast path for equal shards (the bench): one gather into a preallocated [world*b, F] tensor."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return packed
    if dist.get_rank(group) == dst:
        dist.gather(packed, list(out.chunk(dist.get_world_size(group))), dst=dst, group=group)
        return out
    dist.gather(packed, None, dst=dst, group=group)
    return None


def gather_rows(block: torch.Tensor, out: Optional[torch.Tensor], dst: int = 0, group=None):
    """The bench's gather: every rank contributes one flat tensor of n elements (the result records of one or
    several batches), ``out`` [world, n] on ``dst`` receives rank r's block in row r.  One collective, issued on
    the CURRENT stream (the caller puts it on a communication stream of its own)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return block
    if dist.get_rank(group) == dst:
        dist.gather(block, list(out.unbind(0)), dst=dst, group=group)
        return out
    dist.gather(block, None, dst=dst, group=group)
    return None
