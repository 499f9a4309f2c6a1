"""Multi-GPU plumbing: chains shard across ranks with no data-path
collective; the Philox key of a chain is its GLOBAL id, so a rank only needs its
(offset, count).  One all-gather of draws at the end (NCCL on GPUs, gloo in CPU tests)."""
from typing import Tuple


def shard(total_chains: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition: rank r owns global chains [offset, offset + count)."""
    if not (0 <= rank < world_size):
        raise ValueError("0 <= rank < world_size")
    base, rem = divmod(total_chains, world_size)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def gather_draws(local, total_chains: int, group=None):
    """All-gather per-chain rows ([count, ...] on every rank) into [total_chains, ...]
    in global chain order.  `local` is a torch tensor (CUDA for NCCL, CPU for gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = [shard(total_chains, world, r)[1] for r in range(world)]
    if len(set(counts)) == 1:
        out = torch.empty((total_chains,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # uneven shards: pad every rank to the largest block, gather, drop the padding
    mx = max(counts)
    padded = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * mx: r * mx + c] for r, c in enumerate(counts)], dim=0)
