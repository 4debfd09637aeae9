"""Batch sharding of ConvBias across the GPUs of one node.

ConvBiasForward is independent per image (the oracle's outer loop is `n`,
dnn/src/naive/convolution/helper.h:467; megdnn's CUDA algos must accept a non-contiguous batch
stride for the same reason, dnn/src/cuda/conv_bias/algo.h:32-37), so GPU g owns images
[g*N/G, (g+1)*N/G) of src / z / dst while filter, bias and scales are replicated.  There is no
collective on the data path; torch.distributed is used only to gather outputs or checksums for
validation and for the timing barrier.
"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """(first image, image count) of `rank`; the first n % world ranks get one extra image."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad world/rank %d/%d" % (world, rank))
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    """view of the rank's images; dim 0 is the batch in every supported format (NHWC, NCHW, NCHW4)."""
    s, c = shard_range(t.shape[0], world, rank)
    return t.narrow(0, s, c)


def gather_batch(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """validation only: all-gather the per-rank dst shards back into the full batch (every rank
    receives it).  Ragged shards are padded to the largest shard for the collective."""
    world = dist.get_world_size(group)
    counts = [shard_range(n_total, world, r)[1] for r in range(world)]
    cmax = max(counts)
    pad = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def checksum(t: torch.Tensor) -> int:
    """order-sensitive 64-bit checksum of a tensor's raw bytes (sum of byte * (index % 65521 + 1))."""
    b = t.contiguous().view(torch.uint8).reshape(-1).to(torch.int64)
    w = (torch.arange(b.numel(), device=b.device, dtype=torch.int64) % 65521) + 1
    return int((b * w).sum().item())


def gather_checksums(local: torch.Tensor, group=None) -> List[int]:
    """validation only: per-rank checksums of the local dst shard, gathered on every rank."""
    world = dist.get_world_size(group)
    mine = torch.tensor([checksum(local)], dtype=torch.int64, device=local.device)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [int(o.item()) for o in out]
