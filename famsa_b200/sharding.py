"""Multi-GPU plumbing for HP-1: the triangle shards by rows, one process per GPU, and the only exchange step is
one all-gather of the row blocks (NCCL on GPUs, gloo in the CPU tests).  No compute lives here."""
from __future__ import annotations

import math


def tri(r: int) -> int:
    return r * (r - 1) // 2 if r else 0


def row_shards(n: int, parts: int) -> list[int]:
    """Row boundaries b[0]=0 <= ... <= b[parts]=n giving every rank (almost) the same number of pairs
    (row i holds i pairs, so boundaries sit at n*sqrt(r/parts))."""
    total = tri(n)
    bounds = [0]
    for r in range(1, parts):
        target = total * r / parts
        b = int(round((1 + math.sqrt(1 + 8 * target)) / 2))
        bounds.append(min(max(b, bounds[-1]), n))
    bounds.append(n)
    return bounds


def shard_sizes(bounds: list[int]) -> list[int]:
    return [tri(bounds[r + 1]) - tri(bounds[r]) for r in range(len(bounds) - 1)]


def all_gather_blocks(block, bounds, dist, out=None):
    """block: 1-D tensor holding this rank's packed rows [bounds[rank], bounds[rank+1]) padded to the largest
    shard.  Returns the gathered (world * max_shard) tensor; full_triangle() strips the padding."""
    import torch
    world = len(bounds) - 1
    if out is None:
        out = torch.empty(block.numel() * world, dtype=block.dtype, device=block.device)
    # gather raw bytes: every backend (gloo in the CPU tests, NCCL on GPUs) moves uint8
    dist.all_gather_into_tensor(out.view(torch.uint8), block.view(torch.uint8))
    return out


def full_triangle(gathered, bounds):
    """Concatenate the real part of every rank's block -> the packed lower triangle of all n rows."""
    import torch
    sizes = shard_sizes(bounds)
    m = max(sizes)
    return torch.cat([gathered[r * m:r * m + sizes[r]] for r in range(len(sizes))])
