"""Multi-GPU plumbing of the evolve step: block partition of the evolved points and the one collective of the
path, the all-gather of the evolved point sets (BASELINE.json north_star).  torch.distributed is
plumbing only: NCCL over NVLink on the GPU box, gloo in the CPU tests."""
from __future__ import annotations


def partition(total: int, rank: int, world: int):
    """Contiguous block [lo, hi) of `total` points owned by `rank`; blocks differ by at most one point."""
    assert 0 <= rank < world
    return total * rank // world, total * (rank + 1) // world


def counts(total: int, world: int):
    return [partition(total, r, world)[1] - partition(total, r, world)[0] for r in range(world)]


def all_gather_points(local, total: int, group=None):
    """Gather the per-rank blocks of evolved points ((n_r, 4) tensors) into the full (total, 4) tensor on every
    rank, in rank order (= the original point order).  Uneven blocks are padded to the largest block for the collective."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = torch.empty((total, local.shape[1]), dtype=local.dtype, device=local.device)
    cs = counts(total, world)
    if len(set(cs)) == 1:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # uneven blocks: pad every block to the largest one, gather, then drop the padding rows
    mx = max(cs)
    padded = torch.zeros((mx, local.shape[1]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty((world * mx, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    off = 0
    for r, c in enumerate(cs):
        out[off:off + c] = buf[r * mx:r * mx + c]
        off += c
    return out
