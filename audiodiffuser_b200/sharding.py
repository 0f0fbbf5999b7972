"""Batch-sharded sampling across the GPUs of one box.

Each sample's trajectory depends only on its own noise, so the global batch is split into
contiguous shards, one process per GPU, with NO collective on the sampling path. Sample `g` of the
global batch always uses the generator seed `base_seed + g`, so the waveforms are invariant to the
world size. (The reference instead seeds every DDP rank identically — src/train.py:50-51 — and
all ranks generate duplicates, diffunet_complex_module.py:230-266.)
"""
from typing import Tuple

import torch


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the contiguous shard owned by `rank`; shards differ by at most one sample."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def noise_for_indices(indices, length: int, base_seed: int = 0, channels: int = 1) -> torch.Tensor:
    """N(0,1) noise [len(indices), channels, length] (CPU, fp32, pinned when CUDA is present); the row of global sample
    `g` is drawn from torch.Generator().manual_seed(base_seed + g), whatever rank or batch it lands in."""
    indices = list(indices)
    out = torch.empty(len(indices), channels, length, dtype=torch.float32, pin_memory=torch.cuda.is_available())
    g = torch.Generator()
    for i, gi in enumerate(indices):
        g.manual_seed(base_seed + gi)
        out[i] = torch.randn(channels, length, generator=g)
    return out


def shard_noise(global_batch: int, rank: int, world: int, length: int, base_seed: int = 0, channels: int = 1) -> torch.Tensor:
    """Noise for this rank's contiguous shard of the global batch (see noise_for_indices)."""
    start, stop = shard_range(global_batch, rank, world)
    return noise_for_indices(range(start, stop), length, base_seed, channels)


def gather_shards(local: torch.Tensor, global_batch: int, rank: int, world: int):
    """Optional end-of-run collection (NOT on the sampling path): all-gather variable-size shards and
    return the global batch in order on every rank. Uses the default process group."""
    import torch.distributed as dist
    if world == 1:
        return local
    sizes = [shard_range(global_batch, r, world) for r in range(world)]
    maxn = max(b - a for a, b in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([bufs[r][: b - a] for r, (a, b) in enumerate(sizes)], dim=0)
