"""Small helpers with the reference's semantics (`src/models/components/utils.py:9-52`)."""
from typing import Optional

import torch
from torch import Tensor


def exists(val) -> bool:
    return val is not None


def extend_dim(x: Tensor, dim: int) -> Tensor:
    """[b] -> [b, 1, ..., 1] with `dim` dims (utils.py:16-18)."""
    return x.view(*x.shape + (1,) * (dim - x.ndim))


def dynamic_threshold_clip(x: Tensor, dynamic_threshold: float) -> Tensor:
    """Per-sample quantile thresholding (utils.py:24-33). Only used when dynamic_threshold != 0;
    the default clamp(-1, 1) (utils.py:21-22) is fused into the CUDA kernels instead."""
    flat = x.reshape(x.shape[0], -1)
    scale = torch.quantile(flat.abs(), dynamic_threshold, dim=-1).clamp_(min=1.0)
    scale = scale.view(-1, *((1,) * (x.ndim - 1)))
    return x.clamp(-scale, scale) / scale


def to_batch(batch_size: int, device: torch.device, x: Optional[float] = None, xs: Optional[Tensor] = None):
    """Reference semantics (utils.py:41-52): exactly one of x / xs. Returns (sigma_tensor, stride):
    a 1-element device tensor with stride 0 for a scalar sigma (no torch.full, no host sync), or the
    [B] tensor with stride 1."""
    assert exists(x) ^ exists(xs), "Either x or xs must be provided"
    if exists(x):
        if isinstance(x, Tensor):
            s = x.detach().reshape(1).to(device=device, dtype=torch.float32, non_blocking=True)
        else:
            s = torch.tensor([float(x)], dtype=torch.float32, device=device)
        return s, 0
    xs = xs.detach().to(device=device, dtype=torch.float32).contiguous()
    assert xs.numel() == batch_size, f"sigmas has {xs.numel()} values for batch {batch_size}"
    return xs.reshape(batch_size), 1
