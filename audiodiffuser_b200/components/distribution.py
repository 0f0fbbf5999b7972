"""Training sigma distributions. Mirrors `src/models/components/distribution.py:5-16`."""
import torch
from torch import Tensor


class Distribution:
    def __call__(self, num_samples: int, device: torch.device):
        raise NotImplementedError()


class LogNormalDistribution(Distribution):
    """sigma = exp(mean + std * N(0,1))  (distribution.py:14-16; EDM training sigma)."""

    def __init__(self, mean: float, std: float):
        self.mean = mean
        self.std = std

    def __call__(self, num_samples: int, device: torch.device = torch.device("cpu")) -> Tensor:
        return (self.mean + self.std * torch.randn((num_samples,), device=device)).exp()
