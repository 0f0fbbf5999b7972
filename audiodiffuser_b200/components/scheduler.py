"""Sigma schedules (host side). Mirrors `src/models/components/scheduler.py:6-22` of the reference."""
import torch
import torch.nn as nn
from torch import Tensor


class KarrasSchedule(nn.Module):
    """EDM eq. 5: sigma_i = (smax^(1/rho) + i/(N-1) (smin^(1/rho) - smax^(1/rho)))^rho.

    Same constructor and `forward() -> Tensor[N]` (fp32, CPU) as the reference
    (scheduler.py:9-22); the LightningModule evaluates it once at construction
    (diffunet_complex_module.py:64), so this is never on the device hot path.
    """

    def __init__(self, sigma_min: float, sigma_max: float, rho: float = 7.0, num_steps: int = 50):
        super().__init__()
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max
        self.rho = rho
        self.num_steps = num_steps

    def forward(self) -> Tensor:
        inv = 1.0 / self.rho
        i = torch.arange(self.num_steps, dtype=torch.float32)
        lo, hi = self.sigma_min ** inv, self.sigma_max ** inv
        return (hi + i / (self.num_steps - 1) * (lo - hi)) ** self.rho
