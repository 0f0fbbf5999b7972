"""Data-parallel training step on the fused path (BASELINE.json configs[2]).

The reference trains with Lightning DDP + torch.optim.AdamW (configs/trainer/ddp.yaml:9,
configs/model/diffunet_complex.yaml:7-12, loss = diffusion(x, net, sigmas).mean(),
src/models/diffunet_complex_module.py:120-125). Here one step is

    loss[B]  = EluDiffusion.forward(x, net, sigmas)          fused forward, activations kept (adb_wavenet_dsm_forward_train)
    backward                                                  CUDA backward -> ONE flat fp32 gradient (adb_wavenet_dsm_backward)
    all-reduce(sum) of the flat gradient over NCCL            one collective per step (world > 1 only)
    AdamW on the flat parameter vector, grad / world          adb_adamw_step

The module's parameters are re-pointed at views of one flat fp32 vector (state_dict order), so checkpoints,
`load_state_dict` and the sampling path keep working while the optimizer touches a single buffer.
"""
import math
from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from . import _native as N


def allreduce_sum_(flat: Tensor, world: int) -> Tensor:
    """In-place sum over the default process group (the only exchange step of data-parallel training)."""
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


class FusedTrainer:
    def __init__(self, net, diffusion, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01):
        self.net, self.diffusion = net, diffusion
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        params = list(net.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise N.AdbError("FusedTrainer needs the network on a B200 (no CPU path)")
        self.flat = torch.cat([p.detach().reshape(-1).to(torch.float32) for p in params]).contiguous()
        off = 0
        for p in params:                                   # parameters become views of the flat vector
            n = p.numel()
            p.data = self.flat[off:off + n].view(p.shape)
            off += n
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def step(self, x: Tensor, sigmas: Tensor, noise: Optional[Tensor] = None) -> Tensor:
        """One optimisation step on this rank's micro-batch; returns the per-sample losses [B] (detached)."""
        net = self.net
        for p in net.parameters():
            p.grad = None
        loss = self.diffusion(x, net, sigmas=sigmas, noise=noise)
        loss.mean().backward()
        grad = net._last_flat_grad                         # the flat gradient the backward kernel wrote (state_dict order)
        if grad.numel() != self.flat.numel():
            raise N.AdbError("flat gradient size does not match the parameter vector")
        allreduce_sum_(grad, self.world)
        self.step_count += 1
        N.check(N.lib().adb_adamw_step(N.ptr(self.flat), N.ptr(grad), N.ptr(self.exp_avg), N.ptr(self.exp_avg_sq),
                                       self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                       self.step_count, 1.0 / self.world, N.stream_ptr(self.flat.device)))
        net.parameters_updated()
        return loss.detach()
