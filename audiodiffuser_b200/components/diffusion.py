"""EDM-preconditioned denoiser and DSM loss behind the reference's class API.

Mirrors `src/models/components/diffusion.py`: `Diffusion` base (:15-97) and `EluDiffusion`
(:220-258). Arithmetic runs in the fused sm_100a kernels of libadb200 (adb_edm_*); when `net` is the
fused `audiodiffuser_b200.backbones.wavenet.WaveNetNoise` the preconditioning, backbone and eq. 7
combine run as one device-side sequence (adb_wavenet_denoise). Any other `net` (e.g. a reference
backbone) is still accepted: only the arithmetic around the call is fused.
"""
from typing import Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from .. import _native as N
from .utils import dynamic_threshold_clip, extend_dim, to_batch


class Diffusion(nn.Module):
    """Base class with the reference's `denoise_fn` / `forward` contract (diffusion.py:15-97)."""

    def __init__(self, dynamic_threshold: float = 0.0):
        super().__init__()
        self.dynamic_threshold = dynamic_threshold

    def loss_weight(self, sigmas: Tensor) -> Tensor:
        raise NotImplementedError

    def get_scale_weights(self, sigmas: Tensor, ex_dim: int):
        raise NotImplementedError


class EluDiffusion(Diffusion):
    """Elucidated diffusion (EDM) preconditioning — diffusion.py:220-258.

    Same constructor: `EluDiffusion(sigma_data, dynamic_threshold=0.0)`.
    """

    def __init__(self, sigma_data: float, dynamic_threshold: float = 0.0):
        super().__init__(dynamic_threshold)
        self.sigma_data = sigma_data

    # -- API parity helpers (tiny [B] tensors; not on the hot path) --------------------------------
    def get_scale_weights(self, sigmas: Tensor, ex_dim: int) -> Tuple[Tensor, ...]:
        """(c_skip, c_out, c_in)[B,1..], c_noise[B] — diffusion.py:232-241."""
        sd = self.sigma_data
        c_noise = torch.log(sigmas) * 0.25
        s = extend_dim(sigmas, dim=ex_dim)
        c_skip = (sd ** 2) / (s ** 2 + sd ** 2)
        c_out = s * sd * (sd ** 2 + s ** 2) ** -0.5
        c_in = (s ** 2 + sd ** 2) ** -0.5
        return c_skip, c_out, c_in, c_noise

    def loss_weight(self, sigmas: Tensor) -> Tensor:
        """diffusion.py:243-245."""
        return (sigmas ** 2 + self.sigma_data ** 2) * (sigmas * self.sigma_data) ** -2

    # -- the hot path ------------------------------------------------------------------------------
    def denoise_fn(self, x_noisy: Tensor, net: nn.Module = None, inference: bool = False, cond_scale: float = 1.0,
                   sigmas: Optional[Tensor] = None, sigma: Optional[float] = None, **kwargs) -> Tensor:
        """x0_hat = clip(c_skip x + c_out net(c_in x, c_noise)) — diffusion.py:32-63.

        Exactly one of `sigma` (python float or 0-dim tensor, what every sampler passes) and
        `sigmas` ([B], training) — utils.py:47.
        """
        x = N.require_cuda_f32(x_noisy, "x_noisy")
        N.ensure_device(x.device)
        B = x.shape[0]
        n_per = x[0].numel()
        sig, stride = to_batch(B, x.device, x=sigma, xs=sigmas)
        lib = N.lib()
        st = N.stream_ptr(x.device)

        fused = getattr(net, "_adb_fused_denoise", None)
        if fused is not None and self.dynamic_threshold == 0.0 and not torch.is_grad_enabled():
            # unconditional fused backbone: the CFG branch (diffusion.py:52-54) evaluates the same
            # function twice and lerps identical values, so one evaluation is exact.
            return fused(x, sig, stride, float(self.sigma_data))

        net_in = torch.empty_like(x)
        c_noise = torch.empty(B, dtype=torch.float32, device=x.device)
        N.check(lib.adb_edm_precond_in(N.ptr(x), N.ptr(sig), stride, float(self.sigma_data), N.ptr(net_in),
                                       N.ptr(c_noise), B, n_per, st))
        f_null = None
        pair = getattr(net, "_adb_cfg_pair", None)
        if inference and cond_scale != 1.0 and pair is not None and set(kwargs) == {"classes"} and kwargs["classes"] is not None:
            # both guidance branches in one batch-2B evaluation of the fused backbone (diffusion.py:50-53 calls it twice)
            pred, f_null = pair(net_in, c_noise, kwargs["classes"])
        elif inference:
            pred = net(net_in, c_noise, cond_drop_prob=0., **kwargs)
            if cond_scale != 1.0:
                f_null = net(net_in, c_noise, cond_drop_prob=1., **kwargs)
        else:
            pred = net(net_in, c_noise, **kwargs)
        if pred.shape != x.shape:
            raise N.AdbError(f"net returned {tuple(pred.shape)}, expected {tuple(x.shape)} "
                             "(the reference would silently broadcast here, SURVEY.md §0)")
        if torch.is_grad_enabled() and pred.requires_grad:
            raise NotImplementedError(
                "denoise_fn under autograd with a non-fused net: use EluDiffusion.forward (the fused DSM loss) "
                "or wrap the call in torch.no_grad()")
        pred = N.require_cuda_f32(pred, "net output")
        if self.dynamic_threshold != 0.0:
            c_skip, c_out, _, _ = self.get_scale_weights(sig.expand(B) if stride == 0 else sig, x.ndim)
            if f_null is not None:
                pred = f_null + (pred - f_null) * cond_scale
            return dynamic_threshold_clip(c_skip * x + c_out * pred, self.dynamic_threshold)
        out = torch.empty_like(x)
        fn_ptr = N.ptr(N.require_cuda_f32(f_null, "null-cond net output")) if f_null is not None else N.ptr(None)
        N.check(lib.adb_edm_precond_out(N.ptr(x), N.ptr(pred), fn_ptr, float(cond_scale), N.ptr(sig), stride,
                                        float(self.sigma_data), N.ptr(out), B, n_per, st))
        return out

    def forward(self, x: Tensor, net: nn.Module, sigmas: Tensor, inference: bool = False, cond_scale: float = 1.0,
                noise: Optional[Tensor] = None, **kwargs) -> Tensor:
        """Denoising-score-matching loss per sample, shape [B] — diffusion.py:65-97.

        `noise` (optional, same shape as x) replaces the internal `torch.randn_like(x)`
        (diffusion.py:76) so tests can fix it; by default it is drawn exactly like the reference.
        `x_mask` (bool, broadcastable to x; diffusion.py:80-83): masked-out elements count with weight 0.01. Like the
        reference, it is also forwarded to `net` with the other kwargs.

        Gradients (the reference trains whatever `net` it is given):
          * fused `WaveNetNoise`: the loss carries an autograd node whose backward is the CUDA backward pass;
          * any `net` that PyTorch autograd can differentiate (a reference backbone, any torch module): the net runs under
            autograd, the loss value and d loss / d F come from the fused kernels (`_GenericDsmLoss`);
          * a CUDA-kernel backbone without a backward (this package's `UNet1dBase`: sampling only) raises here, at call
            time, instead of returning a loss that cannot be back-propagated.
        """
        x = N.require_cuda_f32(x, "x")
        N.ensure_device(x.device)
        if self.dynamic_threshold != 0.0:
            raise NotImplementedError("dynamic_threshold != 0 is not supported in the fused loss")
        B, n_per = x.shape[0], x[0].numel()
        if noise is None:
            noise = torch.randn_like(x)
        noise = N.require_cuda_f32(noise, "noise")
        sig, _ = to_batch(B, x.device, xs=sigmas)
        mask = None
        if kwargs.get("x_mask") is not None:
            mask = torch.as_tensor(kwargs["x_mask"], device=x.device).to(torch.bool).expand_as(x).to(torch.uint8).contiguous()
        # a plain callable has no parameters of its own: whether a graph is needed shows on its output (pred.requires_grad)
        wants_grad = torch.is_grad_enabled() and isinstance(net, nn.Module) and any(p.requires_grad for p in net.parameters())
        fused_train = getattr(net, "_adb_dsm_loss", None)
        if fused_train is not None and wants_grad:
            if mask is not None:
                raise NotImplementedError("x_mask with the fused DiffWave training step is not supported (the waveform "
                                          "module trains on fixed-length clips); compute the loss without gradients or "
                                          "use a differentiable torch backbone")
            # training step: loss with an autograd node whose backward is the CUDA backward pass
            return fused_train(x, sig.contiguous(), noise, float(self.sigma_data))
        lib, st = N.lib(), N.stream_ptr(x.device)
        x_noisy = torch.empty_like(x)
        net_in = torch.empty_like(x)
        c_noise = torch.empty(B, dtype=torch.float32, device=x.device)
        N.check(lib.adb_edm_noise_in(N.ptr(x), N.ptr(noise), N.ptr(sig), float(self.sigma_data), N.ptr(x_noisy),
                                     N.ptr(net_in), N.ptr(c_noise), B, n_per, st))
        if inference:
            pred = net(net_in, c_noise, cond_drop_prob=0., **kwargs)
        else:
            pred = net(net_in, c_noise, **kwargs)
        if pred.shape != x.shape:
            raise N.AdbError(f"net returned {tuple(pred.shape)}, expected {tuple(x.shape)}")
        if wants_grad and not pred.requires_grad:
            raise NotImplementedError(
                f"{type(net).__name__} has trainable parameters but its forward builds no autograd graph: this backbone runs "
                "on CUDA kernels without a backward pass (UNet1dBase is sampling-only; the fused training step exists for "
                "WaveNetNoise). Wrap the call in torch.no_grad() for the loss value, or train a PyTorch-differentiable net.")
        sig_c = sig.contiguous()
        if pred.requires_grad:
            return _GenericDsmLoss.apply(pred, x, x_noisy, sig_c, mask, float(self.sigma_data))
        pred = N.require_cuda_f32(pred, "net output")
        loss = torch.empty(B, dtype=torch.float32, device=x.device)
        N.check(lib.adb_edm_dsm_loss_masked(N.ptr(x), N.ptr(x_noisy), N.ptr(pred), N.ptr(sig_c), float(self.sigma_data),
                                            N.ptr(mask), N.ptr(loss), B, n_per, st))
        return loss


class _GenericDsmLoss(torch.autograd.Function):
    """loss[b] of diffusion.py:92-95 on the output F of ANY autograd-capable net: value from adb_edm_dsm_loss_masked,
    d loss / d F from adb_edm_dsm_loss_grad (one fused pass each); PyTorch carries the gradient on into the net."""

    @staticmethod
    def forward(ctx, pred, x, x_noisy, sig, mask, sigma_data):
        f = N.require_cuda_f32(pred.detach(), "net output")
        B, n_per = x.shape[0], x[0].numel()
        loss = torch.empty(B, dtype=torch.float32, device=x.device)
        N.check(N.lib().adb_edm_dsm_loss_masked(N.ptr(x), N.ptr(x_noisy), N.ptr(f), N.ptr(sig), sigma_data, N.ptr(mask),
                                                N.ptr(loss), B, n_per, N.stream_ptr(x.device)))
        ctx.save_for_backward(f, x, x_noisy, sig)
        ctx.mask, ctx.sigma_data = mask, sigma_data
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        f, x, x_noisy, sig = ctx.saved_tensors
        B, n_per = x.shape[0], x[0].numel()
        up = N.require_cuda_f32(grad_loss, "grad_loss").reshape(B)
        d_f = torch.empty_like(f)
        N.check(N.lib().adb_edm_dsm_loss_grad(N.ptr(x), N.ptr(x_noisy), N.ptr(f), N.ptr(sig), ctx.sigma_data, N.ptr(ctx.mask),
                                              N.ptr(up), N.ptr(d_f), B, n_per, N.stream_ptr(x.device)))
        return d_f, None, None, None, None, None
