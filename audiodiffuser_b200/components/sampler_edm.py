"""EDM samplers behind the reference's class API.

Mirrors `src/models/components/sampler_edm.py`: `EDMSampler` (:302-397, Heun/Euler + churn) and
`EDMAlphaSampler` (:229-300, general 2nd-order Runge-Kutta). Same constructors and
`forward(noise, fn, net, sigmas, **kwargs)` / `step(...)` signatures.

Two execution modes, chosen per call:
  * fused trajectory — `fn` is `EluDiffusion.denoise_fn` of this package and `net` is the fused
    DiffWave backbone: the whole N-step loop (preconditioning, 2N-1 network evaluations, Heun
    updates, churn) runs device-side through one C call (adb_wavenet_sample_edm), with per-step
    scalars computed once on the host — no per-step host synchronisation.
  * generic — any other `fn` / `net` (e.g. a reference backbone): the loop stays in Python, the
    update arithmetic runs in the fused elementwise kernels (adb_edm_euler / adb_edm_rk2 / ...).
"""
from math import sqrt
from typing import Callable, List, Optional

import ctypes
import torch
import torch.nn as nn
from torch import Tensor

from .. import _native as N


def _host_sigmas(sigmas: Tensor) -> List[float]:
    """One device->host copy per trajectory (the reference syncs 2-4 times per step)."""
    return [float(v) for v in sigmas.detach().to(torch.float32).cpu().tolist()]


def _f32(v: float) -> float:
    """Round a python double to fp32 (the reference's scalars live in fp32 tensors)."""
    return ctypes.c_float(v).value


def _draw_seed() -> int:
    """63-bit seed for the in-kernel churn noise, taken from torch's default CPU generator (so torch.manual_seed governs it)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def _fused_target(fn, net):
    diff = getattr(fn, "__self__", None)
    if diff is None or getattr(fn, "__name__", "") != "denoise_fn":
        return None
    if not hasattr(diff, "sigma_data") or getattr(diff, "dynamic_threshold", 0.0) != 0.0:
        return None
    if getattr(net, "_adb_fused_sample", None) is None:
        return None
    return diff


class EDMSampler(nn.Module):
    """EDM stochastic sampler (Heun + churn) — sampler_edm.py:302-397."""

    def __init__(self, s_tmin: float = 0, s_tmax: float = float('inf'), s_churn: float = 150.0, s_noise: float = 1.04,
                 num_steps: int = 200, cond_scale: float = 1.0, use_heun: bool = True):
        super().__init__()
        self.s_tmin = s_tmin
        self.s_tmax = s_tmax
        self.s_noise = s_noise
        self.s_churn = s_churn
        self.num_steps = num_steps
        self.cond_scale = cond_scale
        self.use_heun = use_heun
        self.last_nfe = 0

    def _gamma(self, sigma: float) -> float:
        on = _f32(min(self.s_churn / self.num_steps, sqrt(2) - 1))       # sampler_edm.py:383-387
        return on if (self.s_tmin <= sigma <= self.s_tmax) else 0.0

    def step(self, x: Tensor, fn: Callable, net: nn.Module, sigma: float, sigma_next: float, gamma: float,
             eps: Optional[Tensor] = None, rng=None, **kwargs) -> Tensor:
        """One step (sampler_edm.py:333-369). sigma / sigma_next / gamma: python floats or 0-dim tensors.

        The churn noise of a step with gamma > 0 is `eps` when given (parity tests), else it is drawn inside the update
        kernel (adb_edm_churn_rng) from `rng = (seed, step_index, sample_offset)`; `rng=None` draws a fresh seed from torch's
        default generator (the reference calls `randn_like(x)` here, sampler_edm.py:346)."""
        sigma, sigma_next, gamma = (_f32(float(v)) for v in (sigma, sigma_next, gamma))
        x = N.require_cuda_f32(x, "x")
        lib, st, n = N.lib(), N.stream_ptr(x.device), x.numel()
        if gamma > 0:
            sigma_hat = _f32(sigma + _f32(gamma * sigma))
            a = _f32(sqrt(_f32(_f32(sigma_hat * sigma_hat) - _f32(sigma * sigma))))
            x_hat = torch.empty_like(x)
            if eps is not None:
                eps = N.require_cuda_f32(eps, "eps")
                noise = torch.empty_like(x)
                N.check(lib.adb_edm_scale(N.ptr(eps), float(self.s_noise), N.ptr(noise), n, st))
                N.check(lib.adb_edm_axpy(N.ptr(x), N.ptr(noise), a, N.ptr(x_hat), n, st))
            else:
                seed, step_index, sample_offset = rng if rng is not None else (_draw_seed(), 0, 0)
                B = x.shape[0]
                N.check(lib.adb_edm_churn_rng(N.ptr(x), N.ptr(x_hat), a, float(self.s_noise), int(seed), int(step_index),
                                              int(sample_offset), B, n // B, st))
        else:
            sigma_hat, x_hat = sigma, x
        den = N.require_cuda_f32(fn(x_hat, net=net, sigma=sigma_hat, inference=True, cond_scale=self.cond_scale,
                                    **kwargs), "denoised")
        self.last_nfe += 1
        h = _f32(sigma_next - sigma_hat)
        d = torch.empty_like(x)
        x_next = torch.empty_like(x)
        N.check(lib.adb_edm_euler(N.ptr(x_hat), N.ptr(den), sigma_hat, h, N.ptr(d), N.ptr(x_next), n, st))
        if sigma_next != 0 and self.use_heun:
            den2 = N.require_cuda_f32(fn(x_next, net=net, sigma=sigma_next, inference=True,
                                         cond_scale=self.cond_scale, **kwargs), "denoised")
            self.last_nfe += 1
            out = torch.empty_like(x)
            N.check(lib.adb_edm_rk2(N.ptr(x_hat), N.ptr(d), N.ptr(x_next), N.ptr(den2), sigma_next, h, 0.5, 0.5,
                                    N.ptr(out), n, st))
            x_next = out
        return x_next

    @torch.no_grad()
    def forward(self, noise: Tensor, fn: Callable, net: nn.Module, sigmas: Tensor, eps: Optional[Tensor] = None,
                churn_seed: Optional[int] = None, sample_offset: int = 0, **kwargs) -> Tensor:
        """noise [B,C,L] ~ N(0,1), sigmas [N] -> x [B,C,L] (sampler_edm.py:371-397).

        Churn noise (steps with gamma > 0): by default it is generated INSIDE the update kernel (Philox, keyed by
        `churn_seed` — drawn from torch's default generator when None, so `torch.manual_seed` makes a run reproducible —
        and by the global sample index `sample_offset + b`, which makes a waveform independent of batch sharding). Nothing
        of size [num_steps, *noise.shape] is allocated: the reference draws one `randn_like(x)` per step
        (sampler_edm.py:346), O(1) memory in the number of steps, and so does this. `eps` (optional,
        [num_steps, *noise.shape]) supplies the noise explicitly instead (parity tests against the reference).
        """
        noise = N.require_cuda_f32(noise, "noise")
        N.ensure_device(noise.device)
        sig = _host_sigmas(sigmas)
        if len(sig) < self.num_steps:
            raise ValueError(f"schedule has {len(sig)} sigmas but num_steps={self.num_steps} (sampler_edm.py:390)")
        self.last_nfe = 0
        churn = any(self._gamma(s) > 0 for s in sig[:self.num_steps])
        if eps is not None:
            eps = N.require_cuda_f32(eps, "eps")
            if eps.numel() != self.num_steps * noise.numel():
                raise N.AdbError(f"eps must have shape [num_steps, *noise.shape]; got {tuple(eps.shape)}")
        seed = 0
        if churn and eps is None:
            seed = _draw_seed() if churn_seed is None else int(churn_seed)
        diff = _fused_target(fn, net)
        force_generic = kwargs.pop("_force_generic", False)
        # device-resident trajectory: the unconditional call (no conditioning kwargs, no guidance); a backbone may decline (None)
        uncond = getattr(net, "_adb_unconditional", False)      # DiffWave: conditioning kwargs / guidance cannot change its output
        if diff is not None and not force_generic and (uncond or (not kwargs and self.cond_scale == 1.0)):
            res = net._adb_fused_sample(noise, sig, self.num_steps, float(diff.sigma_data), float(self.s_tmin),
                                        float(min(self.s_tmax, 3.0e38)), float(self.s_churn), float(self.s_noise),
                                        bool(self.use_heun), -1.0, eps if churn else None, churn_seed=seed,
                                        sample_offset=sample_offset)
            if res is not None:
                x, self.last_nfe = res
                return x
        sig = sig + [0.0]                              # t_N = 0 (sampler_edm.py:377)
        x = torch.empty_like(noise)
        N.check(N.lib().adb_edm_scale(N.ptr(noise), sig[0], N.ptr(x), x.numel(), N.stream_ptr(x.device)))
        for i in range(self.num_steps):
            x = self.step(x, fn=fn, net=net, sigma=sig[i], sigma_next=sig[i + 1], gamma=self._gamma(sig[i]),
                          eps=eps[i] if eps is not None else None, rng=(seed, i, sample_offset), **kwargs)
        return x


class EDMAlphaSampler(nn.Module):
    """EDM deterministic sampler with a general 2nd-order Runge-Kutta step (alpha = 1: Heun) —
    sampler_edm.py:229-300."""

    def __init__(self, alpha: float = 1.0, num_steps: int = 50, cond_scale: float = 1.0, use_heun: bool = True):
        super().__init__()
        self.alpha = alpha
        self.num_steps = num_steps
        self.cond_scale = cond_scale
        self.use_heun = use_heun
        self.last_nfe = 0

    @torch.no_grad()
    def step(self, x: Tensor, fn: Callable, net: nn.Module, sigma: float, sigma_next: float, **kwargs) -> Tensor:
        """sampler_edm.py:251-282."""
        sigma, sigma_next = _f32(float(sigma)), _f32(float(sigma_next))
        x = N.require_cuda_f32(x, "x")
        lib, st, n = N.lib(), N.stream_ptr(x.device), x.numel()
        h = _f32(sigma_next - sigma)
        den = N.require_cuda_f32(fn(x, net=net, sigma=sigma, inference=True, cond_scale=self.cond_scale, **kwargs),
                                 "denoised")
        self.last_nfe += 1
        ah = _f32(self.alpha * h)
        sigma_p = _f32(sigma + ah)
        d = torch.empty_like(x)
        x_p = torch.empty_like(x)
        if sigma_p != 0 and self.use_heun:
            N.check(lib.adb_edm_euler(N.ptr(x), N.ptr(den), sigma, ah, N.ptr(d), N.ptr(x_p), n, st))
            den_p = N.require_cuda_f32(fn(x_p, net=net, sigma=sigma_p, inference=True, cond_scale=self.cond_scale,
                                          **kwargs), "denoised")
            self.last_nfe += 1
            out = torch.empty_like(x)
            w1 = 0.5 / self.alpha
            N.check(lib.adb_edm_rk2(N.ptr(x), N.ptr(d), N.ptr(x_p), N.ptr(den_p), sigma_p, h, _f32(1 - w1), _f32(w1),
                                    N.ptr(out), n, st))
            return out
        N.check(lib.adb_edm_euler(N.ptr(x), N.ptr(den), sigma, h, N.ptr(d), N.ptr(x_p), n, st))
        return x_p

    @torch.no_grad()
    def forward(self, noise: Tensor, fn: Callable, net: nn.Module, sigmas: Tensor, **kwargs) -> Tensor:
        """sampler_edm.py:284-300 — loops num_steps - 1 times and never appends sigma = 0."""
        noise = N.require_cuda_f32(noise, "noise")
        N.ensure_device(noise.device)
        sig = _host_sigmas(sigmas)
        if len(sig) < self.num_steps:
            raise ValueError(f"schedule has {len(sig)} sigmas but num_steps={self.num_steps}")
        self.last_nfe = 0
        diff = _fused_target(fn, net)
        force_generic = kwargs.pop("_force_generic", False)
        uncond = getattr(net, "_adb_unconditional", False)
        if diff is not None and not force_generic and (uncond or (not kwargs and self.cond_scale == 1.0)):
            res = net._adb_fused_sample(noise, sig, self.num_steps, float(diff.sigma_data), 0.0, 3.0e38, 0.0, 1.0,
                                        bool(self.use_heun), float(self.alpha), None)
            if res is not None:
                x, self.last_nfe = res
                return x
        x = torch.empty_like(noise)
        N.check(N.lib().adb_edm_scale(N.ptr(noise), sig[0], N.ptr(x), x.numel(), N.stream_ptr(x.device)))
        for i in range(self.num_steps - 1):
            x = self.step(x, fn=fn, net=net, sigma=sig[i], sigma_next=sig[i + 1], **kwargs)
        return x


class DPM2MSampler(nn.Module):
    """DPM-Solver++(2M) Karras, deterministic multistep — sampler_edm.py:1056-1131.

    Same constructor and `forward(noise, fn, net, sigmas)`; `sigmas` needs `num_steps + 1` entries (the reference
    indexes `sigmas[i + 1]`, :1124), the last of which may be 0. Per step one denoiser call and ONE fused kernel
    (`adb_edm_lincomb`: x <- a x - e (c0 D - c1 D_old)); the step scalars (t = -ln sigma, h, expm1) are computed on the
    host from the schedule, so the loop never synchronises with the device.
    """

    def __init__(self, num_steps: int = 50, cond_scale: float = 1.0):
        super().__init__()
        self.num_steps = num_steps
        self.cond_scale = cond_scale
        self.last_nfe = 0

    @staticmethod
    def _coefficients(sigma_last: Optional[float], sigma: float, sigma_next: float, have_old: bool):
        """(a, e, c0, c1) of x_next = a x - e (c0 D - c1 D_old) — sampler_edm.py:1089-1108, evaluated in fp32 like
        the reference's 0-dim fp32 tensors (np.float32 arithmetic rounds after every operation)."""
        import numpy as np
        f = np.float32
        with np.errstate(divide="ignore", over="ignore"):
            t, t_next = -np.log(f(sigma)), -np.log(f(sigma_next))           # t_fn, :1082
            h = f(t_next - t)
            s_t, s_next = np.exp(-t), np.exp(-t_next)                        # sigma_fn, :1081
            a = f(min(s_next, s_t) / max(s_next, s_t))                       # t_min / t_max, :1092-1093
            if not have_old or sigma_next == 0:
                return float(a), float(np.expm1(f(-h))), 1.0, 0.0           # :1097-1098
            h_last = f(t - (-np.log(f(sigma_last))))
            h_min, h_max = min(h_last, h), max(h_last, h)
            r = f(h_max / h_min)
            h_d = f(f(h_max + h_min) / f(2))
            c0 = f(f(1) + f(f(1) / f(f(2) * r)))
            c1 = f(f(1) / f(f(2) * r))
            return float(a), float(np.expm1(f(-h_d))), float(c0), float(c1)

    @torch.no_grad()
    def forward(self, noise: Tensor, fn: Callable, net: nn.Module, sigmas: Tensor, **kwargs) -> Tensor:
        noise = N.require_cuda_f32(noise, "noise")
        N.ensure_device(noise.device)
        sig = _host_sigmas(sigmas)
        if len(sig) < self.num_steps + 1:
            raise IndexError(f"schedule has {len(sig)} sigmas but DPM2MSampler indexes sigmas[{self.num_steps}] "
                             "(sampler_edm.py:1124)")
        lib, st, n = N.lib(), N.stream_ptr(noise.device), noise.numel()
        self.last_nfe = 0
        x = torch.empty_like(noise)
        N.check(lib.adb_edm_scale(N.ptr(noise), sig[0], N.ptr(x), n, st))           # :1114
        old = None
        for i in range(self.num_steps):
            den = N.require_cuda_f32(fn(x, net=net, sigma=sig[i], inference=True, cond_scale=self.cond_scale, **kwargs),
                                     "denoised")
            self.last_nfe += 1
            a, e, c0, c1 = self._coefficients(sig[i - 1] if i > 0 else None, sig[i], sig[i + 1], old is not None)
            second = old is not None and sig[i + 1] != 0
            nxt = torch.empty_like(x)
            N.check(lib.adb_edm_lincomb(N.ptr(x), N.ptr(den), N.ptr(old) if second else N.ptr(None), a, e, c0, c1, N.ptr(nxt), n, st))
            x, old = nxt, den
        out = torch.empty_like(x)
        N.check(lib.adb_edm_clamp(N.ptr(x), N.ptr(out), n, st))                      # :1131
        return out


class ADPM2Sampler(nn.Module):
    """Ancestral DPM-Solver-2 ('DPM2 a Karras') — stochastic_sampler_edm.py:30-100, the default sampler of the reference's
    model config (configs/model/diffunet_complex.yaml:23). Same constructor and `forward(noise, fn, net, sigmas)`.

    Two denoiser calls per step; the arithmetic around them is three fused kernels (adb_edm_euler, adb_edm_rk2 with
    weights (0, 1), adb_edm_axpy for the ancestral noise) with step scalars computed on the host in fp32.
    `eps` (optional, [num_steps - 1, *noise.shape]) fixes the ancestral noise for parity tests; by default one
    `randn_like(x)` is drawn per step like stochastic_sampler_edm.py:80.
    """

    def __init__(self, rho: float = 1.0, num_steps: int = 50, cond_scale: float = 1.0, eta: float = 1.0):
        super().__init__()
        self.rho = rho
        self.num_steps = num_steps
        self.cond_scale = cond_scale
        self.eta = eta
        self.last_nfe = 0

    def _sigmas(self, sigma: float, sigma_next: float):
        """(sigma_up, sigma_down, sigma_mid) — get_sigmas (:30-33) and :69, in fp32 like the reference's 0-dim tensors."""
        import numpy as np
        f = np.float32
        s, sn = f(sigma), f(sigma_next)
        inner = f(f(sn ** 2) * f(f(s ** 2) - f(sn ** 2))) / f(s ** 2)
        up = min(sn, f(f(self.eta) * f(np.sqrt(f(inner)))))
        down = f(np.sqrt(f(f(sn ** 2) - f(up ** 2))))
        inv = 1.0 / self.rho
        mid = f(f(f(f(s ** f(inv)) + f(down ** f(inv))) / f(2)) ** f(self.rho))
        return float(up), float(down), float(mid)

    @torch.no_grad()
    def forward(self, noise: Tensor, fn: Callable, net: nn.Module, sigmas: Tensor, eps: Optional[Tensor] = None,
                **kwargs) -> Tensor:
        noise = N.require_cuda_f32(noise, "noise")
        N.ensure_device(noise.device)
        sig = _host_sigmas(sigmas)
        if len(sig) < self.num_steps:
            raise IndexError(f"schedule has {len(sig)} sigmas but ADPM2Sampler indexes sigmas[{self.num_steps - 1}]")
        lib, st, n = N.lib(), N.stream_ptr(noise.device), noise.numel()
        self.last_nfe = 0
        x = torch.empty_like(noise)
        N.check(lib.adb_edm_scale(N.ptr(noise), sig[0], N.ptr(x), n, st))
        for i in range(self.num_steps - 1):
            sigma, sigma_next = sig[i], sig[i + 1]
            up, down, mid = self._sigmas(sigma, sigma_next)
            den = N.require_cuda_f32(fn(x, net=net, sigma=sigma, inference=True, cond_scale=self.cond_scale, **kwargs), "denoised")
            d, x_mid = torch.empty_like(x), torch.empty_like(x)
            N.check(lib.adb_edm_euler(N.ptr(x), N.ptr(den), sigma, _f32(mid - sigma), N.ptr(d), N.ptr(x_mid), n, st))        # :66-70
            den_mid = N.require_cuda_f32(fn(x_mid, net=net, sigma=mid, inference=True, cond_scale=self.cond_scale, **kwargs),
                                         "denoised")
            self.last_nfe += 2
            x_det = torch.empty_like(x)
            N.check(lib.adb_edm_rk2(N.ptr(x), N.ptr(d), N.ptr(x_mid), N.ptr(den_mid), mid, _f32(down - sigma), 0.0, 1.0,
                                    N.ptr(x_det), n, st))                                                                      # :76-78
            e = eps[i] if eps is not None else torch.randn_like(x)
            x = torch.empty_like(x_det)
            N.check(lib.adb_edm_axpy(N.ptr(x_det), N.ptr(N.require_cuda_f32(e, "eps")), up, N.ptr(x), n, st))                 # :79
        out = torch.empty_like(x)
        N.check(lib.adb_edm_clamp(N.ptr(x), N.ptr(out), n, st))
        return out


# The reference keeps these two in the same module (sampler_edm.py:495, :807); re-exported so `_target_` paths carry over.
from .sampler_dpm import DPMSampler, UniPCSampler  # noqa: E402,F401
