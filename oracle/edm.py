"""Functional CPU restatement of the EDM denoiser, samplers, schedule and loss (test infrastructure).

Reference locations (all under /root/reference/src/models/components/):
  scale weights / loss weight   diffusion.py:232-245
  denoise_fn                    diffusion.py:32-63   (clip: utils.py:20-33, to_batch: utils.py:41-52)
  DSM loss (Diffusion.forward)  diffusion.py:65-97
  EDMSampler                    sampler_edm.py:302-397
  EDMAlphaSampler               sampler_edm.py:229-300
  KarrasSchedule                scheduler.py:6-22
  LogNormalDistribution         distribution.py:9-16
`net` is any callable net(x_scaled, c_noise[B], **kw) -> tensor like x.
"""
import math
import torch


def scale_weights(sigmas, sigma_data, ndim):
    """(c_skip, c_out, c_in)[B,1,..], c_noise[B] — diffusion.py:232-241."""
    c_noise = torch.log(sigmas) * 0.25
    s = sigmas.view(*sigmas.shape, *((1,) * (ndim - sigmas.ndim)))
    sd2 = sigma_data ** 2
    c_skip = sd2 / (s ** 2 + sd2)
    c_out = s * sigma_data * (sd2 + s ** 2) ** -0.5
    c_in = (s ** 2 + sd2) ** -0.5
    return c_skip, c_out, c_in, c_noise


def loss_weight(sigmas, sigma_data):
    """(sigma^2 + sd^2) / (sigma*sd)^2 — diffusion.py:243-245."""
    return (sigmas ** 2 + sigma_data ** 2) * (sigmas * sigma_data) ** -2


def clip(x, dynamic_threshold=0.0):
    """utils.py:20-33."""
    if dynamic_threshold == 0.0:
        return x.clamp(-1.0, 1.0)
    flat = x.reshape(x.shape[0], -1)
    scale = torch.quantile(flat.abs(), dynamic_threshold, dim=-1).clamp(min=1.0)
    scale = scale.view(-1, *((1,) * (x.ndim - 1)))
    return x.clamp(-scale, scale) / scale


def denoise(x_noisy, net, sigma_data, sigma=None, sigmas=None, inference=True, cond_scale=1.0,
            dynamic_threshold=0.0, **kw):
    """diffusion.py:32-63. Exactly one of sigma (scalar) / sigmas [B]."""
    assert (sigma is None) != (sigmas is None)
    B = x_noisy.shape[0]
    if sigmas is None:
        sigmas = torch.full((B,), float(sigma), dtype=torch.float32).to(x_noisy.dtype)
    c_skip, c_out, c_in, c_noise = scale_weights(sigmas, sigma_data, x_noisy.ndim)
    if inference:
        pred = net(c_in * x_noisy, c_noise, cond_drop_prob=0.0, **kw)
        if cond_scale != 1.0:
            null = net(c_in * x_noisy, c_noise, cond_drop_prob=1.0, **kw)
            pred = null + (pred - null) * cond_scale
    else:
        pred = net(c_in * x_noisy, c_noise, **kw)
    return clip(c_skip * x_noisy + c_out * pred, dynamic_threshold)


def dsm_loss(x, noise, sigmas, net, sigma_data, **kw):
    """Diffusion.forward with the noise made explicit — diffusion.py:65-97. Returns loss [B]."""
    s = sigmas.view(-1, *((1,) * (x.ndim - 1)))
    x_noisy = x + s * noise
    den = denoise(x_noisy, net, sigma_data, sigmas=sigmas, inference=False, **kw)
    per = ((den - x) ** 2).reshape(x.shape[0], -1).sum(dim=1)
    numel = x[0].numel()
    return per * loss_weight(sigmas, sigma_data) / numel


def karras_schedule(sigma_min, sigma_max, rho=7.0, num_steps=50):
    """scheduler.py:17-22 — fp32 arange, python-float powers."""
    rho_inv = 1.0 / rho
    steps = torch.arange(num_steps, dtype=torch.float32)
    return (sigma_max ** rho_inv + steps / (num_steps - 1) *
            (sigma_min ** rho_inv - sigma_max ** rho_inv)) ** rho


def lognormal_sigmas(mean, std, normal):
    """distribution.py:14-16 with the N(0,1) draw made explicit."""
    return (mean + std * normal).exp()


def edm_sampler(noise, denoise_fn, sigmas, num_steps, s_tmin=0.0, s_tmax=float("inf"), s_churn=0.0,
                s_noise=1.0, use_heun=True, eps_fn=None, trace=None):
    """EDMSampler.forward/step — sampler_edm.py:333-397.

    denoise_fn(x, sigma) -> x0 estimate. eps_fn(x) supplies the churn noise (defaults to
    torch.randn_like, drawn EVERY step exactly like sampler_edm.py:346, also when gamma == 0).
    Returns x; appends the number of denoiser calls to `trace` if given.
    """
    if eps_fn is None:
        eps_fn = torch.randn_like
    sig = torch.cat([sigmas, torch.zeros_like(sigmas[:1])])
    x = sig[0] * noise
    gamma_on = min(s_churn / num_steps, math.sqrt(2.0) - 1.0)
    nfe = 0
    for i in range(num_steps):
        s, s_next = sig[i], sig[i + 1]
        gamma = gamma_on if (s >= s_tmin and s <= s_tmax) else 0.0
        eps = s_noise * eps_fn(x)
        if gamma > 0:
            s_hat = s + gamma * s
            x_hat = x + (s_hat ** 2 - s ** 2) ** 0.5 * eps
        else:
            s_hat, x_hat = s, x
        d = (x_hat - denoise_fn(x_hat, s_hat)) / s_hat
        nfe += 1
        x_next = x_hat + (s_next - s_hat) * d
        if s_next != 0 and use_heun:
            d2 = (x_next - denoise_fn(x_next, s_next)) / s_next
            nfe += 1
            x_next = x_hat + 0.5 * (s_next - s_hat) * (d + d2)
        x = x_next
    if trace is not None:
        trace.append(nfe)
    return x


def edm_alpha_sampler(noise, denoise_fn, sigmas, num_steps, alpha=1.0, use_heun=True, trace=None):
    """EDMAlphaSampler — sampler_edm.py:251-300 (loops num_steps-1, never appends sigma=0)."""
    x = sigmas[0] * noise
    nfe = 0
    for i in range(num_steps - 1):
        s, s_next = sigmas[i], sigmas[i + 1]
        h = s_next - s
        d = (x - denoise_fn(x, s)) / s
        nfe += 1
        s_p = s + alpha * h
        if s_p != 0 and use_heun:
            x_p = x + alpha * h * d
            d_p = (x_p - denoise_fn(x_p, s_p)) / s_p
            nfe += 1
            x = x + h * ((1 - 0.5 / alpha) * d + 0.5 / alpha * d_p)
        else:
            x = x + h * d
    if trace is not None:
        trace.append(nfe)
    return x


def dpm2m_sampler(noise, denoise_fn, sigmas, num_steps, trace=None):
    """DPM2MSampler.forward / step — sampler_edm.py:1056-1131 (DPM-Solver++(2M) Karras). sigmas needs num_steps + 1
    entries. denoise_fn(x, sigma) -> x0 estimate."""
    sigma_fn = lambda t: t.neg().exp()                # noqa: E731
    t_fn = lambda s: s.log().neg()                    # noqa: E731
    x = sigmas[0] * noise
    old = None
    nfe = 0
    for i in range(num_steps):
        s_last, s, s_next = sigmas[i - 1], sigmas[i], sigmas[i + 1]
        den = denoise_fn(x, s)
        nfe += 1
        t, t_next = t_fn(s), t_fn(s_next)
        h = t_next - t
        t_min, t_max = min(sigma_fn(t_next), sigma_fn(t)), max(sigma_fn(t_next), sigma_fn(t))
        if old is None or s_next == 0:
            x = (t_min / t_max) * x - (-h).expm1() * den
        else:
            h_last = t - t_fn(s_last)
            h_min, h_max = min(h_last, h), max(h_last, h)
            r = h_max / h_min
            h_d = (h_max + h_min) / 2
            den_d = (1 + 1 / (2 * r)) * den - (1 / (2 * r)) * old
            x = (t_min / t_max) * x - (-h_d).expm1() * den_d
        old = den
    if trace is not None:
        trace.append(nfe)
    return x.clamp(-1.0, 1.0)


def adpm2_sampler(noise, denoise_fn, sigmas, num_steps, rho=1.0, eta=1.0, eps_fn=None):
    """ADPM2Sampler.forward / step — stochastic_sampler_edm.py:30-100 ('DPM2 a Karras', ancestral). eps_fn(x) supplies the
    ancestral noise (default torch.randn_like, one draw per step)."""
    if eps_fn is None:
        eps_fn = torch.randn_like
    x = sigmas[0] * noise
    for i in range(num_steps - 1):
        s, s_next = sigmas[i], sigmas[i + 1]
        s_up = min(s_next, eta * (s_next ** 2 * (s ** 2 - s_next ** 2) / s ** 2) ** 0.5)
        s_down = (s_next ** 2 - s_up ** 2) ** 0.5
        d = (x - denoise_fn(x, s)) / s
        s_mid = ((s ** (1 / rho) + s_down ** (1 / rho)) / 2) ** rho
        x_mid = x + d * (s_mid - s)
        d_mid = (x_mid - denoise_fn(x_mid, s_mid)) / s_mid
        x = x + d_mid * (s_down - s)
        x = x + eps_fn(x) * s_up
    return x.clamp(-1.0, 1.0)
