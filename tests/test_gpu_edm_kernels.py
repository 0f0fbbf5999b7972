"""GPU parity: fused elementwise EDM kernels (through the C ABI) against the CPU oracle."""
import math

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

# fp32 elementwise kernels vs the torch-CPU oracle: same formulas, scalar coefficients rounded in a
# slightly different order (host fp32 vs ATen scalar promotion) -> agreement to ~1 ulp, not bit-exact.
ELEMWISE_TOL = 5e-7


@pytest.fixture(scope="module")
def env():
    from audiodiffuser_b200 import _native as N
    from oracle import edm
    assert torch.cuda.is_available()
    return N, edm, torch.device("cuda:0")


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


@pytest.mark.parametrize("B,C,L", [(2, 1, 1000), (3, 2, 333), (1, 1, 1), (4, 1, 16000)])
def test_precond_in_out(env, B, C, L):
    N, edm, dev = env
    from audiodiffuser_b200 import EluDiffusion
    diff = EluDiffusion(sigma_data=0.2)
    x = _rand((B, C, L), 1, 3.0)
    f = _rand((B, C, L), 2)
    net = lambda xin, cn, **kw: f.to(xin.device)          # noqa: E731
    for sigma in (80.0, 1.0, 0.002):
        want = edm.denoise(x, lambda a, b, **kw: f, 0.2, sigma=sigma)
        got = diff.denoise_fn(x.to(dev), net=net, sigma=sigma, inference=True)
        assert rel_l2(got, want) < ELEMWISE_TOL, (sigma, rel_l2(got, want))
        assert float((got.cpu() - want).abs().max()) <= 2e-7      # a couple of fp32 ulps at |D| <= 1
        got0 = diff.denoise_fn(x.to(dev), net=net, sigma=torch.tensor(sigma, device=dev), inference=True)
        assert torch.equal(got0, got)
    sig = torch.rand(B, generator=torch.Generator().manual_seed(3)) * 5 + 0.01
    want = edm.denoise(x, lambda a, b, **kw: f, 0.2, sigmas=sig, inference=False)
    got = diff.denoise_fn(x.to(dev), net=net, sigmas=sig.to(dev), inference=False)
    assert rel_l2(got, want) < ELEMWISE_TOL
    N.check_async()


def test_net_receives_scaled_input_and_cnoise(env):
    N, edm, dev = env
    from audiodiffuser_b200 import EluDiffusion
    diff = EluDiffusion(sigma_data=0.5)
    x = _rand((2, 1, 257), 4)
    seen = {}

    def net(xin, cn, **kw):
        seen["x"], seen["c"], seen["kw"] = xin.clone(), cn.clone(), kw
        return torch.zeros_like(xin)

    diff.denoise_fn(x.to(dev), net=net, sigma=0.5, inference=True)
    c_skip, c_out, c_in, c_noise = edm.scale_weights(torch.tensor([0.5, 0.5]), 0.5, 3)
    assert torch.equal(seen["x"].cpu(), c_in * x)
    assert torch.allclose(seen["c"].cpu(), c_noise, rtol=1e-6, atol=1e-7)
    assert seen["kw"] == {"cond_drop_prob": 0.0}


def test_cfg_combine(env):
    N, edm, dev = env
    from audiodiffuser_b200 import EluDiffusion
    diff = EluDiffusion(sigma_data=0.2)
    x, f0, f1 = _rand((2, 1, 500), 5), _rand((2, 1, 500), 6), _rand((2, 1, 500), 7)

    def net_cpu(xin, cn, cond_drop_prob=0.0, **kw):
        return f1 if cond_drop_prob == 1.0 else f0

    def net_gpu(xin, cn, cond_drop_prob=0.0, **kw):
        return (f1 if cond_drop_prob == 1.0 else f0).to(xin.device)

    want = edm.denoise(x, net_cpu, 0.2, sigma=0.7, cond_scale=2.5)
    got = diff.denoise_fn(x.to(dev), net=net_gpu, sigma=0.7, inference=True, cond_scale=2.5)
    assert rel_l2(got, want) < ELEMWISE_TOL


def test_exactly_one_sigma(env):
    N, edm, dev = env
    from audiodiffuser_b200 import EluDiffusion
    diff = EluDiffusion(sigma_data=0.2)
    x = _rand((2, 1, 64), 8).to(dev)
    with pytest.raises(AssertionError):
        diff.denoise_fn(x, net=lambda *a, **k: x, inference=True)
    with pytest.raises(AssertionError):
        diff.denoise_fn(x, net=lambda *a, **k: x, sigma=1.0, sigmas=torch.ones(2, device=dev), inference=True)


def test_no_cpu_path(env):
    N, edm, dev = env
    from audiodiffuser_b200 import EluDiffusion
    with pytest.raises(N.AdbError):
        EluDiffusion(0.2).denoise_fn(torch.zeros(1, 1, 8), net=lambda *a, **k: None, sigma=1.0)


@pytest.mark.parametrize("heun,churn", [(True, 0.0), (False, 0.0), (True, 2.0)])
def test_generic_sampler_matches_oracle(env, heun, churn):
    """EDMSampler with an arbitrary torch `net` (generic path): update kernels vs oracle, bit-exact."""
    N, edm, dev = env
    from audiodiffuser_b200 import EluDiffusion, EDMSampler, KarrasSchedule
    steps = 7
    w = _rand((1, 1, 300), 9, 0.3)

    def net_cpu(xin, cn, **kw):
        return torch.tanh(xin * w + cn.view(-1, 1, 1))

    def net_gpu(xin, cn, **kw):
        return torch.tanh(xin.cpu() * w + cn.cpu().view(-1, 1, 1)).to(xin.device)

    noise = _rand((3, 1, 300), 10)
    sig = KarrasSchedule(0.002, 80.0, 7.0, steps)()
    eps = _rand((steps, 3, 1, 300), 11)
    it = iter(eps)
    want = edm.edm_sampler(noise, lambda x, s: edm.denoise(x, net_cpu, 0.2, sigma=float(s)), sig, steps,
                           s_tmin=0.05, s_tmax=50.0, s_churn=churn, s_noise=1.003, use_heun=heun,
                           eps_fn=lambda x: next(it))
    diff = EluDiffusion(0.2)
    smp = EDMSampler(s_tmin=0.05, s_tmax=50.0, s_churn=churn, s_noise=1.003, num_steps=steps, use_heun=heun)
    got = smp(noise.to(dev), fn=diff.denoise_fn, net=net_gpu, sigmas=sig.to(dev), eps=eps.to(dev))
    assert smp.last_nfe == (2 * steps - 1 if heun else steps)
    assert rel_l2(got, want) < 1e-6
    N.check_async()


@pytest.mark.parametrize("alpha", [1.0, 0.5])
def test_generic_alpha_sampler(env, alpha):
    N, edm, dev = env
    from audiodiffuser_b200 import EluDiffusion, EDMAlphaSampler, KarrasSchedule
    steps = 6
    w = _rand((1, 1, 200), 12, 0.3)
    net_cpu = lambda xin, cn, **kw: torch.tanh(xin * w + cn.view(-1, 1, 1))                     # noqa: E731
    net_gpu = lambda xin, cn, **kw: torch.tanh(xin.cpu() * w + cn.cpu().view(-1, 1, 1)).to(xin.device)  # noqa: E731
    noise = _rand((2, 1, 200), 13)
    sig = KarrasSchedule(0.002, 80.0, 7.0, steps)()
    want = edm.edm_alpha_sampler(noise, lambda x, s: edm.denoise(x, net_cpu, 0.2, sigma=float(s)), sig, steps, alpha=alpha)
    smp = EDMAlphaSampler(alpha=alpha, num_steps=steps)
    got = smp(noise.to(dev), fn=EluDiffusion(0.2).denoise_fn, net=net_gpu, sigmas=sig.to(dev))
    assert smp.last_nfe == 2 * (steps - 1)
    assert rel_l2(got, want) < 1e-6


def test_dsm_loss(env):
    N, edm, dev = env
    from audiodiffuser_b200 import EluDiffusion
    B, L = 3, 4001
    x = _rand((B, 1, L), 14, 0.2).clamp(-1, 1)
    noise = _rand((B, 1, L), 15)
    sig = torch.tensor([0.7, 0.03, 5.0])
    w = _rand((1, 1, L), 16, 0.5)
    net_cpu = lambda xin, cn, **kw: torch.tanh(xin * w + cn.view(-1, 1, 1))                     # noqa: E731
    net_gpu = lambda xin, cn, **kw: torch.tanh(xin.cpu() * w + cn.cpu().view(-1, 1, 1)).to(xin.device)  # noqa: E731
    want = edm.dsm_loss(x, noise, sig, net_cpu, 0.2)
    got = EluDiffusion(0.2)(x.to(dev), net_gpu, sigmas=sig.to(dev), noise=noise.to(dev))
    assert torch.allclose(got.cpu(), want, rtol=2e-5)
    N.check_async()


@pytest.mark.parametrize("n", [1, 1003, 64000])
def test_fused_heun_mid_post(env, n):
    """adb_edm_heun_mid / adb_edm_heun_post (what the fused trajectory launches between network evaluations)
    against one oracle Heun step written on raw network outputs."""
    N, edm, dev = env
    sd, s0, s1 = 0.2, 3.0, 1.7
    x, f1, f2 = _rand((1, 1, n), 20, 3.0), _rand((1, 1, n), 21), _rand((1, 1, n), 22)
    fs = iter([f1, f2])
    want = edm.edm_sampler(x / s0, lambda xx, s: edm.denoise(xx, lambda a, b, **kw: next(fs), sd, sigma=float(s)),
                           torch.tensor([s0, s1]), 1)
    # one step s0 -> s1 of a two-entry schedule is Heun (s1 != 0); edm_sampler scales the noise by sig[0]
    lib, st = N.lib(), N.stream_ptr(dev)
    xd, f1d, f2d = x.to(dev), f1.to(dev), f2.to(dev)
    d, x1, out = torch.empty_like(xd), torch.empty_like(xd), torch.empty_like(xd)
    h = N.ctypes.c_float(N.ctypes.c_float(s1).value - N.ctypes.c_float(s0).value).value
    N.check(lib.adb_edm_heun_mid(N.ptr(xd), N.ptr(f1d), s0, sd, h, N.ptr(d), N.ptr(x1), n, st))
    N.check(lib.adb_edm_heun_post(N.ptr(xd), N.ptr(d), N.ptr(f2d), s1, sd, h, N.ptr(out), n, st))
    N.check_async()
    assert rel_l2(out, want) < 1e-6


def test_churn_rng_matches_philox_oracle_and_is_shard_invariant(env):
    """adb_edm_churn_rng (in-kernel Philox churn noise, replaces randn_like of sampler_edm.py:346): values against the numpy
    restatement, determinism, and independence of where in a batch / on which shard a sample is drawn."""
    import numpy as np
    from oracle.philox import churn_normals
    N, edm, dev = env
    lib, st = N.lib(), N.stream_ptr(dev)
    seed, step = (1 << 61) + 12345, 7
    for B, n_per, s0 in ((3, 1000, 0), (2, 333, 5), (1, 5, (1 << 33) + 2)):        # vectorised, ragged/unaligned, 64-bit sample index
        x = _rand((B, n_per), 3).to(dev)
        out = torch.empty_like(x)
        N.check(lib.adb_edm_churn_rng(N.ptr(x), N.ptr(out), 0.75, 1.003, seed, step, s0, B, n_per, st))
        N.check_async()
        eps = np.stack([churn_normals(seed, s0 + b, step, n_per) for b in range(B)])
        want = x.cpu().numpy() + np.float32(0.75) * (np.float32(1.003) * eps)
        assert np.abs(out.cpu().numpy() - want).max() < 2e-5          # Box-Muller through different libm's
        again = torch.empty_like(x)
        N.check(lib.adb_edm_churn_rng(N.ptr(x), N.ptr(again), 0.75, 1.003, seed, step, s0, B, n_per, st))
        assert torch.equal(out, again)
    # a sample's noise depends on its GLOBAL index only: batch [0, 4) == shards [0, 2) + [2, 4) with sample0 = 2
    x = torch.zeros(4, 4096, device=dev)
    full, lo, hi = torch.empty_like(x), torch.empty(2, 4096, device=dev), torch.empty(2, 4096, device=dev)
    N.check(lib.adb_edm_churn_rng(N.ptr(x), N.ptr(full), 1.0, 1.0, seed, 0, 0, 4, 4096, st))
    N.check(lib.adb_edm_churn_rng(N.ptr(x), N.ptr(lo), 1.0, 1.0, seed, 0, 0, 2, 4096, st))
    N.check(lib.adb_edm_churn_rng(N.ptr(x), N.ptr(hi), 1.0, 1.0, seed, 0, 2, 2, 4096, st))
    N.check_async()
    assert torch.equal(full[:2], lo) and torch.equal(full[2:], hi)
    assert not torch.equal(full[0], full[1])


def test_churn_rng_statistics(env):
    """Mean / variance / kurtosis / chi-square of 2^22 in-kernel normals, and no correlation between samples or steps."""
    N, edm, dev = env
    lib, st = N.lib(), N.stream_ptr(dev)
    B, n = 4, 1 << 20
    zero = torch.zeros(B, n, device=dev)
    a, b = torch.empty_like(zero), torch.empty_like(zero)
    N.check(lib.adb_edm_churn_rng(N.ptr(zero), N.ptr(a), 1.0, 1.0, 42, 3, 0, B, n, st))
    N.check(lib.adb_edm_churn_rng(N.ptr(zero), N.ptr(b), 1.0, 1.0, 42, 4, 0, B, n, st))
    N.check_async()
    z = a.double().flatten()
    m = z.numel()
    assert abs(float(z.mean())) < 4 / math.sqrt(m)
    assert abs(float(z.var()) - 1.0) < 4 * math.sqrt(2.0 / m)
    assert abs(float((z ** 4).mean()) - 3.0) < 4 * math.sqrt(96.0 / m)
    edges = torch.tensor([-1e9, -2.5, -2.0, -1.5, -1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 1e9], dtype=torch.float64, device=dev)
    obs = torch.histogram(z.cpu(), bins=edges.cpu())[0]
    cdf = 0.5 * (1 + torch.erf(edges.cpu() / math.sqrt(2)))
    exp = (cdf[1:] - cdf[:-1]) * m
    chi2 = float(((obs - exp) ** 2 / exp).sum())
    assert chi2 < 40.0, chi2                                           # 11 degrees of freedom: P(chi2 > 40) ~ 4e-5
    for u, v in ((a[0], a[1]), (a[0], b[0]), (a[2, :-1], a[2, 1:])):       # across samples, across steps, neighbouring elements
        r = float((u.double() * v.double()).mean())
        assert abs(r) < 5 / math.sqrt(u.numel()), r


def test_default_sampler_churn_needs_no_eps_tensor(env):
    """EDMSampler with the reference's default churn (s_churn = 150) and no `eps`: fused trajectory and generic Python loop
    draw the same in-kernel noise (same seed) and agree; the result is reproducible under torch.manual_seed, depends on the
    seed, and is invariant to sharding the batch (sample_offset). No [num_steps, B, C, L] tensor is involved."""
    N, edm, dev = env
    from audiodiffuser_b200 import EluDiffusion, EDMSampler, KarrasSchedule, WaveNetNoise
    from oracle.weights import make_wavenet_state_dict
    net = WaveNetNoise(64, 3, 2, precision="fp32")
    net.load_state_dict(make_wavenet_state_dict(64, 3, seed=3), strict=True)
    net = net.to(dev)
    diff = EluDiffusion(0.2)
    steps = 6
    sig = KarrasSchedule(0.002, 80.0, 7.0, steps)().to(dev)
    noise = _rand((4, 1, 500), 21).to(dev)
    smp = EDMSampler(num_steps=steps)                                  # reference defaults: s_churn = 150, s_noise = 1.04
    fused = smp(noise, fn=diff.denoise_fn, net=net, sigmas=sig, churn_seed=99)
    generic = smp(noise, fn=diff.denoise_fn, net=net, sigmas=sig, churn_seed=99, _force_generic=True)
    N.check_async()
    assert rel_l2(fused, generic) < 1e-5, rel_l2(fused, generic)
    assert rel_l2(smp(noise, fn=diff.denoise_fn, net=net, sigmas=sig, churn_seed=100), fused) > 1e-2
    halves = torch.cat([smp(noise[:2].contiguous(), fn=diff.denoise_fn, net=net, sigmas=sig, churn_seed=99),
                        smp(noise[2:].contiguous(), fn=diff.denoise_fn, net=net, sigmas=sig, churn_seed=99, sample_offset=2)])
    assert rel_l2(halves, fused) < 1e-6, rel_l2(halves, fused)
    torch.manual_seed(5)
    r1 = smp(noise, fn=diff.denoise_fn, net=net, sigmas=sig)
    torch.manual_seed(5)
    r2 = smp(noise, fn=diff.denoise_fn, net=net, sigmas=sig)
    assert torch.equal(r1, r2) and not torch.equal(r1, fused)
    # a churn window that contains no sigma: s_churn > 0 but no step draws noise (was a spurious "needs eps" error)
    quiet = EDMSampler(s_tmin=1e3, s_tmax=1e4, s_churn=30.0, num_steps=steps)
    plain = EDMSampler(s_churn=0.0, num_steps=steps)
    assert torch.equal(quiet(noise, fn=diff.denoise_fn, net=net, sigmas=sig), plain(noise, fn=diff.denoise_fn, net=net, sigmas=sig))


def test_dsm_loss_x_mask_and_generic_autograd(env):
    """Diffusion.forward on an arbitrary differentiable torch net (diffusion.py:65-97 trains whatever net it gets): the loss
    value with and without x_mask against the oracle formula, and the gradients that reach the net's parameters through
    _GenericDsmLoss against PyTorch autograd of the same loss written in torch ops."""
    N, edm, dev = env
    from audiodiffuser_b200 import EluDiffusion
    torch.manual_seed(0)
    B, C, L = 3, 2, 257
    net_mod = torch.nn.Conv1d(C, C, 5, padding=2).to(dev)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = net_mod

        def forward(self, x, t, **kw):
            return self.conv(x) * (1.0 + 0.1 * t.view(-1, 1, 1))

    net = Net()
    diff = EluDiffusion(sigma_data=0.2)
    x = (_rand((B, C, L), 5) * 0.3).to(dev)
    noise = _rand((B, C, L), 6).to(dev)
    sig = torch.tensor([0.05, 0.7, 9.0], device=dev)
    mask = (torch.arange(L, device=dev)[None, None, :] < torch.tensor([257, 100, 31], device=dev)[:, None, None])   # [B,1,L] padding mask

    def torch_loss(m):
        sp = sig.view(-1, 1, 1)
        xn = x + sp * noise
        sd = 0.2
        c_skip = sd ** 2 / (sp ** 2 + sd ** 2)
        c_out = sp * sd * (sd ** 2 + sp ** 2) ** -0.5
        c_in = (sp ** 2 + sd ** 2) ** -0.5
        den = (c_skip * xn + c_out * net(c_in * xn, torch.log(sig) * 0.25)).clamp(-1, 1)
        w = torch.ones_like(x) if m is None else (torch.ones_like(x) * m + torch.ones_like(x) * (~m) * 0.01)
        lam = (sig ** 2 + sd ** 2) * (sig * sd) ** -2
        return ((den - x) ** 2 * w).flatten(1).sum(1) * lam / (C * L)

    for m in (None, mask):
        kw = {} if m is None else {"x_mask": m}
        net.zero_grad()
        want = torch_loss(m)
        want.mean().backward()
        gw, gb = net.conv.weight.grad.clone(), net.conv.bias.grad.clone()
        with torch.no_grad():
            val = diff(x, net, sigmas=sig, noise=noise, **kw)
        assert not val.requires_grad and rel_l2(val, want) < 1e-5, rel_l2(val, want)
        net.zero_grad()
        got = diff(x, net, sigmas=sig, noise=noise, **kw)
        assert got.requires_grad and rel_l2(got, want) < 1e-5
        got.mean().backward()
        assert rel_l2(net.conv.weight.grad, gw) < 1e-4, rel_l2(net.conv.weight.grad, gw)
        assert rel_l2(net.conv.bias.grad, gb) < 1e-4
    N.check_async()


def test_kernel_backbone_without_backward_raises_at_call_time(env):
    """A CUDA-kernel backbone with parameters but no backward pass (UNet1dBase is sampling-only) must refuse a training call
    when it is made, not hand back a loss that fails later in .backward()."""
    N, edm, dev = env
    from audiodiffuser_b200 import EluDiffusion, UNet1dBase
    from oracle.weights import UNET_SMALL
    net = UNet1dBase(precision="fp32", **UNET_SMALL).to(dev)
    diff = EluDiffusion(0.2)
    x = _rand((1, 2, 256), 1).to(dev)
    with pytest.raises(NotImplementedError, match="sampling-only"):
        diff(x, net, sigmas=torch.tensor([0.5], device=dev))
    with torch.no_grad():
        assert diff(x, net, sigmas=torch.tensor([0.5], device=dev)).shape == (1,)
