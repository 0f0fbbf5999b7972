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
