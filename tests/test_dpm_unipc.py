"""DPMSampler / UniPCSampler: oracle and CUDA path against goldens generated from the reference's own
classes (oracle/make_golden_dpm.py): every solver family, order, prediction form and time spacing the reference has."""
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle.make_golden_dpm import DPM_CASES, UNIPC_CASES

FP32_TOL = 1e-5          # rel-L2 on the final clamped waveforms, fp32 kernels (north_star's fp32 gate)


def _schedule(points):
    from oracle import edm
    return edm.karras_schedule(0.01, 20.0, 5.0, points)


def test_oracle_dpm_and_unipc_match_reference():
    from oracle import dpm_solvers as ds, edm, wavenet as owav
    from oracle.weights import make_wavenet_state_dict
    g = load_golden("dpm_unipc_small")
    C, layers, cycle, B, L, seed = (int(v) for v in g["cfg"])
    net_fn = owav.make_net_fn(make_wavenet_state_dict(C, layers, seed), cycle)
    calls = [0]

    def den(x, s):
        calls[0] += 1
        return edm.denoise(x, net_fn, 0.2, sigma=s)

    noise = torch.from_numpy(g["noise"])
    with torch.no_grad():
        for name, kw, points in DPM_CASES:
            calls[0] = 0
            out = ds.dpm_sampler(noise, den, _schedule(points), **kw)
            assert rel_l2(out, g["dpm_" + name]) < 1e-6, name
            assert calls[0] == int(g["nfe_dpm_" + name]), name
        for name, kw, points in UNIPC_CASES:
            calls[0] = 0
            out = ds.unipc_sampler(noise, den, _schedule(points), **kw)
            assert rel_l2(out, g["unipc_" + name]) < 1e-6, name
            assert calls[0] == int(g["nfe_unipc_" + name]) == kw["num_steps"] - (0 if kw["log_time_spacing"] else 1), name


def test_order_schedule_and_constructor_conventions():
    from audiodiffuser_b200.components.sampler_edm import DPMSampler, UniPCSampler
    assert DPMSampler(1.0, order=3, num_steps=9)._orders() == [3, 3, 2, 1]            # sampler_edm.py:776-781
    assert DPMSampler(1.0, order=3, num_steps=8)._orders() == [3, 3, 2]
    assert DPMSampler(1.0, order=3, num_steps=10)._orders() == [3, 3, 3, 1]
    assert DPMSampler(1.0, order=2, num_steps=7)._orders() == [2, 2, 2, 1]
    assert DPMSampler(1.0, order=1, num_steps=4)._orders() == [1, 1, 1, 1]
    with pytest.raises(ValueError):
        DPMSampler(1.0, order=4, num_steps=8)._orders()
    assert DPMSampler(1.0, num_steps=10, log_time_spacing=False).num_steps == 9         # :514
    assert UniPCSampler(num_steps=10, log_time_spacing=False).num_steps == 9            # :830
    with pytest.raises(Exception):                                                      # no CPU path
        DPMSampler(1.0, num_steps=4)(torch.zeros(1, 1, 8), fn=None, net=None, sigmas=_schedule(5))


def test_host_coefficient_folding_on_cpu(monkeypatch):
    """The solvers' host logic — time grid, phi-functions, UniPC solves, history shifts, folding of every update into
    `a x + sum c_k m_k` — without a GPU: the one-launch update and the device checks are replaced by torch stand-ins, the
    denoiser is the oracle's. Every golden case must come out (the scalars are folded in double precision)."""
    from audiodiffuser_b200 import _native as N
    from audiodiffuser_b200.components import sampler_dpm
    from oracle import edm, wavenet as owav
    from oracle.weights import make_wavenet_state_dict

    def lincomb_cpu(x, a, terms, clamp=False):
        assert len(terms) <= 4                                             # what one adb_edm_lincomb_n launch can take
        out = float(a) * x.double()
        for c, m in terms:
            out = out + float(c) * m.double()
        out = out.float()
        return out.clamp(-1.0, 1.0) if clamp else out

    monkeypatch.setattr(sampler_dpm, "lincomb", lincomb_cpu)
    monkeypatch.setattr(N, "require_cuda_f32", lambda t, name: t)
    monkeypatch.setattr(N, "ensure_device", lambda d: None)
    g = load_golden("dpm_unipc_small")
    C, layers, cycle, B, L, seed = (int(v) for v in g["cfg"])
    net_fn = owav.make_net_fn(make_wavenet_state_dict(C, layers, seed), cycle)
    fn = lambda x, net, sigma, inference, cond_scale, **kw: edm.denoise(x, net, 0.2, sigma=sigma)      # noqa: E731
    noise = torch.from_numpy(g["noise"])
    for name, kw, points in DPM_CASES:
        smp = sampler_dpm.DPMSampler(cond_scale=1.0, **kw)
        out = smp(noise, fn=fn, net=net_fn, sigmas=_schedule(points))
        assert rel_l2(out, g["dpm_" + name]) < 3e-6, (name, rel_l2(out, g["dpm_" + name]))
        assert smp.last_nfe == int(g["nfe_dpm_" + name]), name
    for name, kw, points in UNIPC_CASES:
        smp = sampler_dpm.UniPCSampler(cond_scale=1.0, **kw)
        out = smp(noise, fn=fn, net=net_fn, sigmas=_schedule(points))
        assert rel_l2(out, g["unipc_" + name]) < 3e-6, (name, rel_l2(out, g["unipc_" + name]))
        assert smp.last_nfe == int(g["nfe_unipc_" + name]), name


def _gpu_net(g, dev, precision="fp32"):
    from audiodiffuser_b200.backbones.wavenet import WaveNetNoise
    from audiodiffuser_b200.components.diffusion import EluDiffusion
    from oracle.weights import make_wavenet_state_dict
    C, layers, cycle, B, L, seed = (int(v) for v in g["cfg"])
    net = WaveNetNoise(C, layers, cycle, precision=precision)
    net.load_state_dict(make_wavenet_state_dict(C, layers, seed), strict=True)
    return net.to(dev), EluDiffusion(sigma_data=0.2)


@pytest.mark.gpu
def test_lincomb_n_kernel_against_torch():
    from audiodiffuser_b200.components.sampler_dpm import lincomb
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(0)
    for n in (1, 5, 4096, 16000 * 3 + 2):
        x = torch.randn(n, generator=gen).to(dev)
        ms = [torch.randn(n, generator=gen).to(dev) for _ in range(4)]
        cs = [0.3, -1.7, 2.5, -0.01]
        for k in range(5):
            want = 0.9 * x.double() + sum(c * m.double() for c, m in zip(cs[:k], ms[:k]))
            got = lincomb(x, 0.9, list(zip(cs[:k], ms[:k])))
            assert (got.double() - want).abs().max() < 2e-6 * max(1.0, float(want.abs().max())), (n, k)
            assert torch.equal(lincomb(x, 0.9, list(zip(cs[:k], ms[:k])), clamp=True), got.clamp(-1, 1))
    odd = torch.randn(4099, generator=gen).to(dev)[1:]                                  # unaligned view -> scalar kernel
    assert torch.allclose(lincomb(odd.contiguous(), 2.0, [(1.0, odd.contiguous())]), 3 * odd)
    with pytest.raises(Exception):
        lincomb(x, 1.0, [(1.0, m) for m in ms] + [(1.0, ms[0])])                        # five terms


@pytest.mark.gpu
def test_dpm_sampler_gpu_vs_reference_golden():
    from audiodiffuser_b200.components.sampler_edm import DPMSampler
    dev = torch.device("cuda:0")
    g = load_golden("dpm_unipc_small")
    net, diff = _gpu_net(g, dev)
    noise = torch.from_numpy(g["noise"]).to(dev)
    for name, kw, points in DPM_CASES:
        smp = DPMSampler(cond_scale=1.0, **kw)
        out = smp(noise, fn=diff.denoise_fn, net=net, sigmas=_schedule(points).to(dev))
        assert rel_l2(out, g["dpm_" + name]) < FP32_TOL, (name, rel_l2(out, g["dpm_" + name]))
        assert smp.last_nfe == int(g["nfe_dpm_" + name]), name
        assert float(out.abs().max()) <= 1.0


@pytest.mark.gpu
def test_unipc_sampler_gpu_vs_reference_golden():
    from audiodiffuser_b200.components.sampler_edm import UniPCSampler
    dev = torch.device("cuda:0")
    g = load_golden("dpm_unipc_small")
    net, diff = _gpu_net(g, dev)
    noise = torch.from_numpy(g["noise"]).to(dev)
    for name, kw, points in UNIPC_CASES:
        smp = UniPCSampler(cond_scale=1.0, **kw)
        out = smp(noise, fn=diff.denoise_fn, net=net, sigmas=_schedule(points).to(dev))
        assert rel_l2(out, g["unipc_" + name]) < FP32_TOL, (name, rel_l2(out, g["unipc_" + name]))
        assert smp.last_nfe == int(g["nfe_unipc_" + name]), name
    # the reference's own state layout for UniPC is 4-D; the same call works on [B,1,1,L] through a reshaping net
    name, kw, points = UNIPC_CASES[2]
    net4 = lambda x, t, **k: net(x.reshape(x.shape[0], 1, -1), t, **k).reshape(x.shape)      # noqa: E731
    out4 = UniPCSampler(**kw)(noise[:, :, None, :], fn=diff.denoise_fn, net=net4, sigmas=_schedule(points).to(dev))
    assert rel_l2(out4[:, :, 0, :], g["unipc_" + name]) < FP32_TOL


@pytest.mark.gpu
def test_dpm_and_unipc_bf16_path_within_gate():
    """The tensor-core backbone needs C = 256 (the goldens are C = 64): oracle run here on the same seeded inputs."""
    from audiodiffuser_b200.backbones.wavenet import WaveNetNoise
    from audiodiffuser_b200.components.diffusion import EluDiffusion
    from audiodiffuser_b200.components.sampler_edm import DPMSampler, UniPCSampler
    from oracle import dpm_solvers as ds, edm, wavenet as owav
    from oracle.weights import make_wavenet_state_dict
    dev = torch.device("cuda:0")
    C, layers, cycle, seed, B, L = 256, 3, 3, 77, 2, 640
    sd = make_wavenet_state_dict(C, layers, seed)
    net = WaveNetNoise(C, layers, cycle, precision="bf16")
    net.load_state_dict(sd, strict=True)
    net, diff = net.to(dev), EluDiffusion(sigma_data=0.2)
    net_fn = owav.make_net_fn(sd, cycle)
    den = lambda x, s: edm.denoise(x, net_fn, 0.2, sigma=s)                      # noqa: E731
    noise = torch.randn(B, 1, L, generator=torch.Generator().manual_seed(seed + 1))
    kw = dict(order=3, num_steps=6, multisteps=True, x0_pred=True, log_time_spacing=True)
    with torch.no_grad():
        want = ds.dpm_sampler(noise, den, _schedule(7), **kw)
        want_pc = ds.unipc_sampler(noise, den, _schedule(7), num_steps=6, order=2)
    out = DPMSampler(cond_scale=1.0, **kw)(noise.to(dev), fn=diff.denoise_fn, net=net, sigmas=_schedule(7).to(dev))
    assert rel_l2(out, want) < 2e-2, rel_l2(out, want)
    out = UniPCSampler(num_steps=6, order=2)(noise.to(dev), fn=diff.denoise_fn, net=net, sigmas=_schedule(7).to(dev))
    assert rel_l2(out, want_pc) < 2e-2, rel_l2(out, want_pc)
