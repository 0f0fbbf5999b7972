"""SURVEY §8(f) widening items: DPM-Solver++(2M) sampler and the EMA classes, against goldens produced by the
reference's own DPM2MSampler / phema classes (oracle/make_golden_extra.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2


def test_oracle_dpm2m_matches_reference():
    from oracle import edm, wavenet as owav
    from oracle.weights import make_wavenet_state_dict
    g = load_golden("dpm2m_small")
    C, layers, cycle, B, L, seed, N = (int(v) for v in g["cfg"])
    net_fn = owav.make_net_fn(make_wavenet_state_dict(C, layers, seed), cycle)
    den = lambda x, s: edm.denoise(x, net_fn, 0.2, sigma=float(s))          # noqa: E731
    noise = torch.from_numpy(g["noise"])
    with torch.no_grad():
        trace = []
        out = edm.dpm2m_sampler(noise, den, torch.from_numpy(g["sigmas"]), N, trace=trace)
        out2 = edm.dpm2m_sampler(noise, den, torch.from_numpy(g["sigmas2"]), N)
    assert trace[0] == int(g["nfe"]) == N
    assert rel_l2(out, g["out"]) < 1e-5 and rel_l2(out2, g["out2"]) < 1e-5


def test_dpm2m_coefficients_host_math():
    """The host-side step scalars against the reference formulas evaluated with fp32 tensors (sampler_edm.py:1089-1108)."""
    from audiodiffuser_b200 import DPM2MSampler
    s_last, s, s_next = torch.tensor(3.0), torch.tensor(1.5), torch.tensor(0.7)
    t_fn = lambda x: x.log().neg()                                            # noqa: E731
    t, t_next = t_fn(s), t_fn(s_next)
    h, h_last = t_next - t, t - t_fn(s_last)
    r = max(h_last, h) / min(h_last, h)
    want = (float(s_next / s), float((-((max(h_last, h) + min(h_last, h)) / 2)).expm1()), float(1 + 1 / (2 * r)), float(1 / (2 * r)))
    got = DPM2MSampler._coefficients(3.0, 1.5, 0.7, True)
    assert np.allclose(got, want, rtol=2e-6)
    a, e, c0, c1 = DPM2MSampler._coefficients(None, 1.5, 0.0, False)          # last step onto sigma = 0: x <- D
    assert (a, e, c0, c1) == (0.0, -1.0, 1.0, 0.0)


@pytest.mark.gpu
def test_dpm2m_sampler_gpu_vs_reference_golden():
    from audiodiffuser_b200 import DPM2MSampler, EluDiffusion, WaveNetNoise, _native as N
    from oracle.weights import make_wavenet_state_dict
    dev = torch.device("cuda:0")
    g = load_golden("dpm2m_small")
    C, layers, cycle, B, L, seed, steps = (int(v) for v in g["cfg"])
    net = WaveNetNoise(C, layers, cycle, precision="fp32")
    net.load_state_dict(make_wavenet_state_dict(C, layers, seed), strict=True)
    net = net.to(dev)
    diff, smp = EluDiffusion(0.2), DPM2MSampler(num_steps=steps)
    noise = torch.from_numpy(g["noise"]).to(dev)
    out = smp(noise, fn=diff.denoise_fn, net=net, sigmas=torch.from_numpy(g["sigmas"]).to(dev))
    assert smp.last_nfe == int(g["nfe"])
    out2 = smp(noise, fn=diff.denoise_fn, net=net, sigmas=torch.from_numpy(g["sigmas2"]).to(dev))
    N.check_async()
    e1, e2 = rel_l2(out, g["out"]), rel_l2(out2, g["out2"])
    print(f"DPM2M fp32: rel-L2 {e1:.3e} (schedule ending in 0), {e2:.3e} (all second-order)")
    assert e1 < 2e-5 and e2 < 2e-5
    assert float(out.abs().max()) <= 1.0
    with pytest.raises(IndexError):
        smp(noise, fn=diff.denoise_fn, net=net, sigmas=torch.from_numpy(g["sigmas"][:steps]).to(dev))


@pytest.mark.gpu
def test_ema_classes_gpu_vs_reference_golden():
    from audiodiffuser_b200.ema import PowerFunctionEMA, TraditionalEMA
    dev = torch.device("cuda:0")
    g = load_golden("ema_small")
    lin = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3)).to(dev)

    def set_params(flat):
        off = 0
        with torch.no_grad():
            for p in lin.parameters():
                n = p.numel()
                p.copy_(torch.from_numpy(flat[off:off + n]).view(p.shape))
                off += n

    set_params(g["p0"])
    pf, tr = PowerFunctionEMA(lin, stds=[0.05, 0.1]), TraditionalEMA(lin, halflife_Mimg=0.5, rampup_ratio=0.09)
    nimg = 0
    for step in g["params"]:
        set_params(step)
        nimg += 64
        pf.update(cur_nimg=nimg, batch_size=64)
        tr.update(cur_nimg=nimg, batch_size=64)
    assert rel_l2(pf.emas[0], g["pf0"]) < 1e-6 and rel_l2(pf.emas[1], g["pf1"]) < 1e-6
    assert rel_l2(tr.ema, g["trad"]) < 1e-6
    m = tr.get()
    assert rel_l2(torch.cat([p.reshape(-1) for p in m.parameters()]), g["trad"]) < 1e-6
    assert [sfx for _, sfx in pf.get()] == ["-0.050", "-0.100"]
    # checkpoint resume (phema.py:119-123, :162-163): state_dict() -> fresh objects -> load_state_dict() restores every copy,
    # and the next update continues identically; the state also loads into the reference's own classes when they exist
    st_pf, st_tr = pf.state_dict(), tr.state_dict()
    lin2 = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3)).to(dev)
    pf2, tr2 = PowerFunctionEMA(lin2, stds=[0.3, 0.4]), TraditionalEMA(lin2, halflife_Mimg=0.5, rampup_ratio=0.09)
    pf2.load_state_dict(st_pf)
    tr2.load_state_dict(st_tr)
    assert pf2.stds == [0.05, 0.1]
    assert all(torch.equal(a, b) for a, b in zip(pf2.emas, pf.emas)) and torch.equal(tr2.ema, tr.ema)
    lin2.load_state_dict(lin.state_dict())
    flat = torch.cat([p.detach().reshape(-1) for p in lin.parameters()]).contiguous()
    pf3 = PowerFunctionEMA(lin2, stds=[0.05, 0.1], flat_params=flat)        # the trainer's flat vector instead of torch.cat per step
    pf3.load_state_dict(st_pf)
    pf.update(cur_nimg=nimg + 64, batch_size=64)
    pf2.update(cur_nimg=nimg + 64, batch_size=64)
    pf3.update(cur_nimg=nimg + 64, batch_size=64)
    assert all(torch.equal(a, b) for a, b in zip(pf2.emas, pf.emas)) and all(torch.equal(a, b) for a, b in zip(pf3.emas, pf.emas))


def _adpm2_eps(g, B, L, steps):
    torch.manual_seed(int(g["seed"]))
    return torch.stack([torch.randn(B, 1, L) for _ in range(steps - 1)])       # one randn_like per step, in order


def test_oracle_adpm2_matches_reference():
    from oracle import edm, wavenet as owav
    from oracle.weights import make_wavenet_state_dict
    g = load_golden("adpm2_small")
    C, layers, cycle, B, L, seed, steps = (int(v) for v in g["cfg"])
    net_fn = owav.make_net_fn(make_wavenet_state_dict(C, layers, seed), cycle)
    den = lambda x, s: edm.denoise(x, net_fn, 0.2, sigma=float(s))          # noqa: E731
    it = iter(_adpm2_eps(g, B, L, steps))
    with torch.no_grad():
        out = edm.adpm2_sampler(torch.from_numpy(g["noise"]), den, torch.from_numpy(g["sigmas"]), steps, eps_fn=lambda x: next(it))
        out0 = edm.adpm2_sampler(torch.from_numpy(g["noise"]), den, torch.from_numpy(g["sigmas"]), steps, rho=7.0, eta=0.0)
    assert rel_l2(out, g["out"]) < 1e-5 and rel_l2(out0, g["out_rho7_eta0"]) < 1e-5


@pytest.mark.gpu
def test_adpm2_sampler_gpu_vs_reference_golden():
    from audiodiffuser_b200 import ADPM2Sampler, EluDiffusion, WaveNetNoise, _native as N
    from oracle.weights import make_wavenet_state_dict
    dev = torch.device("cuda:0")
    g = load_golden("adpm2_small")
    C, layers, cycle, B, L, seed, steps = (int(v) for v in g["cfg"])
    net = WaveNetNoise(C, layers, cycle, precision="fp32")
    net.load_state_dict(make_wavenet_state_dict(C, layers, seed), strict=True)
    net = net.to(dev)
    diff = EluDiffusion(0.2)
    noise, sig = torch.from_numpy(g["noise"]).to(dev), torch.from_numpy(g["sigmas"]).to(dev)
    smp = ADPM2Sampler(rho=1.0, num_steps=steps)
    out = smp(noise, fn=diff.denoise_fn, net=net, sigmas=sig, eps=_adpm2_eps(g, B, L, steps).to(dev))
    assert smp.last_nfe == 2 * (steps - 1)
    out0 = ADPM2Sampler(rho=7.0, num_steps=steps, eta=0.0)(noise, fn=diff.denoise_fn, net=net, sigmas=sig)
    N.check_async()
    e, e0 = rel_l2(out, g["out"]), rel_l2(out0, g["out_rho7_eta0"])
    print(f"ADPM2 fp32: rel-L2 {e:.3e} (ancestral noise replayed), {e0:.3e} (eta = 0, rho = 7)")
    assert e < 2e-5 and e0 < 2e-5
