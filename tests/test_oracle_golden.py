"""Pin the CPU oracle against outputs of the reference's own code (tests/golden, produced by
oracle/make_golden.py in the build container). CPU-only."""
import math

import numpy as np
import pytest
import torch

from oracle import edm, wavenet
from oracle.weights import make_wavenet_state_dict
from conftest import load_golden, rel_l2

TOL = 2e-6      # fp32 CPU vs fp32 CPU, different op order (functional vs nn.Module)


def test_scalar_kats():
    g = load_golden("kat_scalars")
    for key, vals in zip(g["keys"], g["values"]):
        sd_, s_ = (float(v) for v in str(key).split("_"))
        c_skip, c_out, c_in, c_noise = edm.scale_weights(torch.tensor([s_]), sd_, 3)
        got = [float(c_skip), float(c_out), float(c_in), float(c_noise),
               float(edm.loss_weight(torch.tensor([s_]), sd_))]
        np.testing.assert_allclose(got, vals, rtol=1e-6)
    # the closed-form values SURVEY.md §8(c) lists for sigma_data = sigma = 0.5
    c_skip, c_out, c_in, c_noise = edm.scale_weights(torch.tensor([0.5]), 0.5, 3)
    assert abs(float(c_skip) - 0.5) < 1e-7 and abs(float(c_out) - 0.5 / math.sqrt(2)) < 1e-7
    assert abs(float(c_in) - math.sqrt(2)) < 1e-6 and abs(float(c_noise) - 0.25 * math.log(0.5)) < 1e-7
    assert abs(float(edm.loss_weight(torch.tensor([0.5]), 0.5)) - 8.0) < 1e-5


def test_karras_schedule_bitexact():
    g = load_golden("kat_scalars")
    assert np.array_equal(edm.karras_schedule(0.002, 80.0, 7.0, 18).numpy(), g["karras18"])
    assert np.array_equal(edm.karras_schedule(0.002, 80.0, 7.0, 50).numpy(), g["karras50"])
    assert np.array_equal(edm.karras_schedule(0.01, 10.0, 3.0, 5).numpy(), g["karras5_rho3"])


@pytest.mark.parametrize("name", ["wavenet_c64_l4", "wavenet_c256_l3", "wavenet_c256_l13_dil2048",
                                  "wavenet_c256_l2_short"])
def test_wavenet_forward(name):
    g = load_golden(name)
    C, layers, cycle, B, L, seed = (int(v) for v in g["cfg"])
    sd = make_wavenet_state_dict(C, layers, seed)
    out = wavenet.wavenet_forward(sd, torch.from_numpy(g["audio"]), torch.from_numpy(g["t"]), cycle)
    assert out.shape == (B, 1, L)
    assert float(np.abs(g["out"]).max()) > 1e-3          # non-vacuous
    assert rel_l2(out, g["out"]) < TOL


def _small_net(g):
    C, layers, cycle, B, L, seed = (int(v) for v in g["cfg"][:6])
    return wavenet.make_net_fn(make_wavenet_state_dict(C, layers, seed), cycle)


def test_denoise():
    g = load_golden("denoise_small")
    net = _small_net(g)
    x = torch.from_numpy(g["x"])
    for s_ in [80.0, 10.0, 1.0, 0.1, 0.002]:
        out = edm.denoise(x * s_, net, 0.2, sigma=s_)
        assert rel_l2(out, g[f"sigma_{s_}"]) < TOL
        assert float(out.abs().max()) <= 1.0
    out = edm.denoise(x, net, 0.2, sigmas=torch.from_numpy(g["sigmas_per_sample"]), inference=False)
    assert rel_l2(out, g["per_sample"]) < TOL


def test_samplers():
    g = load_golden("sampler_small")
    net = _small_net(g)
    N = int(g["cfg"][6])
    noise, sig = torch.from_numpy(g["noise"]), torch.from_numpy(g["sigmas"])
    den = lambda x, s: edm.denoise(x, net, 0.2, sigma=float(s))
    tr = []
    assert rel_l2(edm.edm_sampler(noise, den, sig, N, trace=tr), g["heun"]) < 5e-6
    assert tr[-1] == int(g["nfe_heun"]) == 2 * N - 1
    assert rel_l2(edm.edm_sampler(noise, den, sig, N, use_heun=False, trace=tr), g["euler"]) < 5e-6
    assert tr[-1] == int(g["nfe_euler"]) == N
    torch.manual_seed(int(g["churn_seed"]))
    out = edm.edm_sampler(noise, den, sig, N, s_tmin=0.05, s_tmax=50.0, s_churn=2.0, s_noise=1.003, trace=tr)
    assert rel_l2(out, g["churn"]) < 5e-6 and tr[-1] == int(g["nfe_churn"])
    assert rel_l2(edm.edm_alpha_sampler(noise, den, sig, N, alpha=1.0, trace=tr), g["alpha1"]) < 5e-6
    assert tr[-1] == int(g["nfe_alpha1"]) == 2 * (N - 1)
    assert rel_l2(edm.edm_alpha_sampler(noise, den, sig, N, alpha=0.5, trace=tr), g["alpha05"]) < 5e-6
    assert int(g["nfe_heun18"]) == 35 and int(g["nfe_alpha18"]) == 34


def test_dsm_loss():
    g = load_golden("dsm_loss_small")
    net = _small_net(g)
    loss = edm.dsm_loss(torch.from_numpy(g["x"]), torch.from_numpy(g["noise"]),
                        torch.from_numpy(g["sigmas"]), net, 0.2)
    np.testing.assert_allclose(loss.numpy(), g["loss"], rtol=2e-5)


def test_full_size_single_call():
    """Full BASELINE-shape network (C=256, 36 layers, L=16000), one denoiser call at B=1."""
    import os
    from conftest import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, "full_diffwave_b1.npz")):
        pytest.skip("full-size golden not generated")
    g = load_golden("full_diffwave_b1")
    C, layers, cycle, B, L, seed = (int(v) for v in g["cfg"][:6])
    net = wavenet.make_net_fn(make_wavenet_state_dict(C, layers, seed), cycle)
    noise = torch.from_numpy(g["noise"])
    with torch.no_grad():
        out = edm.denoise(noise * 1.0, net, 0.2, sigma=1.0)
    assert rel_l2(out, g["den_sigma_1.0"]) < 5e-6
