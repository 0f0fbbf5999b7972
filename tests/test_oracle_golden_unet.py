"""CPU: the UNet1d oracle restatement against goldens produced by the reference's own UNet1dBase
(oracle/make_golden_unet.py). Bit-exact on the same torch build; 1e-6 relative otherwise."""
import pytest
import torch

from conftest import load_golden, rel_l2


@pytest.mark.parametrize("name", ["unet1d_small", "unet1d_small_ragged", "unet1d_mid"])
def test_unet_oracle_matches_reference_golden(name):
    from oracle import unet1d as ou
    from oracle.weights import UNET_CASES, make_unet1d_state_dict
    cfg, B, L, seed = UNET_CASES[name]
    g = load_golden(name)
    assert tuple(int(v) for v in g["cfg"]) == (B, L, seed)
    sd = make_unet1d_state_dict(cfg, seed)
    with torch.no_grad():
        out = ou.unet1d_forward(sd, cfg, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]))
    assert out.shape == (B, cfg["in_channels"], L)
    assert rel_l2(out, g["out"]) < 1e-6


def test_unet_param_count_config4():
    """102.25 M parameters for the audio-diffusion-pytorch default configuration."""
    from oracle.weights import UNET1D_CONFIG4, unet1d_param_shapes
    import math
    n = sum(math.prod(s) for s in unet1d_param_shapes(UNET1D_CONFIG4).values())
    assert n == 102_245_568


def test_unet_edm_denoise_and_sampler_golden():
    from oracle import edm, unet1d as ou
    from oracle.weights import UNET_MID, make_unet1d_state_dict
    g = load_golden("unet1d_mid_edm")
    B, L, seed, steps = (int(v) for v in g["cfg"])
    net_fn = ou.make_net_fn(make_unet1d_state_dict(UNET_MID, seed), UNET_MID)
    noise = torch.from_numpy(g["noise"])
    with torch.no_grad():
        for s in (80.0, 1.0, 0.002):
            assert rel_l2(edm.denoise(noise * s, net_fn, 0.2, sigma=s), g[f"den_sigma_{s}"]) < 1e-6
        x = edm.edm_sampler(noise, lambda x_, s_: edm.denoise(x_, net_fn, 0.2, sigma=float(s_)), torch.from_numpy(g["sigmas"]), steps)
    assert rel_l2(x, g["heun"]) < 1e-5


def test_unet_class_conditioning_oracle_golden():
    from oracle import edm, unet1d as ou
    from oracle.weights import UNET_CLASS, make_unet1d_state_dict
    g = load_golden("unet1d_class_cfg")
    B, L, seed = (int(v) for v in g["cfg"])
    sd = make_unet1d_state_dict(UNET_CLASS, seed)
    x, t, cls = (torch.from_numpy(g[k]) for k in ("x", "t", "classes"))
    with torch.no_grad():
        assert rel_l2(ou.unet1d_forward(sd, UNET_CLASS, x, t, cls, 0.0), g["f_cond"]) < 1e-6
        assert rel_l2(ou.unet1d_forward(sd, UNET_CLASS, x, t, cls, 1.0), g["f_null"]) < 1e-6
        den = edm.denoise(x * 0.5, ou.make_net_fn(sd, UNET_CLASS), 0.2, sigma=0.5, cond_scale=2.5, classes=cls)
    assert rel_l2(den, g["den_cfg_sigma_0.5"]) < 1e-6
