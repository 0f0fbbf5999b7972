"""Generate tests/golden/unet1d_*.npz by running the REFERENCE's own UNet1dBase / EluDiffusion / EDMSampler
(build container only; see oracle/make_golden.py for the conventions).

    python -m oracle.make_golden_unet [case ...]
"""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import import_reference                      # noqa: E402
from oracle.make_golden import save, seeded                          # noqa: E402
from oracle.weights import (make_unet1d_state_dict, unet1d_param_shapes, UNET1D_CONFIG4, UNET_SMALL, UNET_MID,   # noqa: E402
                            UNET_CLASS, UNET_CASES as CASES)
from oracle import unet1d as ounet                                   # noqa: E402

def build_ref(R, cfg, seed):
    net = R.UNet1dBase(**cfg).eval()
    ref_shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    mine = dict(unet1d_param_shapes(cfg))
    assert list(ref_shapes.keys()) == list(mine.keys()) and ref_shapes == mine, "weight factory differs from the reference"
    sd = make_unet1d_state_dict(cfg, seed)
    net.load_state_dict(sd, strict=True)
    return net, sd


@torch.no_grad()
def main():
    torch.set_num_threads(os.cpu_count())
    ref = import_reference()
    R = importlib.import_module("src.models.backbones.unet1d")
    only = [a for a in sys.argv[1:] if not a.startswith("-")]           # optional: regenerate just the named cases
    for name, (cfg, B, L, seed) in CASES.items():
        if only and name not in only:
            continue
        net, sd = build_ref(R, cfg, seed)
        x = seeded((B, cfg["in_channels"], L), seed + 1000)
        t = seeded((B,), seed + 2000, 1.5)
        out = net(x, t)
        assert torch.equal(ounet.unet1d_forward(sd, cfg, x, t), out), "oracle no longer bit-identical to the reference"
        save(name, x=x, t=t, out=out, cfg=np.array([B, L, seed], dtype=np.int64))

    if only:
        return
    # denoiser + sampler through the reference's own EluDiffusion / EDMSampler (no adapter needed)
    cfg, B, L, seed, steps = UNET_MID, 2, 4096, 105, 5
    net, sd = build_ref(R, cfg, seed)
    diff = ref.diffusion.EluDiffusion(sigma_data=0.2)
    noise = seeded((B, 2, L), seed + 1000)
    res = {}
    for s_ in (80.0, 1.0, 0.002):
        res[f"den_sigma_{s_}"] = diff.denoise_fn(noise * s_, net=net, sigma=torch.tensor(s_), inference=True)
    sig = ref.scheduler.KarrasSchedule(0.002, 80.0, 7.0, steps)()
    smp = ref.sampler_edm.EDMSampler(s_tmin=0, s_tmax=float("inf"), s_churn=0.0, s_noise=1.0, num_steps=steps, cond_scale=1.0,
                                     use_heun=True)
    res["heun"] = smp(noise, fn=diff.denoise_fn, net=net, sigmas=sig)
    save("unet1d_mid_edm", noise=noise, sigmas=sig, cfg=np.array([B, L, seed, steps], dtype=np.int64), **res)

    # label conditioning + classifier-free guidance (conditioner.py:59-111, diffusion.py:50-54)
    cfg, B, L, seed = UNET_CLASS, 2, 4096, 106
    net, sd = build_ref(R, cfg, seed)
    x = seeded((B, 2, L), seed + 1000)
    t = seeded((B,), seed + 2000, 1.5)
    cls = torch.tensor([3, 7])
    res = {"f_cond": net(x, t, classes=cls, cond_drop_prob=0.0), "f_null": net(x, t, classes=cls, cond_drop_prob=1.0)}
    for s_ in (10.0, 0.5):
        res[f"den_cfg_sigma_{s_}"] = diff.denoise_fn(x * s_, net=net, sigma=torch.tensor(s_), inference=True, cond_scale=2.5, classes=cls)
    save("unet1d_class_cfg", x=x, t=t, classes=cls, cfg=np.array([B, L, seed], dtype=np.int64), **res)


if __name__ == "__main__":
    main()
