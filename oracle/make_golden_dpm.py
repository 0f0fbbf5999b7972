"""Goldens for DPMSampler (sampler_edm.py:495-805) and UniPCSampler (:807-1053), from the REFERENCE's own classes
(build container only; /root/reference is not read at test time).

    python -m oracle.make_golden_dpm

One small DiffWave denoiser, one noise tensor; every case stores only the final waveforms and the number of denoiser calls.
UniPC runs on a 4-D state [B,1,1,L] (its einsum `k,bkchw->bchw` accepts nothing else) through an adapter that flattens it.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import import_reference, WaveNetAdapter      # noqa: E402
from oracle.make_golden import build_ref_net, save, seeded           # noqa: E402

# (name, class, constructor kwargs, schedule points). Schedules: Karras(0.01, 20, rho 5, points).
DPM_CASES = [
    ("ms1_x0_log", dict(order=1, num_steps=6, multisteps=True, x0_pred=True, log_time_spacing=True), 7),
    ("ms2_x0_log", dict(order=2, num_steps=7, multisteps=True, x0_pred=True, log_time_spacing=True), 9),
    ("ms3_x0_log", dict(order=3, num_steps=8, multisteps=True, x0_pred=True, log_time_spacing=True), 9),
    ("ms3_eps_log", dict(order=3, num_steps=8, multisteps=True, x0_pred=False, log_time_spacing=True), 9),
    ("ms2_x0_sig", dict(order=2, num_steps=8, multisteps=True, x0_pred=True, log_time_spacing=False), 8),
    ("ms3_eps_sig", dict(order=3, num_steps=8, multisteps=True, x0_pred=False, log_time_spacing=False), 8),
    ("ss3_x0_log", dict(order=3, num_steps=9, multisteps=False, x0_pred=True, log_time_spacing=True), 9),
    ("ss3_eps_log", dict(order=3, num_steps=8, multisteps=False, x0_pred=False, log_time_spacing=True), 9),
    ("ss2_x0_log", dict(order=2, num_steps=7, multisteps=False, x0_pred=True, log_time_spacing=True), 9),
    ("ss1_eps_log", dict(order=1, num_steps=5, multisteps=False, x0_pred=False, log_time_spacing=True), 9),
    ("ss2_x0_sig", dict(order=2, num_steps=8, multisteps=False, x0_pred=True, log_time_spacing=False), 8),
]
UNIPC_CASES = [
    ("pc1_x0_log", dict(num_steps=6, order=1, x0_pred=True, log_time_spacing=True), 7),
    ("pc2_x0_log", dict(num_steps=7, order=2, x0_pred=True, log_time_spacing=True), 9),
    ("pc3_x0_log", dict(num_steps=8, order=3, x0_pred=True, log_time_spacing=True), 9),
    ("pc3_eps_log", dict(num_steps=8, order=3, x0_pred=False, log_time_spacing=True), 9),
    ("pc2_x0_sig", dict(num_steps=8, order=2, x0_pred=True, log_time_spacing=False), 8),
]
NET = dict(C=64, layers=4, cycle=2, seed=501, B=2, L=256)


def schedule(ref, points):
    return ref.scheduler.KarrasSchedule(0.01, 20.0, 5.0, points)()


class Flat4d:
    """net(x [B,1,1,L], t) for the 4-D state UniPC needs."""

    def __init__(self, net):
        self.net = net

    def __call__(self, x, t, **kw):
        return self.net(x.reshape(x.shape[0], x.shape[-1]), t).reshape(x.shape)


@torch.no_grad()
def main():
    ref = import_reference()
    net = build_ref_net(ref, NET["C"], NET["layers"], NET["cycle"], NET["seed"])
    diff = ref.diffusion.EluDiffusion(sigma_data=0.2)
    noise = seeded((NET["B"], 1, NET["L"]), NET["seed"] + 1)
    arrays = dict(noise=noise, cfg=np.array([NET[k] for k in ("C", "layers", "cycle", "B", "L", "seed")], dtype=np.int64))
    calls = []

    def counting_fn(*a, **k):
        calls.append(1)
        return diff.denoise_fn(*a, **k)

    for name, kw, points in DPM_CASES:
        calls.clear()
        out = ref.sampler_edm.DPMSampler(cond_scale=1.0, **kw)(noise, fn=counting_fn, net=WaveNetAdapter(net),
                                                               sigmas=schedule(ref, points))
        arrays["dpm_" + name] = out
        arrays["nfe_dpm_" + name] = np.int64(len(calls))
        print(name, len(calls), float(out.abs().mean()))
    for name, kw, points in UNIPC_CASES:
        calls.clear()
        out = ref.sampler_edm.UniPCSampler(cond_scale=1.0, **kw)(noise[:, :, None, :], fn=counting_fn, net=Flat4d(net),
                                                                 sigmas=schedule(ref, points))
        arrays["unipc_" + name] = out[:, :, 0, :]
        arrays["nfe_unipc_" + name] = np.int64(len(calls))
        print(name, len(calls), float(out.abs().mean()))
    save("dpm_unipc_small", **arrays)


if __name__ == "__main__":
    main()
