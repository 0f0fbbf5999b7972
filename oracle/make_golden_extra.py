"""Goldens for the SURVEY §8(f) widening items, from the REFERENCE's own code (build container only):
DPM2MSampler (sampler_edm.py:1056-1131) and the EMA classes (src/models/phema.py).

    python -m oracle.make_golden_extra
"""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import import_reference, WaveNetAdapter      # noqa: E402
from oracle.make_golden import build_ref_net, save, seeded           # noqa: E402


@torch.no_grad()
def main():
    ref = import_reference()
    # ---- DPM-Solver++(2M) on a small DiffWave, schedule = Karras + 0 appended ----
    C, layers, cycle, seed, B, L, N = 64, 4, 2, 401, 2, 256, 8
    net = build_ref_net(ref, C, layers, cycle, seed)
    adapter = WaveNetAdapter(net)
    diff = ref.diffusion.EluDiffusion(sigma_data=0.2)
    noise = seeded((B, 1, L), seed + 1)
    sig = torch.cat([ref.scheduler.KarrasSchedule(0.002, 80.0, 7.0, N)(), torch.zeros(1)])
    calls = []

    def counting_fn(*a, **k):
        calls.append(1)
        return diff.denoise_fn(*a, **k)

    out = ref.sampler_edm.DPM2MSampler(num_steps=N)(noise, fn=counting_fn, net=adapter, sigmas=sig)
    sig2 = ref.scheduler.KarrasSchedule(0.01, 20.0, 5.0, N + 1)()                 # no zero at the end: all steps 2nd order
    out2 = ref.sampler_edm.DPM2MSampler(num_steps=N)(noise, fn=diff.denoise_fn, net=adapter, sigmas=sig2)
    save("dpm2m_small", noise=noise, sigmas=sig, out=out, nfe=np.int64(len(calls)), sigmas2=sig2, out2=out2,
         cfg=np.array([C, layers, cycle, B, L, seed, N], dtype=np.int64))

    # ---- ancestral DPM-Solver-2 (the reference's default sampler); noise replayed from the recorded seed ----
    N2 = 6
    sig3 = ref.scheduler.KarrasSchedule(0.002, 80.0, 7.0, N2)()
    torch.manual_seed(909)
    out3 = ref.stochastic_sampler_edm.ADPM2Sampler(rho=1.0, num_steps=N2)(noise, fn=diff.denoise_fn, net=adapter, sigmas=sig3)
    out3b = ref.stochastic_sampler_edm.ADPM2Sampler(rho=7.0, num_steps=N2, eta=0.0)(noise, fn=diff.denoise_fn, net=adapter, sigmas=sig3)
    save("adpm2_small", noise=noise, sigmas=sig3, out=out3, out_rho7_eta0=out3b, seed=np.int64(909),
         cfg=np.array([C, layers, cycle, B, L, seed, N2], dtype=np.int64))

    # ---- EMA classes on a tiny module ----
    phema = importlib.import_module("src.models.phema")
    torch.manual_seed(5)
    lin = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    p0 = torch.cat([p.detach().reshape(-1) for p in lin.parameters()]).clone()
    pf = phema.PowerFunctionEMA(lin, stds=[0.05, 0.1])
    tr = phema.TraditionalEMA(lin, halflife_Mimg=0.5, rampup_ratio=0.09)
    steps = []
    g = torch.Generator().manual_seed(6)
    nimg = 0
    for _ in range(5):
        for p in lin.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.1)
        nimg += 64
        pf.update(cur_nimg=nimg, batch_size=64)
        tr.update(cur_nimg=nimg, batch_size=64)
        steps.append(torch.cat([p.detach().reshape(-1) for p in lin.parameters()]).clone())
    flat = lambda m: torch.cat([p.detach().reshape(-1) for p in m.parameters()])      # noqa: E731
    save("ema_small", p0=p0, params=torch.stack(steps), pf0=flat(pf.emas[0]), pf1=flat(pf.emas[1]), trad=flat(tr.ema))


if __name__ == "__main__":
    main()
