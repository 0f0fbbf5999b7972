"""BASELINE.json configs[4], "vs reference PyTorch GPU": the reference's OWN classes (WaveNetNoise + EluDiffusion + EDMSampler,
imported unmodified from baseline/_ref or /root/reference through oracle/ref_loader.py) run by PyTorch on the same B200 — the
full 18-step EDM-Heun trajectory of the headline workload — with the switches a user of the reference would sensibly turn on:
TF32 matmul / cuDNN, cudnn.benchmark, and either fp32 or bf16 autocast (Lightning's bf16-mixed). One JSON line per case.
Not part of bench.py (the reference arm there is the host-CPU path the tier asks for); context for the sweep table.

    python tools/ref_gpu.py [batch ...]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader                                   # noqa: E402
from oracle.weights import make_wavenet_state_dict              # noqa: E402

C, LAYERS, CYCLE, L, STEPS, SIGMA_DATA = 256, 36, 12, 16000, 18, 0.2
batches = [int(a) for a in sys.argv[1:]] or [64, 256]
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True

ref = ref_loader.import_reference()
net = ref.wavenet.WaveNetNoise(residual_channels=C, residual_layers=LAYERS, dilation_cycle=CYCLE)
net.load_state_dict(make_wavenet_state_dict(C, LAYERS, seed=0), strict=True)
net = net.to(dev).eval()
adapter = ref_loader.WaveNetAdapter(net)
diff = ref.diffusion.EluDiffusion(sigma_data=SIGMA_DATA)
sampler = ref.sampler_edm.EDMSampler(s_tmin=0, s_tmax=float("inf"), s_churn=0.0, s_noise=1.0, num_steps=STEPS, cond_scale=1.0, use_heun=True)
sigmas = ref.scheduler.KarrasSchedule(sigma_min=0.002, sigma_max=80.0, rho=7.0, num_steps=STEPS)().to(dev)

for B in batches:
    noise = torch.randn(B, 1, L, device=dev)
    for name, ctx in (("fp32 (TF32 on)", torch.autocast("cuda", enabled=False)), ("bf16 autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        with torch.no_grad(), ctx:
            one = diff.denoise_fn(noise, net=adapter, sigma=1.0, inference=True)          # warm-up: cuDNN autotune
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = sampler(noise, fn=diff.denoise_fn, net=adapter, sigmas=sigmas)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        assert torch.isfinite(out).all()
        print(json.dumps({"impl": "reference classes on the GPU (PyTorch eager, cuDNN/cuBLAS)", "precision": name, "batch": B,
                          "seconds_per_trajectory": dt, "samples_per_s": B / dt, "tflops": 606.093e9 * 35 * B / dt / 1e12,
                          "reference_root": ref_loader.REF_ROOT}), flush=True)
