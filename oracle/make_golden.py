"""Generate tests/golden/*.npz by running the REFERENCE's own code (build container only).

    python -m oracle.make_golden [--full]

Inputs and weights come from seeded CPU generators (oracle/weights.py), are loaded into the
reference's `WaveNetNoise` with `load_state_dict(strict=True)`, and the reference's
`EluDiffusion`, `EDMSampler`, `EDMAlphaSampler`, `KarrasSchedule` are run unmodified on them
(with the 4-line adapter of SURVEY.md §8(c)).  The outputs are the pins the oracle — and through
it the CUDA path — is checked against.  `--full` adds the full-size (C=256, 36 layers, L=16000)
cases, which take a few minutes of CPU.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import import_reference, WaveNetAdapter      # noqa: E402
from oracle.weights import make_wavenet_state_dict, wavenet_param_shapes  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def seeded(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(shape, generator=g, dtype=torch.float32) * scale


def build_ref_net(ref, C, layers, cycle, seed):
    net = ref.wavenet.WaveNetNoise(residual_channels=C, residual_layers=layers, dilation_cycle=cycle)
    ref_shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    mine = {k: tuple(v) for k, v in wavenet_param_shapes(C, layers).items()}
    assert ref_shapes == mine, "weight factory key/shape set differs from the reference"
    assert list(ref_shapes.keys()) == list(mine.keys())
    net.load_state_dict(make_wavenet_state_dict(C, layers, seed), strict=True)
    net.eval()
    return net


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


@torch.no_grad()
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    ref = import_reference()

    # ---- scalar known answers -------------------------------------------------------------
    kat = {}
    for sd_, s_ in [(0.5, 0.5), (0.2, 80.0), (0.2, 0.002), (0.2, 1.0)]:
        d = ref.diffusion.EluDiffusion(sigma_data=sd_)
        sig = torch.tensor([s_], dtype=torch.float32)
        c_skip, c_out, c_in, c_noise = d.get_scale_weights(sig, 3)
        kat[f"{sd_}_{s_}"] = [float(c_skip), float(c_out), float(c_in), float(c_noise),
                              float(d.loss_weight(sig))]
    save("kat_scalars",
         keys=np.array(list(kat.keys())), values=np.array(list(kat.values()), dtype=np.float64),
         karras18=ref.scheduler.KarrasSchedule(0.002, 80.0, 7.0, 18)(),
         karras50=ref.scheduler.KarrasSchedule(0.002, 80.0, 7.0, 50)(),
         karras5_rho3=ref.scheduler.KarrasSchedule(0.01, 10.0, 3.0, 5)())

    # ---- backbone forward, small / ragged / large-dilation ---------------------------------
    cases = [
        # name, C, layers, cycle, B, L, seed
        ("wavenet_c64_l4", 64, 4, 2, 2, 300, 11),
        ("wavenet_c256_l3", 256, 3, 12, 2, 1000, 12),
        ("wavenet_c256_l13_dil2048", 256, 13, 12, 1, 4500, 13),
        ("wavenet_c256_l2_short", 256, 2, 12, 3, 77, 14),
    ]
    for name, C, layers, cycle, B, L, seed in cases:
        net = build_ref_net(ref, C, layers, cycle, seed)
        audio = seeded((B, L), seed + 1000)
        t = seeded((B,), seed + 2000, 1.5)
        out = net(audio, t)
        save(name, audio=audio, t=t, out=out,
             cfg=np.array([C, layers, cycle, B, L, seed], dtype=np.int64))

    # ---- denoiser, samplers, loss on a small net -------------------------------------------
    C, layers, cycle, seed, B, L = 64, 4, 2, 21, 2, 256
    net = build_ref_net(ref, C, layers, cycle, seed)
    adapter = WaveNetAdapter(net)
    diff = ref.diffusion.EluDiffusion(sigma_data=0.2)
    x = seeded((B, 1, L), 31)
    den = {}
    for s_ in [80.0, 10.0, 1.0, 0.1, 0.002]:
        den[f"sigma_{s_}"] = diff.denoise_fn(x * s_, net=adapter, sigma=torch.tensor(s_), inference=True)
    per_sample = torch.tensor([3.0, 0.05])
    den["per_sample"] = diff.denoise_fn(x, net=adapter, sigmas=per_sample, inference=False)
    save("denoise_small", x=x, sigmas_per_sample=per_sample,
         cfg=np.array([C, layers, cycle, B, L, seed], dtype=np.int64), **den)

    N = 6
    sig = ref.scheduler.KarrasSchedule(0.002, 80.0, 7.0, N)()
    noise = seeded((B, 1, L), 41)
    calls = []

    def counting_fn(*a, **k):
        calls.append(1)
        return diff.denoise_fn(*a, **k)

    res = {}
    s1 = ref.sampler_edm.EDMSampler(s_tmin=0, s_tmax=float("inf"), s_churn=0.0, s_noise=1.0,
                                    num_steps=N, cond_scale=1.0, use_heun=True)
    res["heun"] = s1(noise, fn=counting_fn, net=adapter, sigmas=sig); res["nfe_heun"] = len(calls); calls.clear()
    s2 = ref.sampler_edm.EDMSampler(s_churn=0.0, s_noise=1.0, num_steps=N, use_heun=False)
    res["euler"] = s2(noise, fn=counting_fn, net=adapter, sigmas=sig); res["nfe_euler"] = len(calls); calls.clear()
    s3 = ref.sampler_edm.EDMSampler(s_tmin=0.05, s_tmax=50.0, s_churn=2.0, s_noise=1.003,
                                    num_steps=N, use_heun=True)
    torch.manual_seed(777)
    res["churn"] = s3(noise, fn=counting_fn, net=adapter, sigmas=sig); res["nfe_churn"] = len(calls); calls.clear()
    s4 = ref.sampler_edm.EDMAlphaSampler(alpha=1.0, num_steps=N, use_heun=True)
    res["alpha1"] = s4(noise, fn=counting_fn, net=adapter, sigmas=sig); res["nfe_alpha1"] = len(calls); calls.clear()
    s5 = ref.sampler_edm.EDMAlphaSampler(alpha=0.5, num_steps=N, use_heun=True)
    res["alpha05"] = s5(noise, fn=counting_fn, net=adapter, sigmas=sig); res["nfe_alpha05"] = len(calls); calls.clear()
    # NFE counts at the BASELINE step count (cheap: a zero net)
    zero = lambda x_, t_, **k: torch.zeros_like(x_)
    sig18 = ref.scheduler.KarrasSchedule(0.002, 80.0, 7.0, 18)()
    ref.sampler_edm.EDMSampler(s_churn=0.0, s_noise=1.0, num_steps=18)(noise, fn=counting_fn, net=zero, sigmas=sig18)
    res["nfe_heun18"] = len(calls); calls.clear()
    ref.sampler_edm.EDMAlphaSampler(num_steps=18)(noise, fn=counting_fn, net=zero, sigmas=sig18)
    res["nfe_alpha18"] = len(calls); calls.clear()
    save("sampler_small", noise=noise, sigmas=sig, churn_seed=np.int64(777),
         cfg=np.array([C, layers, cycle, B, L, seed, N], dtype=np.int64), **res)

    # DSM loss: Diffusion.forward draws randn_like(x) internally (diffusion.py:76)
    x0 = seeded((B, 1, L), 51, 0.2).clamp(-1, 1)
    sig_b = torch.tensor([0.7, 0.03])
    torch.manual_seed(888)
    loss = diff(x0, adapter, sigmas=sig_b)
    torch.manual_seed(888)
    loss_noise = torch.randn_like(x0)
    save("dsm_loss_small", x=x0, sigmas=sig_b, noise=loss_noise, loss=loss, seed=np.int64(888),
         cfg=np.array([C, layers, cycle, B, L, seed], dtype=np.int64))

    # ---- full-size DiffWave (BASELINE config shape, B=1) -----------------------------------
    if args.full:
        C, layers, cycle, seed, B, L = 256, 36, 12, 0, 1, 16000
        net = build_ref_net(ref, C, layers, cycle, seed)
        adapter = WaveNetAdapter(net)
        noise = seeded((B, 1, L), 61)
        outs = {}
        for s_ in [80.0, 1.0, 0.002]:
            outs[f"net_sigma_{s_}"] = net((noise[:, 0] * s_) * float((s_ ** 2 + 0.04) ** -0.5),
                                          torch.full((B,), 0.25 * float(np.log(s_))))
            outs[f"den_sigma_{s_}"] = diff.denoise_fn(noise * s_, net=adapter, sigma=torch.tensor(s_), inference=True)
        # The same algorithm evaluated in fp64 (the reference modules themselves cannot run in double:
        # diffusion_embedding, wavenet.py:88-92, builds an fp32 table). The oracle restatement, which
        # reproduces the reference's fp32 outputs bit-for-bit (asserted here), is run in fp64 instead.
        # It shows how far the reference's own fp32 output is from exact arithmetic (1.0e-5 relative at
        # sigma = 80), which bounds what any fp32-vs-fp32 comparison can show.
        from oracle import edm as oedm, wavenet as owav
        sd32 = make_wavenet_state_dict(C, layers, seed)
        sd64 = {k: v.double() for k, v in sd32.items()}
        for s_ in [80.0, 1.0, 0.002]:
            o32 = oedm.denoise(noise * s_, owav.make_net_fn(sd32, cycle), 0.2, sigma=s_)
            assert torch.equal(o32, outs[f"den_sigma_{s_}"]), "oracle no longer bit-identical to the reference"
            outs[f"den64_sigma_{s_}"] = oedm.denoise(noise.double() * s_, owav.make_net_fn(sd64, cycle), 0.2, sigma=s_)
        sampler = ref.sampler_edm.EDMSampler(s_tmin=0, s_tmax=float("inf"), s_churn=0.0, s_noise=1.0,
                                             num_steps=18, cond_scale=1.0, use_heun=True)
        outs["heun18"] = sampler(noise, fn=diff.denoise_fn, net=adapter, sigmas=sig18)
        save("full_diffwave_b1", noise=noise, sigmas=sig18,
             cfg=np.array([C, layers, cycle, B, L, seed, 18], dtype=np.int64), **outs)

    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "torch": torch.__version__,
                   "reference": "AgentCooper2002/AudioDiffuser @ /root/reference (unmodified)",
                   "sigma_data": 0.2, "full": bool(args.full) or os.path.exists(os.path.join(OUT, "full_diffwave_b1.npz"))},
                  f, indent=1)


if __name__ == "__main__":
    main()
