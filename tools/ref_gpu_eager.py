"""Context number for BASELINE.json configs[4] ("vs reference PyTorch GPU ... paths"): the reference ALGORITHM (the oracle
restatement, which is plain torch code and bit-identical to the reference's modules on CPU) run with PyTorch eager on the
same B200 — fp32 and bf16 autocast — next to the fused path. Not part of bench.py; prints one JSON line per case.

    python tools/ref_gpu_eager.py [B]
"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import edm as oedm, wavenet as owav
from oracle.weights import make_wavenet_state_dict
from audiodiffuser_b200 import WaveNetNoise, EluDiffusion, _native

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
C, LAYERS, CYCLE, L = 256, 36, 12, 16000
dev = torch.device("cuda:0")
sd = make_wavenet_state_dict(C, LAYERS, seed=0)
sd_dev = {k: v.to(dev) for k, v in sd.items()}
x = torch.randn(B, 1, L, device=dev)
sig = torch.full((B,), 1.0, device=dev)


def timed(fn, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


net_fn = owav.make_net_fn(sd_dev, CYCLE)
with torch.no_grad():
    t_fp32 = timed(lambda: oedm.denoise(x, net_fn, 0.2, sigmas=sig))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        t_bf16 = timed(lambda: oedm.denoise(x, net_fn, 0.2, sigmas=sig))
net = WaveNetNoise(C, LAYERS, CYCLE, precision="bf16")
net.load_state_dict(sd, strict=True)
net = net.to(dev)
diff = EluDiffusion(0.2)
t_ours = timed(lambda: diff.denoise_fn(x, net=net, sigma=1.0, inference=True))
_native.check_async()
for name, t in (("torch eager fp32 (cuDNN)", t_fp32), ("torch eager bf16 autocast (cuDNN)", t_bf16), ("adb200 bf16 fused", t_ours)):
    print(json.dumps({"case": name, "batch": B, "ms_per_denoiser_call": 1e3 * t, "samples_per_s_heun18": B / (35 * t),
                      "tflops": 606.093e9 * B / t / 1e12}))

# ---- training step (BASELINE.json configs[2]: DSM loss forward + backward + AdamW), same box, same batch ----
if os.environ.get("ADB_EAGER_TRAIN", "1") != "0":
    from audiodiffuser_b200.training import FusedTrainer
    Bt = min(B, 16)
    xt = (torch.rand(Bt, 1, L, device=dev) * 2 - 1)
    st = torch.exp(-1.2 + 1.2 * torch.randn(Bt, device=dev))
    params = {k: v.clone().requires_grad_(True) for k, v in sd_dev.items()}
    opt = torch.optim.AdamW(list(params.values()), lr=1e-4, betas=(0.9, 0.999), weight_decay=0.01, fused=True)
    p_net = owav.make_net_fn(params, CYCLE)

    def eager_step(autocast):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss = oedm.dsm_loss(xt, torch.randn_like(xt), st, p_net, 0.2).mean()
        loss.backward()
        opt.step()

    t_tr32 = timed(lambda: eager_step(False), reps=2)
    t_tr16 = timed(lambda: eager_step(True), reps=2)
    del params, opt
    torch.cuda.empty_cache()
    tnet = WaveNetNoise(C, LAYERS, CYCLE, precision="bf16")
    tnet.load_state_dict(sd, strict=True)
    trainer = FusedTrainer(tnet.to(dev), EluDiffusion(0.2), lr=1e-4)
    t_trours = timed(lambda: trainer.step(xt, st), reps=3)
    _native.check_async()
    for name, t in (("torch eager fp32 autograd + fused AdamW", t_tr32), ("torch eager bf16 autocast autograd + fused AdamW", t_tr16),
                    ("adb200 bf16 fused training step", t_trours)):
        print(json.dumps({"case": name, "batch": Bt, "ms_per_train_step": 1e3 * t, "train_samples_per_s": Bt / t,
                          "tflops_3x_forward": 3 * 606.093e9 * Bt / t / 1e12}))
