"""Generate tests/golden/train_*.npz: DSM loss and parameter gradients from the REFERENCE's own modules under
PyTorch autograd (build container only).

    python -m oracle.make_golden_train

loss = EluDiffusion(sigma_data)(x, adapter(WaveNetNoise), sigmas=...)   (diffusion.py:65-97; the noise it draws with
randn_like is replayed from the recorded seed), total = loss.mean(), total.backward(). Gradients are stored as the
flat vector in state_dict order, sub-sampled with a fixed stride (the full vector of the C=256 case is 9 MB), plus
the exact L2 norm of every parameter's gradient.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import import_reference, WaveNetAdapter      # noqa: E402
from oracle.make_golden import build_ref_net, save, seeded           # noqa: E402

TRAIN_CASES = {   # name: (C, layers, cycle, B, L, seed, stride)
    "train_c64_l4": (64, 4, 2, 2, 256, 301, 3),
    "train_c256_l3": (256, 3, 12, 2, 600, 302, 16),
    "train_c256_l13_dil": (256, 13, 12, 1, 4500, 303, 64),     # dilation 1..2048 and back to 1: padding on both sides
}


def main():
    torch.set_num_threads(os.cpu_count())
    ref = import_reference()
    for name, (C, layers, cycle, B, L, seed, stride) in TRAIN_CASES.items():
        net = build_ref_net(ref, C, layers, cycle, seed)
        net.train()
        adapter = WaveNetAdapter(net)
        diff = ref.diffusion.EluDiffusion(sigma_data=0.2)
        x0 = seeded((B, 1, L), seed + 1000, 0.2).clamp(-1, 1)
        sig = (seeded((B,), seed + 2000) * 1.2 - 1.2).exp()            # LogNormal(-1.2, 1.2), sc09 experiment values
        torch.manual_seed(seed + 3000)
        loss = diff(x0, adapter, sigmas=sig)
        torch.manual_seed(seed + 3000)
        noise = torch.randn_like(x0)
        loss.mean().backward()
        grads = [p.grad.detach().reshape(-1) for p in net.state_dict(keep_vars=True).values()]
        flat = torch.cat(grads)
        norms = np.array([float(g.double().norm()) for g in grads])
        # the same computation in fp64 through the oracle restatement (whose fp32 autograd matches the reference's to
        # round-off, asserted here): the reference's own fp32 gradient is up to 6e-4 away from it on the 13-layer case,
        # which bounds what an fp32-vs-fp32 comparison can show
        from oracle import edm as oedm, wavenet as owav
        from oracle.weights import make_wavenet_state_dict
        flats = {}
        for dt in (torch.float32, torch.float64):
            sd = {k: v.to(dt).clone().requires_grad_(True) for k, v in make_wavenet_state_dict(C, layers, seed).items()}
            l64 = oedm.dsm_loss(x0.to(dt), noise.to(dt), sig.to(dt), owav.make_net_fn(sd, cycle), 0.2)
            l64.mean().backward()
            flats[dt] = torch.cat([v.grad.reshape(-1) for v in sd.values()])
        err32 = float((flats[torch.float32] - flat).norm() / flat.norm())
        assert err32 < 1e-5, ("oracle autograd no longer matches the reference", err32)
        grads64 = flats[torch.float64]
        norms64, o = [], 0
        for g_ in grads:
            norms64.append(float(grads64[o:o + g_.numel()].norm())); o += g_.numel()
        print(name, "reference fp32 vs fp64 evaluation:", float((grads64 - flat.double()).norm() / grads64.norm()))
        save(name, x=x0, sigmas=sig, noise=noise, loss=loss.detach(), grad_sub=flat[::stride].clone(), grad_norms=norms,
             grad64_sub=grads64[::stride].float().clone(), grad64_norms=np.array(norms64),
             cfg=np.array([C, layers, cycle, B, L, seed, stride], dtype=np.int64))


if __name__ == "__main__":
    main()
