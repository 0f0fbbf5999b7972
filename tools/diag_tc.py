"""GPU diagnostic: layer-by-layer comparison of the tcgen05 path against the fp32 CUDA path and the
CPU oracle, with error-pattern summaries that localise descriptor / layout mistakes. Writes a report
to gpurun_out/diag_tc.txt. Test infrastructure (may import oracle/)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiodiffuser_b200 import WaveNetNoise, _native          # noqa: E402
from oracle import wavenet as owav                            # noqa: E402
from oracle.weights import make_wavenet_state_dict            # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
lines = []


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    lines.append(s)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def pattern(name, got, want):
    """got/want: [B, L, C]. Summarise where the error lives."""
    err = (got - want).abs()
    scale = want.abs().mean().item() + 1e-12
    P(f"  {name}: rel-L2={rel(got, want):.3e} max-abs={err.max().item():.3e} mean|want|={scale:.3e} "
      f"nan={int(torch.isnan(got).sum())}")
    B, L, C = got.shape
    e_c = err.mean(dim=(0, 1)) / scale                  # per channel
    e_t = err.mean(dim=(0, 2)) / scale                  # per time
    P("    by 32-channel group:", np.array2string(e_c.view(-1, 32).mean(1).cpu().numpy(), precision=3))
    tt = e_t[: (L // 128) * 128].view(-1, 128) if L >= 128 else e_t.view(1, -1)
    P("    by row-in-tile (16-row groups):", np.array2string(tt.mean(0).view(-1, 16).mean(1).cpu().numpy(), precision=3))
    P("    by tile:", np.array2string(tt.mean(1).cpu().numpy()[:16], precision=3))


def main():
    dev = torch.device("cuda:0")
    P("device:", torch.cuda.get_device_name(0), "lib version", _native.lib().adb_version())
    C, layers, cycle, B, L = 256, 3, 12, 2, 1000
    sd = make_wavenet_state_dict(C, layers, seed=12)
    g = torch.Generator().manual_seed(1012)
    audio = torch.randn(B, L, generator=g)
    t = torch.randn(B, generator=g) * 1.5
    want, inter = owav.wavenet_forward(sd, audio, t, cycle, return_intermediates=True)

    nets = {}
    for prec in ("fp32", "bf16"):
        net = WaveNetNoise(C, layers, cycle, precision=prec)
        net.load_state_dict(sd, strict=True)
        nets[prec] = net.to(dev)

    res = {}
    for prec in ("fp32", "bf16"):
        P(f"== {prec} path, C={C} layers={layers} B={B} L={L}")
        try:
            out, dh, ds = nets[prec].forward_debug(audio.to(dev), t.to(dev), layers)
            _native.check_async()
        except Exception as e:                                 # noqa: BLE001
            P("  FAILED:", repr(e))
            continue
        res[prec] = (out.cpu(), dh.cpu(), ds.cpu())
        for n in range(layers):
            pattern(f"h{n} vs oracle", dh[n].cpu(), inter[f"h{n}"].permute(0, 2, 1))
            pattern(f"skip{n} vs oracle", ds[n].cpu(), inter[f"skip{n}"].permute(0, 2, 1))
        P(f"  out vs oracle rel-L2 = {rel(out.cpu(), want):.3e}")

    # ragged / large dilation / short
    for (C, layers, cycle, B, L, seed) in [(256, 13, 12, 1, 4500, 13), (256, 2, 12, 3, 77, 14)]:
        sd = make_wavenet_state_dict(C, layers, seed)
        g = torch.Generator().manual_seed(seed + 1000)
        audio = torch.randn(B, L, generator=g)
        g2 = torch.Generator().manual_seed(seed + 2000)
        t = torch.randn(B, generator=g2) * 1.5
        want = owav.wavenet_forward(sd, audio, t, cycle)
        for prec in ("fp32", "bf16"):
            net = WaveNetNoise(C, layers, cycle, precision=prec)
            net.load_state_dict(sd, strict=True)
            net = net.to(dev)
            try:
                out = net(audio.to(dev), t.to(dev))
                _native.check_async()
                P(f"== {prec} C={C} layers={layers} B={B} L={L}: out rel-L2 vs oracle = {rel(out.cpu(), want):.3e}")
            except Exception as e:                             # noqa: BLE001
                P(f"== {prec} C={C} layers={layers} B={B} L={L}: FAILED {e!r}")

    # rough timing of the full-size network
    C, layers, cycle = 256, 36, 12
    sd = make_wavenet_state_dict(C, layers, 0)
    net = WaveNetNoise(C, layers, cycle, precision="bf16")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev)
    for B in (8, 64):
        x = torch.randn(B, 16000, device=dev)
        tt = torch.zeros(B, device=dev)
        try:
            net(x, tt)
            _native.check_async()
            net.set_timing(True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                net(x, tt)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 3
            tm = net.timers()
            net.set_timing(False)
            fl = 606.093e9 * B
            P(f"== bf16 full net B={B}: {dt * 1e3:.2f} ms/eval -> {fl / dt / 1e12:.1f} TFLOP/s ; "
              f"conv kernels {tm['conv'][0] / 3:.2f} ms ({fl / (tm['conv'][0] / 3 * 1e-3) / 1e12:.1f} TFLOP/s), "
              f"aux {tm['aux'][0] / 3:.2f} ms")
        except Exception as e:                                 # noqa: BLE001
            P(f"== bf16 full net B={B}: FAILED {e!r}")
            break

    with open(os.path.join(OUT, "diag_tc.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
