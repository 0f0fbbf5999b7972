"""GPU parity: DiffWave backbone, fused denoiser and fused sampling trajectory (through the C ABI)
against the CPU oracle and the committed reference goldens.

Tolerances (BASELINE.json north_star): fp32 path <= 1e-5 relative (rel-L2 over the tensor) per
denoiser call; bf16 path <= 2e-2 relative per call.
"""
import os

import pytest
import torch

from conftest import GOLDEN, load_golden, record_parity, rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "bf16": 2e-2}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def make_net(C, layers, cycle, seed, precision, dev):
    from audiodiffuser_b200 import WaveNetNoise
    from oracle.weights import make_wavenet_state_dict
    net = WaveNetNoise(C, layers, cycle, precision=precision)
    net.load_state_dict(make_wavenet_state_dict(C, layers, seed), strict=True)
    return net.to(dev)


@pytest.mark.parametrize("name,precisions", [
    ("wavenet_c64_l4", ["fp32"]),                       # C = 64: fp32 path only (tensor-core path is C = 256)
    ("wavenet_c256_l3", ["fp32", "bf16"]),
    ("wavenet_c256_l13_dil2048", ["fp32", "bf16"]),     # dilation up to 2048 > L/2: padding on both sides
    ("wavenet_c256_l2_short", ["fp32", "bf16"]),        # L = 77 < one tile, ragged
])
def test_backbone_vs_reference_golden(dev, name, precisions):
    from audiodiffuser_b200 import _native as N
    g = load_golden(name)
    C, layers, cycle, B, L, seed = (int(v) for v in g["cfg"])
    for precision in precisions:
        net = make_net(C, layers, cycle, seed, precision, dev)
        out = net(torch.from_numpy(g["audio"]).to(dev), torch.from_numpy(g["t"]).to(dev))
        N.check_async()
        assert out.shape == (B, 1, L)
        assert rel_l2(out, g["out"]) < TOL[precision], (precision, rel_l2(out, g["out"]))


def test_bf16_path_rejects_other_widths(dev):
    from audiodiffuser_b200 import _native as N
    net = make_net(64, 2, 2, 1, "bf16", dev)
    with pytest.raises(N.AdbError):
        net(torch.zeros(1, 128, device=dev), torch.zeros(1, device=dev))


def test_random_init_is_zero_like_reference(dev):
    """wavenet.py:57-66: the reference's own init gives an exactly-zero network output."""
    from audiodiffuser_b200 import WaveNetNoise
    torch.manual_seed(0)
    net = WaveNetNoise(256, 2, 12, precision="bf16").to(dev)
    out = net(torch.randn(1, 300, device=dev), torch.zeros(1, device=dev))
    assert float(out.abs().max()) == 0.0


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_denoiser_vs_golden(dev, precision):
    """denoise_small golden is C = 64 (fp32 path); the bf16 case uses the oracle directly at C = 256."""
    from audiodiffuser_b200 import EluDiffusion, EDMDenoiser, _native as N
    diff = EluDiffusion(sigma_data=0.2)
    if precision == "fp32":
        g = load_golden("denoise_small")
        C, layers, cycle, B, L, seed = (int(v) for v in g["cfg"])
        net = make_net(C, layers, cycle, seed, "fp32", dev)
        x = torch.from_numpy(g["x"]).to(dev)
        D = EDMDenoiser(net, diff)
        for s in (80.0, 10.0, 1.0, 0.1, 0.002):
            out = D(x * s, s)
            assert rel_l2(out, g[f"sigma_{s}"]) < TOL["fp32"], s
        out = diff.denoise_fn(x, net=net, sigmas=torch.from_numpy(g["sigmas_per_sample"]).to(dev), inference=False)
        assert rel_l2(out, g["per_sample"]) < TOL["fp32"]
    else:
        from oracle import edm, wavenet
        from oracle.weights import make_wavenet_state_dict
        C, layers, cycle, seed, B, L = 256, 5, 12, 33, 2, 700
        sd = make_wavenet_state_dict(C, layers, seed)
        net = make_net(C, layers, cycle, seed, "bf16", dev)
        x = torch.randn(B, 1, L, generator=torch.Generator().manual_seed(34))
        for s in (80.0, 1.0, 0.002):
            want = edm.denoise(x * s, wavenet.make_net_fn(sd, cycle), 0.2, sigma=s)
            out = diff.denoise_fn((x * s).to(dev), net=net, sigma=s, inference=True)
            assert rel_l2(out, want) < TOL["bf16"], s
    N.check_async()


def test_fused_samplers_vs_golden_fp32(dev):
    from audiodiffuser_b200 import EluDiffusion, EDMSampler, EDMAlphaSampler, _native as N
    g = load_golden("sampler_small")
    C, layers, cycle, B, L, seed, steps = (int(v) for v in g["cfg"])
    net = make_net(C, layers, cycle, seed, "fp32", dev)
    diff = EluDiffusion(0.2)
    noise, sig = torch.from_numpy(g["noise"]).to(dev), torch.from_numpy(g["sigmas"]).to(dev)
    s = EDMSampler(s_churn=0.0, s_noise=1.0, num_steps=steps)
    assert rel_l2(s(noise, fn=diff.denoise_fn, net=net, sigmas=sig), g["heun"]) < 5e-5
    assert s.last_nfe == int(g["nfe_heun"])
    # the generic (python-loop) path must agree with the fused trajectory
    assert rel_l2(s(noise, fn=diff.denoise_fn, net=net, sigmas=sig, _force_generic=True), g["heun"]) < 5e-5
    s = EDMSampler(s_churn=0.0, s_noise=1.0, num_steps=steps, use_heun=False)
    assert rel_l2(s(noise, fn=diff.denoise_fn, net=net, sigmas=sig), g["euler"]) < 5e-5
    assert s.last_nfe == int(g["nfe_euler"])
    # churn: replay the reference's CPU RNG stream (one randn_like per step, sampler_edm.py:346)
    torch.manual_seed(int(g["churn_seed"]))
    eps = torch.stack([torch.randn(B, 1, L) for _ in range(steps)]).to(dev)
    s = EDMSampler(s_tmin=0.05, s_tmax=50.0, s_churn=2.0, s_noise=1.003, num_steps=steps)
    assert rel_l2(s(noise, fn=diff.denoise_fn, net=net, sigmas=sig, eps=eps), g["churn"]) < 5e-5
    assert s.last_nfe == int(g["nfe_churn"])
    for alpha, key in ((1.0, "alpha1"), (0.5, "alpha05")):
        a = EDMAlphaSampler(alpha=alpha, num_steps=steps)
        assert rel_l2(a(noise, fn=diff.denoise_fn, net=net, sigmas=sig), g[key]) < 5e-5
        assert a.last_nfe == int(g["nfe_" + key])
    N.check_async()


def test_dsm_loss_vs_golden(dev):
    from audiodiffuser_b200 import EluDiffusion
    g = load_golden("dsm_loss_small")
    C, layers, cycle, B, L, seed = (int(v) for v in g["cfg"])
    net = make_net(C, layers, cycle, seed, "fp32", dev)
    loss = EluDiffusion(0.2)(torch.from_numpy(g["x"]).to(dev), net, sigmas=torch.from_numpy(g["sigmas"]).to(dev),
                             noise=torch.from_numpy(g["noise"]).to(dev))
    assert torch.allclose(loss.cpu(), torch.from_numpy(g["loss"]), rtol=1e-4)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_size_vs_reference_golden(dev, precision):
    """BASELINE shape: C = 256, 36 layers, L = 16000, B = 1 — per-call denoiser output and the final
    18-step Heun waveform against the reference's own outputs."""
    if not os.path.exists(os.path.join(GOLDEN, "full_diffwave_b1.npz")):
        pytest.skip("full-size golden missing")
    from audiodiffuser_b200 import EluDiffusion, EDMSampler, _native as N
    g = load_golden("full_diffwave_b1")
    C, layers, cycle, B, L, seed, steps = (int(v) for v in g["cfg"])
    net = make_net(C, layers, cycle, seed, precision, dev)
    diff = EluDiffusion(0.2)
    noise = torch.from_numpy(g["noise"]).to(dev)
    for s in (80.0, 1.0, 0.002):
        out = diff.denoise_fn(noise * s, net=net, sigma=s, inference=True)
        e32 = rel_l2(out, g[f"den_sigma_{s}"])           # vs the reference's own fp32 output
        e64 = rel_l2(out, g[f"den64_sigma_{s}"])         # vs the same algorithm evaluated in fp64
        print(f"full-size {precision} sigma={s}: rel-L2 vs ref fp32 {e32:.3e}, vs fp64 {e64:.3e}")
        record_parity(f"full_size_b1_{precision}_denoiser_sigma_{s}", rel_l2_vs_reference_fp32=e32, rel_l2_vs_fp64=e64)
        if precision == "fp32":
            # The reference's fp32 output is itself 1.03e-5 (sigma=80) / 8.3e-6 (sigma=1) away from the fp64
            # evaluation after 36 layers, so two independent fp32 implementations can differ by up to the sum
            # of their errors: hold 1e-5 against the fp64 evaluation and 2e-5 against the reference's fp32.
            assert e64 < 1e-5, (s, e64)
            assert e32 < 2e-5, (s, e32)
        else:
            assert e32 < TOL["bf16"], (s, e32)
    smp = EDMSampler(s_churn=0.0, s_noise=1.0, num_steps=steps)
    x = smp(noise, fn=diff.denoise_fn, net=net, sigmas=torch.from_numpy(g["sigmas"]).to(dev))
    N.check_async()
    assert smp.last_nfe == 35
    want = torch.from_numpy(g["heun18"]).double()
    err = (x.cpu().double() - want).norm() / want.norm()
    snr_db = -20.0 * torch.log10(err)
    record_parity(f"full_size_b1_{precision}_heun18_waveform", rel_l2_vs_reference=float(err), snr_db=float(snr_db), nfe=smp.last_nfe)
    # final-waveform tolerance: SNR >= 80 dB fp32, >= 25 dB bf16 after 35 evaluations
    assert snr_db > (80.0 if precision == "fp32" else 25.0), (precision, float(snr_db))


@pytest.mark.parametrize("B", [64, 300])
def test_full_size_batch_vs_reference_golden(dev, B):
    """Parity AT the benchmarked configuration (BASELINE.json configs[1]: 36 layers, L = 16000, bf16, 18-step EDM-Heun, a
    batch that spans several tiles-per-SM rounds; B = 300 also crosses the 256-sample pass boundary of the z-stash path).
    Rows 0, B/2 and B-1 carry the reference golden's noise: after the whole trajectory (i) each of them is within the bf16
    gate of the reference's own waveform, (ii) all three are BIT-identical to each other and to the same noise sampled
    alone (B = 1) — whatever tile, CTA-pair half, batch position or pass computed them; (iii) no other row is disturbed
    (finite, different noise gives a different waveform)."""
    if not os.path.exists(os.path.join(GOLDEN, "full_diffwave_b1.npz")):
        pytest.skip("full-size golden missing")
    from audiodiffuser_b200 import EluDiffusion, EDMSampler, _native as N
    g = load_golden("full_diffwave_b1")
    C, layers, cycle, _, L, seed, steps = (int(v) for v in g["cfg"])
    net = make_net(C, layers, cycle, seed, "bf16", dev)
    diff = EluDiffusion(0.2)
    gold = torch.from_numpy(g["noise"]).to(dev)                    # [1, 1, L]
    noise = torch.randn(B, 1, L, generator=torch.Generator().manual_seed(99)).to(dev)
    rows = [0, B // 2, B - 1]
    for r in rows:
        noise[r] = gold[0]
    sig = torch.from_numpy(g["sigmas"]).to(dev)
    smp = EDMSampler(s_churn=0.0, s_noise=1.0, num_steps=steps)
    x = smp(noise, fn=diff.denoise_fn, net=net, sigmas=sig)
    alone = smp(gold, fn=diff.denoise_fn, net=net, sigmas=sig)
    N.check_async()
    assert smp.last_nfe == 35 and torch.isfinite(x).all()
    want = torch.from_numpy(g["heun18"]).double()
    errs = [float((x[r].cpu().double() - want[0]).norm() / want.norm()) for r in rows]
    snr = [-20.0 * float(torch.log10(torch.tensor(e))) for e in errs]
    bit_equal = all(torch.equal(x[r], alone[0]) for r in rows)
    other = float((x[1].cpu().double() - want[0]).norm() / want.norm())
    record_parity(f"full_size_batch_{B}_bf16_heun18", rows=rows, rel_l2_vs_reference=errs, snr_db=snr,
                  rows_bit_identical_to_b1=bit_equal, rel_l2_of_an_unrelated_row=other)
    assert min(snr) > 25.0, snr
    assert bit_equal
    assert other > 0.1                                              # a row with different noise is a different waveform


def test_batch_rows_independent(dev):
    """Sharding property: a sample's output does not depend on what else is in the batch."""
    net = make_net(256, 3, 12, 77, "bf16", dev)
    x = torch.randn(5, 640, device=dev)
    t = torch.randn(5, device=dev)
    full = net(x, t)
    part = net(x[2:4].contiguous(), t[2:4].contiguous())
    assert torch.equal(full[2:4], part)


@pytest.mark.parametrize("name", ["wavenet_c256_l3", "wavenet_c256_l13_dil2048", "wavenet_c256_l2_short"])
def test_block_kernels_vs_reference_golden(dev, name, monkeypatch):
    """All four residual-block kernels of the bf16 path (ADB_BLOCK_KERNEL, read when the native handle is created) against
    the same reference goldens: 0 single-CTA, 1 CTA pair, 2 CTA pair + fp16 skip stash (the training forward / debug
    entry), 3 z-stash kernel + skip GEMM (the default sampling path). The pair kernel without the stash is bit-identical
    to the single-CTA kernel (same MMA order per accumulator element)."""
    from audiodiffuser_b200 import _native as N
    g = load_golden(name)
    C, layers, cycle, B, L, seed = (int(v) for v in g["cfg"])
    audio, t = torch.from_numpy(g["audio"]).to(dev), torch.from_numpy(g["t"]).to(dev)
    outs = {}
    for k in (0, 1, 2, 3):
        monkeypatch.setenv("ADB_BLOCK_KERNEL", str(k))
        outs[k] = make_net(C, layers, cycle, seed, "bf16", dev)(audio, t)
    monkeypatch.delenv("ADB_BLOCK_KERNEL")
    default = make_net(C, layers, cycle, seed, "bf16", dev)(audio, t)
    N.check_async()
    assert torch.equal(default, outs[3])
    assert rel_l2(outs[1], outs[0]) < 1e-6
    errs = {k: rel_l2(v, g["out"]) for k, v in outs.items()}
    print(f"{name}: rel-L2 vs reference per block kernel {errs}")
    for k, e in errs.items():
        assert e < TOL["bf16"], (k, e)
    # one fp16 rounding of every second layer's skip term: far below the bf16 operand error itself
    assert rel_l2(outs[2], outs[1]) < 3e-3
    # the z-stash path sums the skip terms of all blocks in ONE fp32 accumulator (no intermediate rounding at all)
    assert rel_l2(outs[3], outs[1]) < 3e-3, rel_l2(outs[3], outs[1])
    assert errs[3] < 1.05 * errs[1] + 1e-4                       # and it does not move the error against the reference


@pytest.mark.parametrize("chunk,stash_gb", [(2, 80.0), (256, 0.001), (3, 0.035)])
def test_zstash_passes_and_layer_groups(dev, monkeypatch, chunk, stash_gb):
    """The z-stash path processes the batch in passes of ADB_CHUNK samples and contracts the stash every G blocks (G from
    ADB_STASH_GB; here 13 blocks at once, one at a time, and groups of 5 + 5 + 3): any pass size / group size gives the same waveforms (bit-identical across passes; the group size only
    changes where the fp32 partial sums are added)."""
    from audiodiffuser_b200 import _native as N
    g = load_golden("wavenet_c256_l13_dil2048")
    C, layers, cycle, B, L, seed = (int(v) for v in g["cfg"])
    audio, t = torch.from_numpy(g["audio"]).to(dev), torch.from_numpy(g["t"]).to(dev)
    x = torch.cat([audio, audio.flip(0), audio[:1]], 0).contiguous()       # batch 2 B + 1: a ragged last pass
    tt = torch.cat([t, t.flip(0), t[:1]], 0).contiguous()
    base = make_net(C, layers, cycle, seed, "bf16", dev)(x, tt)
    monkeypatch.setenv("ADB_CHUNK", str(chunk))
    monkeypatch.setenv("ADB_STASH_GB", str(stash_gb))
    out = make_net(C, layers, cycle, seed, "bf16", dev)(x, tt)
    N.check_async()
    # a different group size re-associates the fp32 skip sum; the tail rounds it to bf16, so a last-bit change of the sum
    # moves a few outputs by one bf16 step (measured 1.1e-4) — far below the bf16 operand error itself (4e-3)
    assert rel_l2(out, base) < 5e-4, rel_l2(out, base)
    assert rel_l2(out[:B], g["out"]) < TOL["bf16"]
    if stash_gb >= 1.0:
        assert torch.equal(out, base)                                  # passes alone do not change a single bit


@pytest.mark.parametrize("S,pipe", [(1, 0), (2, 0), (2, 1), (5, 1), (4, 1)])
def test_multilayer_wavefront_launch_is_bit_identical(dev, monkeypatch, S, pipe):
    """ADB_ZS_ML = S (opt-in): all blocks of a chunk in ONE launch as a wavefront over (sub-pass of S samples, layer,
    tile group) with per-tile completion flags, ping / pong reused in L2. Same arithmetic per tile, so the waveforms must be
    bit-identical to the one-launch-per-block path — for any sub-pass size, including a ragged last sub-pass and a dilation
    (2048) that reaches across 16 tiles, and with either job order."""
    from audiodiffuser_b200 import _native as N
    g = load_golden("wavenet_c256_l13_dil2048")
    C, layers, cycle, B, L, seed = (int(v) for v in g["cfg"])
    audio, t = torch.from_numpy(g["audio"]).to(dev), torch.from_numpy(g["t"]).to(dev)
    x = torch.cat([audio, audio.flip(0), audio[:1]], 0).contiguous()       # batch 2 B + 1: ragged sub-passes
    tt = torch.cat([t, t.flip(0), t[:1]], 0).contiguous()
    base = make_net(C, layers, cycle, seed, "bf16", dev)(x, tt)
    monkeypatch.setenv("ADB_ZS_ML", str(S))
    monkeypatch.setenv("ADB_ZS_PIPE", str(pipe))
    out = make_net(C, layers, cycle, seed, "bf16", dev)(x, tt)
    N.check_async()
    assert torch.equal(out, base)
    assert rel_l2(out[:B], g["out"]) < TOL["bf16"]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_debug_per_block_outputs_vs_oracle(dev, precision):
    """adb_wavenet_forward_debug: the residual stream h and the running skip sum after every block against the oracle's
    intermediates on the same seeded inputs (the skip stash is switched off on this entry point, so every running sum
    is complete)."""
    from oracle import wavenet as owav
    from oracle.weights import make_wavenet_state_dict
    C, layers, cycle, seed, B, L = 256, 4, 3, 31, 2, 700
    sd = make_wavenet_state_dict(C, layers, seed)
    audio = torch.randn(B, L, generator=torch.Generator().manual_seed(seed + 1))
    t = torch.tensor([-0.4, 0.9])
    with torch.no_grad():
        want, inter = owav.wavenet_forward(sd, audio, t, cycle, return_intermediates=True)
    net = make_net(C, layers, cycle, seed, precision, dev)
    out, dh, ds = net.forward_debug(audio.to(dev), t.to(dev), layers)
    tol = TOL[precision]
    assert rel_l2(out, want) < tol
    for n in range(layers):
        if n + 1 < layers:                                           # the last block's residual output is never computed
            assert rel_l2(dh[n].transpose(1, 2), inter[f"h{n}"]) < tol, (n, "h")
        assert rel_l2(ds[n].transpose(1, 2), inter[f"skip{n}"]) < tol, (n, "skip")
    assert rel_l2(out, net(audio.to(dev), t.to(dev))) < (1e-6 if precision == "fp32" else 3e-3)   # same result as the plain entry
