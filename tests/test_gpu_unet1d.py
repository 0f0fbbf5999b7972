"""GPU parity: the 1-D U-Net (through the adb_cl_* C ABI) against goldens from the reference's own UNet1dBase.

Tolerances (BASELINE.json north_star): fp32 path <= 1e-5 rel-L2 per call, bf16 path <= 2e-2."""
import pytest
import torch

from conftest import record_parity, load_golden, rel_l2

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-5, "bf16": 2e-2}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def make_unet(cfg, seed, precision, dev):
    from audiodiffuser_b200 import UNet1dBase
    from oracle.weights import make_unet1d_state_dict
    net = UNet1dBase(precision=precision, **cfg)
    net.load_state_dict(make_unet1d_state_dict(cfg, seed), strict=True)
    return net.to(dev)


@pytest.mark.parametrize("name,precisions", [
    ("unet1d_small", ["fp32"]),                 # channels = 32: K not a multiple of 64 everywhere -> fp32 path only
    ("unet1d_small_ragged", ["fp32"]),
    ("unet1d_mid", ["fp32", "bf16"]),
    ("unet1d_cfg4_l65536", ["fp32", "bf16"]),   # BASELINE config 4 architecture (102 M parameters)
    ("unet1d_cfg4_l262144", ["fp32", "bf16"]),  # the same at the config's real length, 2 x 262144
])
def test_unet_vs_reference_golden(dev, name, precisions):
    from audiodiffuser_b200 import _native as N
    from oracle.weights import UNET_CASES
    cfg, B, L, seed = UNET_CASES[name]
    g = load_golden(name)
    x, t = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["t"]).to(dev)
    for precision in precisions:
        net = make_unet(cfg, seed, precision, dev)
        net.use_cuda_graph = False
        out = net(x, t)
        N.check_async()
        assert out.shape == (B, cfg["in_channels"], L)
        e = rel_l2(out, g["out"])
        print(f"{name} {precision}: rel-L2 {e:.3e}")
        record_parity(f"{name}_{precision}", rel_l2_vs_reference=e)
        assert e < TOL[precision], (precision, e)
        net.use_cuda_graph = True               # the captured graph must reproduce the eager launch sequence exactly
        out_g = net(x, t)
        out_g2 = net(x, t)
        N.check_async()
        assert torch.equal(out_g, out) and torch.equal(out_g2, out)


def test_unet_zero_init_like_reference(dev):
    """unet1d.py:619: the reference zero-initialises the output transposed conv -> exactly-zero output."""
    from audiodiffuser_b200 import UNet1dBase
    from oracle.weights import UNET_SMALL
    torch.manual_seed(0)
    net = UNet1dBase(precision="fp32", **UNET_SMALL).to(dev)
    out = net(torch.randn(1, 2, 256, device=dev), torch.zeros(1, device=dev))
    assert float(out.abs().max()) == 0.0


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unet_edm_denoiser_and_sampler(dev, precision):
    """EluDiffusion.denoise_fn / EDMSampler driving the fused UNet (generic fn/net protocol, diffusion.py:50)."""
    from audiodiffuser_b200 import EluDiffusion, EDMSampler, _native as N
    from oracle.weights import UNET_MID
    g = load_golden("unet1d_mid_edm")
    B, L, seed, steps = (int(v) for v in g["cfg"])
    net = make_unet(UNET_MID, seed, precision, dev)
    diff = EluDiffusion(0.2)
    noise = torch.from_numpy(g["noise"]).to(dev)
    for s in (80.0, 1.0, 0.002):
        e = rel_l2(diff.denoise_fn(noise * s, net=net, sigma=s, inference=True), g[f"den_sigma_{s}"])
        assert e < TOL[precision], (precision, s, e)
    smp = EDMSampler(s_churn=0.0, s_noise=1.0, num_steps=steps)
    x = smp(noise, fn=diff.denoise_fn, net=net, sigmas=torch.from_numpy(g["sigmas"]).to(dev))
    N.check_async()
    assert smp.last_nfe == 2 * steps - 1
    e = rel_l2(x, g["heun"])
    print(f"unet EDM heun {precision}: rel-L2 {e:.3e}")
    record_parity(f"unet1d_mid_edm_heun_{precision}", rel_l2_vs_reference=e, nfe=smp.last_nfe)
    # measured 3.6e-7 (fp32) / 1.5e-3 (bf16): the gates leave a factor ~6, not 60
    assert e < (5e-6 if precision == "fp32" else 1e-2), (precision, e)


def test_unet_rejects_conditioning(dev):
    from audiodiffuser_b200 import UNet1dBase
    from oracle.weights import UNET_SMALL
    with pytest.raises(NotImplementedError):
        UNet1dBase(**dict(UNET_SMALL, text_cond=True))
    net = UNet1dBase(precision="fp32", **UNET_SMALL).to(dev)
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 2, 256, device=dev), torch.zeros(1, device=dev), classes=torch.zeros(1, dtype=torch.long, device=dev))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unet_class_conditioning_and_cfg(dev, precision):
    """Label conditioning (LabelEmbedder, conditioner.py:59-111) and classifier-free guidance through denoise_fn
    (diffusion.py:50-54): the fused path evaluates both guidance branches as one batch-2B network call."""
    from audiodiffuser_b200 import EluDiffusion, _native as N
    from oracle.weights import UNET_CLASS
    g = load_golden("unet1d_class_cfg")
    B, L, seed = (int(v) for v in g["cfg"])
    net = make_unet(UNET_CLASS, seed, precision, dev)
    x, t, cls = (torch.from_numpy(g[k]).to(dev) for k in ("x", "t", "classes"))
    for p, key in ((0.0, "f_cond"), (1.0, "f_null")):
        e = rel_l2(net(x, t, classes=cls, cond_drop_prob=p), g[key])
        assert e < TOL[precision], (precision, key, e)
    diff = EluDiffusion(0.2)
    for s in (10.0, 0.5):
        out = diff.denoise_fn(x * s, net=net, sigma=s, inference=True, cond_scale=2.5, classes=cls)
        e = rel_l2(out, g[f"den_cfg_sigma_{s}"])
        print(f"class-conditioned CFG {precision} sigma={s}: rel-L2 {e:.3e}")
        assert e < (1e-5 if precision == "fp32" else 3e-2), (precision, s, e)
    N.check_async()
    with pytest.raises(NotImplementedError):
        net(x, t, classes=cls, cond_drop_prob=0.3)


@pytest.mark.parametrize("L", [16, 32, 64, 128, 24])
def test_attention_kernels_vs_fp64_softmax_attention(dev, L, monkeypatch):
    """adb_cl_attention (attention_utils.py:163-184) at the U-Net's head dimension 64: the mma.sync tensor-core kernel
    (L % 16 == 0, L <= 128), the CUDA-core kernel it falls back to (L = 24; ADB_NO_MMA_ATTENTION) and the fp32 kernel,
    against softmax(q k^T / 8) v evaluated in fp64 on the same bf16-rounded inputs."""
    from audiodiffuser_b200 import _native as N
    lib, st = N.lib(), N.stream_ptr(dev)
    B, heads, d = 3, 8, 64
    C = heads * d
    gen = torch.Generator().manual_seed(L)
    q = torch.randn(B, L, C, generator=gen).to(dev).to(torch.bfloat16)
    kv = (torch.randn(B, L, 2 * C, generator=gen) * 1.5).to(dev).to(torch.bfloat16)
    qd, kd, vd = q.double(), kv[..., :C].double(), kv[..., C:].double()
    split = lambda a: a.reshape(B, L, heads, d).transpose(1, 2)            # noqa: E731
    p = torch.softmax(split(qd) @ split(kd).transpose(-1, -2) / 8.0, dim=-1)
    want = (p @ split(vd)).transpose(1, 2).reshape(B, L, C)

    def run(dtype_id, qq, kk):
        o = torch.empty_like(qq)
        N.check(lib.adb_cl_attention(N.ptr(qq), N.ptr(kk), N.ptr(o), B, L, C, heads, dtype_id, st))
        N.check_async()
        return o

    got = run(1, q, kv)                                                    # bf16: mma kernel when L % 16 == 0
    assert rel_l2(got, want) < 6e-3, rel_l2(got, want)
    monkeypatch.setenv("ADB_NO_MMA_ATTENTION", "1")
    ref = run(1, q, kv)                                                    # bf16 CUDA-core kernel
    assert rel_l2(ref, want) < 4e-3
    assert rel_l2(got, ref) < 6e-3
    assert rel_l2(run(0, q.float(), kv.float()), want) < 1e-5              # fp32 kernel


@pytest.mark.parametrize("precision,churn", [("fp32", 0.0), ("bf16", 0.0), ("fp32", 40.0)])
def test_unet_device_resident_trajectory_matches_generic_loop(dev, precision, churn):
    """UNet1dBase._adb_fused_sample (state in static buffers, two CUDA graphs, fused mid / post kernels) against the generic
    Python loop of the same sampler and against the reference golden: same arithmetic, so fp32 agrees to round-off."""
    from audiodiffuser_b200 import EluDiffusion, EDMSampler, _native as N
    from oracle.weights import UNET_MID
    g = load_golden("unet1d_mid_edm")
    B, L, seed, steps = (int(v) for v in g["cfg"])
    net = make_unet(UNET_MID, seed, precision, dev)
    diff = EluDiffusion(0.2)
    noise = torch.from_numpy(g["noise"]).to(dev)
    sig = torch.from_numpy(g["sigmas"]).to(dev)
    smp = EDMSampler(s_churn=churn, s_noise=1.003, num_steps=steps)
    fused = smp(noise, fn=diff.denoise_fn, net=net, sigmas=sig, churn_seed=7)
    nfe = smp.last_nfe
    generic = smp(noise, fn=diff.denoise_fn, net=net, sigmas=sig, churn_seed=7, _force_generic=True)
    N.check_async()
    assert nfe == smp.last_nfe == 2 * steps - 1
    assert ("traj", B, L, dev.index, 0.2) in net._graphs                      # the device-resident path really ran
    e = rel_l2(fused, generic)
    record_parity(f"unet1d_fused_trajectory_vs_generic_{precision}_churn{churn}", rel_l2=e)
    assert e < (2e-6 if precision == "fp32" else 5e-3), e
    if churn == 0.0:
        assert rel_l2(fused, g["heun"]) < (5e-6 if precision == "fp32" else 1e-2)
    again = smp(noise, fn=diff.denoise_fn, net=net, sigmas=sig, churn_seed=7)
    assert torch.equal(again, fused)                                        # static buffers are fully rewritten per call
