"""GPU parity of the fused GroupNorm-apply convolution (adb_cl_gn_conv3, ConvBlock1d of unet1d.py:160-207 with the channel
concatenation of UpsampleBlock1d :552-556 folded in) and of the two-input plain convolution (adb_cl_conv_cat):

  * against the unfused C-ABI sequence adb_cl_concat -> adb_cl_groupnorm -> adb_cl_conv on the same bf16 inputs (the fused kernel
    feeds the tensor core fp16 operands — normalised activations and weights with 11 significand bits — where the unfused path
    rounds both to bf16, so it is the more accurate of the two),
  * against torch's group_norm / silu / conv1d evaluated in fp64 on the same bf16-rounded inputs.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import record_parity, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


CASES = [  # B, L, C1, C2, N, cond, res
    (2, 300, 64, 0, 64, False, False),        # ragged length, one K-block, N tile 64
    (3, 1000, 128, 128, 256, True, True),     # concatenation, (scale, shift) conditioning, residual
    (2, 384, 512, 512, 512, True, True),      # Cin = 1024, two n-tiles, odd number of m-tiles (padding tile in the last pair)
    (5, 17, 256, 0, 128, True, False),        # tiles far from full
    (1, 4096, 256, 0, 256, False, True),      # 32 tiles of one sample
]


@pytest.mark.parametrize("B,L,C1,C2,N,cond,res", CASES)
def test_gn_conv3_vs_unfused_and_fp64(dev, B, L, C1, C2, N, cond, res):
    from audiodiffuser_b200 import _native as Nn
    lib, st = Nn.lib(), Nn.stream_ptr(dev)
    G, eps, s2 = 8, 1e-5, 2 ** -0.5
    Cin = C1 + C2
    gen = torch.Generator().manual_seed(B * 1000 + L)
    rnd = lambda *s: torch.randn(*s, generator=gen)                                # noqa: E731
    h = (rnd(B, L, C1) * 1.7 + 0.3).to(dev).to(torch.bfloat16)
    sk = (rnd(B, L, C2) * 0.8 - 0.2).to(dev).to(torch.bfloat16) if C2 else None
    gamma, beta = (1 + 0.3 * rnd(Cin)).to(dev), (0.2 * rnd(Cin)).to(dev)
    ss = (0.3 * rnd(B, 2 * Cin + 6)).to(dev) if cond else None                     # row pitch larger than 2 Cin on purpose
    w = (rnd(3, Cin, N) / math.sqrt(3 * Cin)).to(dev)
    bias = (0.1 * rnd(N)).to(dev)
    r = rnd(B, L, N).to(dev).to(torch.bfloat16) if res else None
    packed = torch.empty(lib.adb_cl_conv_packed_elems(Cin, N, 3), dtype=torch.bfloat16, device=dev)
    Nn.check(lib.adb_cl_pack_conv_weights(Nn.ptr(w), Nn.ptr(packed), Cin, N, 3, st))
    packed16 = torch.empty(lib.adb_cl_conv_packed_elems(Cin, N, 3), dtype=torch.float16, device=dev)     # the fused kernel's fp16 operand
    Nn.check(lib.adb_cl_pack_conv_weights_f16(Nn.ptr(w), Nn.ptr(packed16), Cin, N, 3, st))

    # fused: statistics + coefficients of the raw inputs, then one kernel
    sums = torch.zeros(B * G * 2, dtype=torch.float64, device=dev)
    tickets = torch.zeros(B, dtype=torch.int32, device=dev)
    coef = torch.empty(2, B, Cin, dtype=torch.float32, device=dev)
    cpg = Cin // G
    g1 = C1 // cpg
    ss_ld = ss.shape[1] if cond else 0
    out = torch.empty(B, L, N, dtype=torch.bfloat16, device=dev)
    for rep in range(2):                                   # twice: the kernels must leave sums / tickets zero for the next use
        Nn.check(lib.adb_cl_gn_coef(Nn.ptr(h), Nn.ptr(sums), Nn.ptr(tickets), Nn.ptr(coef), B, L, C1, g1, G, 0, 0, Cin, Nn.ptr(gamma),
                                    Nn.ptr(beta), Nn.ptr(ss), ss_ld, eps, 1.0, st))
        if C2:
            Nn.check(lib.adb_cl_gn_coef(Nn.ptr(sk), Nn.ptr(sums), Nn.ptr(tickets), Nn.ptr(coef), B, L, C2, G - g1, G, g1, C1, Cin,
                                        Nn.ptr(gamma), Nn.ptr(beta), Nn.ptr(ss), ss_ld, eps, s2, st))
        out.zero_()
        Nn.check(lib.adb_cl_gn_conv3(Nn.ptr(h), C1, Nn.ptr(sk), C2, Nn.ptr(coef), Nn.ptr(packed16), Nn.ptr(bias), Nn.ptr(r), Nn.ptr(out),
                                     B, L, N, st))
        Nn.check_async()
        assert float(sums.abs().max()) == 0.0 and int(tickets.abs().max()) == 0

    # unfused sequence of the same library
    if C2:
        cat = torch.empty(B, L, Cin, dtype=torch.bfloat16, device=dev)
        Nn.check(lib.adb_cl_concat(Nn.ptr(h), Nn.ptr(sk), s2, Nn.ptr(cat), B * L, C1, C2, 1, st))
    else:
        cat = h
    xn = torch.empty_like(cat)
    sums2 = torch.empty(B * G * 2, dtype=torch.float64, device=dev)
    Nn.check(lib.adb_cl_groupnorm(Nn.ptr(cat), Nn.ptr(gamma), Nn.ptr(beta), Nn.ptr(ss), ss.shape[1] if cond else 0, Nn.ptr(xn),
                                  Nn.ptr(sums2), B, L, Cin, G, eps, 2, 1, st))
    ref = torch.empty(B, L, N, dtype=torch.bfloat16, device=dev)
    Nn.check(lib.adb_cl_conv(Nn.ptr(xn), Nn.ptr(packed), Nn.ptr(bias), Nn.ptr(r), Nn.ptr(ref), B, L, L, Cin, N, 3, -1, 1, 0, 0, 0, 0, 1, st))
    Nn.check_async()

    # fp64 torch evaluation on the same bf16-rounded inputs (channels-first)
    x64 = h.double() if not C2 else torch.cat([h.double(), sk.double() * s2], dim=2)
    y = F.group_norm(x64.transpose(1, 2), G, gamma.double(), beta.double(), eps)
    if cond:
        y = y * (ss[:, :Cin].double().unsqueeze(2) + 1) + ss[:, Cin:2 * Cin].double().unsqueeze(2)
    y = F.silu(y)
    want = F.conv1d(y, w.double().permute(2, 1, 0), bias.double(), padding=1).transpose(1, 2)     # unrounded weights: the two paths round them differently
    if res:
        want = want + r.double()

    e_ref, e_fused, e_pair = rel_l2(ref, want), rel_l2(out, want), rel_l2(out, ref)
    print(f"gn_conv3 B{B} L{L} C{C1}+{C2} N{N}: unfused {e_ref:.2e} fused {e_fused:.2e} fused-vs-unfused {e_pair:.2e}")
    record_parity(f"gn_conv3_B{B}_L{L}_C{C1}+{C2}_N{N}", fused_vs_fp64=e_fused, unfused_vs_fp64=e_ref, fused_vs_unfused=e_pair)
    assert e_fused < 6e-3, e_fused                       # bf16 operands + bf16 output rounding
    assert e_fused < 1.2 * e_ref + 1e-4                  # not worse than the unfused path
    assert e_pair < 6e-3, e_pair


@pytest.mark.parametrize("B,L,C1,C2,N,taps", [(2, 300, 64, 64, 64, 1), (3, 777, 256, 256, 256, 1), (2, 256, 128, 64, 128, 3)])
def test_conv_cat_vs_conv_on_concatenation(dev, B, L, C1, C2, N, taps):
    from audiodiffuser_b200 import _native as Nn
    lib, st = Nn.lib(), Nn.stream_ptr(dev)
    Cin = C1 + C2
    gen = torch.Generator().manual_seed(L + taps)
    a = torch.randn(B, L, C1, generator=gen).to(dev).to(torch.bfloat16)
    b = torch.randn(B, L, C2, generator=gen).to(dev).to(torch.bfloat16)
    w = (torch.randn(taps, Cin, N, generator=gen) / math.sqrt(taps * Cin)).to(dev)
    bias = (0.1 * torch.randn(N, generator=gen)).to(dev)
    packed = torch.empty(lib.adb_cl_conv_packed_elems(Cin, N, taps), dtype=torch.bfloat16, device=dev)
    Nn.check(lib.adb_cl_pack_conv_weights(Nn.ptr(w), Nn.ptr(packed), Cin, N, taps, st))
    cat = torch.cat([a, b], dim=2).contiguous()
    ref = torch.empty(B, L, N, dtype=torch.bfloat16, device=dev)
    out = torch.empty_like(ref)
    off0 = -(taps // 2)
    Nn.check(lib.adb_cl_conv(Nn.ptr(cat), Nn.ptr(packed), Nn.ptr(bias), Nn.ptr(None), Nn.ptr(ref), B, L, L, Cin, N, taps, off0, 1, 0, 0, 0, 0, 1, st))
    Nn.check(lib.adb_cl_conv_cat(Nn.ptr(a), C1, Nn.ptr(b), C2, Nn.ptr(packed), Nn.ptr(bias), Nn.ptr(None), Nn.ptr(out), B, L, N, taps,
                                 off0, 1, 0, st))
    Nn.check_async()
    assert torch.equal(out, ref)                          # same K order, same operands: bit-identical


def test_conv_ktrim_skips_only_zero_blocks(dev):
    """adb_cl_conv_ktrim on a strided-convolution weight (Downsample1d on the [L/f][f*C] view, unet1d.py:214-225: kernel 2f + 1, the
    third coarse tap holds ONE fine tap) must be bit-identical to adb_cl_conv on the same packed weights: the skipped K-blocks
    multiply zeros."""
    from audiodiffuser_b200 import _native as Nn
    lib, st = Nn.lib(), Nn.stream_ptr(dev)
    B, L, ci, f, co = 3, 1000, 128, 4, 256
    gen = torch.Generator().manual_seed(7)
    w = torch.randn(co, ci, 2 * f + 1, generator=gen) / math.sqrt(ci * (2 * f + 1))
    wc = torch.zeros(3, f * ci, co)
    for j in range(3):
        for ph in range(f):
            kk = j * f + ph
            if kk < 2 * f + 1:
                wc[j, ph * ci:(ph + 1) * ci, :] = w[:, :, kk].t()
    wc = wc.to(dev)
    bias = (0.1 * torch.randn(co, generator=gen)).to(dev)
    x = torch.randn(B, L // f, f * ci, generator=gen).to(dev).to(torch.bfloat16)          # the [L/f][f*C] view
    packed = torch.empty(lib.adb_cl_conv_packed_elems(f * ci, co, 3), dtype=torch.bfloat16, device=dev)
    Nn.check(lib.adb_cl_pack_conv_weights(Nn.ptr(wc), Nn.ptr(packed), f * ci, co, 3, st))
    rows = L // f
    full = torch.empty(B, rows, co, dtype=torch.bfloat16, device=dev)
    trim = torch.empty_like(full)
    Nn.check(lib.adb_cl_conv(Nn.ptr(x), Nn.ptr(packed), Nn.ptr(bias), Nn.ptr(None), Nn.ptr(full), B, rows, rows, f * ci, co, 3, -1, 1, 0, 0, 0, 0, 1, st))
    Nn.check(lib.adb_cl_conv_ktrim(Nn.ptr(x), Nn.ptr(packed), Nn.ptr(bias), Nn.ptr(None), Nn.ptr(trim), B, rows, rows, f * ci, co, 3, -1, 1, 0,
                                   ci // 64, 1, st))
    Nn.check_async()
    assert torch.equal(trim, full)
    want = F.conv1d(x.double().reshape(B, L, ci).transpose(1, 2), w.to(dev).to(torch.bfloat16).double(), bias.double(), stride=f, padding=f)
    assert rel_l2(trim, want.transpose(1, 2)) < 4e-3


def test_wavenc_prep_folds_the_input_scale(dev):
    """adb_cl_wavenc_prep with a per-sample scale == the same re-layout of the pre-scaled input (the EDM input scale c_in of the
    device-resident U-Net trajectory, diffusion.py:46-48), and adb_edm_precond_coef yields the c_in / c_noise that
    adb_edm_precond_in applies."""
    from audiodiffuser_b200 import _native as Nn
    lib, st = Nn.lib(), Nn.stream_ptr(dev)
    B, cin, L, W, S = 3, 2, 4096, 32, 16
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(B, cin, L, generator=gen).to(dev)
    sig = torch.tensor([0.7], device=dev)
    c_in, c_noise = torch.empty(B, device=dev), torch.empty(B, device=dev)
    Nn.check(lib.adb_edm_precond_coef(Nn.ptr(sig), 0, 0.2, Nn.ptr(c_in), Nn.ptr(c_noise), B, st))
    net_in, c_noise2 = torch.empty_like(x), torch.empty(B, device=dev)
    Nn.check(lib.adb_edm_precond_in(Nn.ptr(x), Nn.ptr(sig), 0, 0.2, Nn.ptr(net_in), Nn.ptr(c_noise2), B, cin * L, st))
    rows = L // W + 1
    a = torch.empty(B, rows, W * cin, dtype=torch.bfloat16, device=dev)
    b = torch.empty_like(a)
    Nn.check(lib.adb_cl_wavenc_prep(Nn.ptr(x), Nn.ptr(a), B, cin, L, W, S, Nn.ptr(c_in), st))
    Nn.check(lib.adb_cl_wavenc_prep(Nn.ptr(net_in), Nn.ptr(b), B, cin, L, W, S, Nn.ptr(None), st))
    Nn.check_async()
    assert torch.equal(c_noise, c_noise2)
    assert torch.allclose(c_in, torch.full((B,), (0.7 ** 2 + 0.2 ** 2) ** -0.5, device=dev), rtol=1e-6)
    assert torch.equal(a, b)
