"""CPU check of the algebra behind the fused GroupNorm convolution (csrc/cl_ops.cuh GnCoefArgs, csrc/cl_conv_gn_tc.cuh): the
per-(sample, channel) slope / offset pairs the statistics kernel's last block derives from the fp64 group sums of the RAW
inputs reproduce torch's

    SiLU( GroupNorm(cat(x, s * skip)) * (scale + 1) + shift )            (unet1d.py:160-161, :195-207, :552-556)

without ever forming the concatenation or the scaled skip: the constant s is folded into the second input's statistics
(mean -> s mean, var -> s^2 var) and into its slope; the kernel stores the pairs halved and evaluates
SiLU(y) = h tanh(h) + h with h = y / 2. No GPU, no library call: this pins the formulas, the GPU tests pin the kernels."""
import torch
import torch.nn.functional as F


def coefficients(sums, count, gamma, beta, scale_shift, eps, s, c_off, C, G_this, Cin):
    """GnCoefArgs finalize, restated: sums [B][G_this][2] fp64 of ONE raw input with C channels at channel offset c_off."""
    B = sums.shape[0]
    mean = s * (sums[..., 0] / count)
    var = (s * s * (sums[..., 1] / count) - mean * mean).clamp_min(0.0)
    rstd = (1.0 / torch.sqrt(var + eps)).float()
    mean = mean.float()
    cpg = C // G_this
    g_of_c = torch.arange(C) // cpg
    a = rstd[:, g_of_c] * gamma[c_off:c_off + C]
    b = beta[c_off:c_off + C] - mean[:, g_of_c] * a
    if scale_shift is not None:
        sc = scale_shift[:, c_off:c_off + C] + 1.0
        sh = scale_shift[:, Cin + c_off:Cin + c_off + C]
        a = a * sc
        b = b * sc + sh
    return 0.5 * a * s, 0.5 * b                      # halved; the slope acts on the raw (unscaled) input


def test_folded_coefficients_reproduce_groupnorm_of_the_concatenation():
    gen = torch.Generator().manual_seed(0)
    B, L, C1, C2, G, eps, s = 3, 50, 64, 64, 8, 1e-5, 2 ** -0.5
    Cin = C1 + C2
    x = torch.randn(B, L, C1, generator=gen) * 1.7 + 0.3
    sk = torch.randn(B, L, C2, generator=gen) * 0.6 - 0.4
    gamma, beta = 1 + 0.3 * torch.randn(Cin, generator=gen), 0.2 * torch.randn(Cin, generator=gen)
    ss = 0.3 * torch.randn(B, 2 * Cin, generator=gen)
    cpg = Cin // G
    g1 = C1 // cpg
    assert C1 % cpg == 0                               # groups do not straddle the two inputs (what the host checks)

    def group_sums(t, groups):                         # what cl_gn_stats_vec_kernel accumulates, in fp64
        v = t.double().reshape(B, L, groups, -1)
        return torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1)

    a1, b1 = coefficients(group_sums(x, g1), cpg * L, gamma, beta, ss, eps, 1.0, 0, C1, g1, Cin)
    a2, b2 = coefficients(group_sums(sk, G - g1), cpg * L, gamma, beta, ss, eps, s, C1, C2, G - g1, Cin)
    h = torch.cat([x * a1[:, None, :] + b1[:, None, :], sk * a2[:, None, :] + b2[:, None, :]], dim=2)
    got = h * torch.tanh(h) + h                        # SiLU(2 h)

    cat = torch.cat([x, sk * s], dim=2).double()
    y = F.group_norm(cat.transpose(1, 2), G, gamma.double(), beta.double(), eps)
    y = y * (ss[:, :Cin].double().unsqueeze(2) + 1) + ss[:, Cin:].double().unsqueeze(2)
    want = F.silu(y).transpose(1, 2)
    err = float((got.double() - want).norm() / want.norm())
    assert err < 2e-6, err                             # fp32 coefficient arithmetic against an fp64 evaluation


def test_row_shifted_taps_are_the_convolution():
    """The three taps of the k = 3 'same' convolution as three views of ONE staged box with a one-row halo on either side (zero
    rows outside the sample): sum_tap box[tap : tap + rows] @ W[tap] == conv1d(padding = 1)."""
    gen = torch.Generator().manual_seed(1)
    L, Cin, N = 37, 16, 8
    x = torch.randn(L, Cin, generator=gen).double()
    w = torch.randn(3, Cin, N, generator=gen).double()
    box = torch.zeros(L + 2, Cin, dtype=torch.float64)
    box[1:L + 1] = x                                   # row r of the box is sample row r - 1 (TMA zero fill = the padding)
    got = sum(box[tap:tap + L] @ w[tap] for tap in range(3))
    want = F.conv1d(x.t().unsqueeze(0), w.permute(2, 1, 0), padding=1)[0].t()
    assert torch.allclose(got, want, atol=1e-12)
