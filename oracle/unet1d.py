"""Functional CPU restatement of the 1-D U-Net backbone (test infrastructure, see oracle/__init__).

Follows `src/models/backbones/unet1d.py` of the reference for the unconditional configuration
(class_cond = text_cond = False, no inj_embeddings / inj_channels, use_nearest_upsample = False):
  WAVenc1d / WAVdec1d            unet1d.py:572-622   strided conv k=W,s=S,p=W//2-S//2 / its transpose, no bias
  to_time                        unet1d.py:128-148, :678-684   [t, sin(2 pi t w), cos(2 pi t w)] -> Linear -> SiLU -> Linear
  ResnetBlock1d / ConvBlock1d    unet1d.py:163-207, :257-316   GN -> (scale+1, shift) -> SiLU -> conv k3, + to_out(x)
  Downsample1d / Upsample1d      unet1d.py:214-255   conv k=2f+1,s=f,p=f / ConvTranspose k=2f,s=f,p=f//2+f%2
  TransformerBlock1d             unet1d.py:67-122    x + Attn(LayerNorm(x)) in [b l c]; x + FF(x) in [b c l]
  Attention (self-attention)     attention_utils.py:78-184 (branch :157, :163-184) softmax in fp32, scale d^-1/2
  FeedForward1d / LayerNorm1d    unet1d.py:32-61     LN(no bias) -> 1x1 -> GELU -> LN(no bias) -> 1x1
  Down / Bottleneck / Up blocks  unet1d.py:323-570   skips per resnet block (+ transformer), cat([x, skip * 2^-1/2])
  UNet1d.forward                 unet1d.py:769-816 ; UNet1dBase.forward :856-893

`cfg` is a dict with the reference's constructor kwargs (UNet1dBase, unet1d.py:821-841 + UNet1d :625-648).
`sd` is a state_dict with the reference's parameter names (prefix "unet.").
"""
import math

import torch
import torch.nn.functional as F


def _p(sd, key):
    return sd["unet." + key]


def time_embedding(sd, t):
    """unet1d.py:128-148 + :678-684."""
    w = _p(sd, "to_time.0.0.weights")
    x = t[:, None]
    freqs = x * w[None, :] * 2 * math.pi
    f = torch.cat((x, freqs.sin(), freqs.cos()), dim=-1)
    h = F.linear(f, _p(sd, "to_time.0.1.weight"), _p(sd, "to_time.0.1.bias"))
    return F.linear(F.silu(h), _p(sd, "to_time.2.weight"), _p(sd, "to_time.2.bias"))


def conv_block(sd, prefix, x, groups, scale_shift=None):
    """ConvBlock1d (unet1d.py:163-207): GN -> optional x*(scale+1)+shift -> SiLU -> conv k3 pad 1."""
    x = F.group_norm(x, groups, _p(sd, prefix + ".groupnorm.weight"), _p(sd, prefix + ".groupnorm.bias"), eps=1e-5)
    if scale_shift is not None:
        x = x * (scale_shift[0] + 1) + scale_shift[1]
    x = F.silu(x)
    return F.conv1d(x, _p(sd, prefix + ".project.weight"), _p(sd, prefix + ".project.bias"), padding=1)


def label_embedding(sd, classes, cond_drop_prob):
    """LabelEmbedder.forward for integer labels (conditioner.py:59-111): cond_drop_prob = 0 keeps every label, 1 replaces
    every label by the learned null embedding (operator_utils.py:46-52); then LayerNorm -> Linear -> SiLU -> Linear."""
    emb = F.embedding(classes, sd["label_conditioner.label_emb.weight"])
    if cond_drop_prob >= 1:
        emb = sd["label_conditioner.null_classes_emb"].expand_as(emb)
    elif cond_drop_prob > 0:
        raise NotImplementedError("random label dropout (training) is not restated")
    C = emb.shape[-1]
    h = F.layer_norm(emb, (C,), sd["label_conditioner.class_to_cond.0.weight"], sd["label_conditioner.class_to_cond.0.bias"], eps=1e-5)
    h = F.silu(F.linear(h, sd["label_conditioner.class_to_cond.1.weight"], sd["label_conditioner.class_to_cond.1.bias"]))
    return F.linear(h, sd["label_conditioner.class_to_cond.3.weight"], sd["label_conditioner.class_to_cond.3.bias"])


def resnet_block(sd, prefix, x, temb, groups):
    """ResnetBlock1d (unet1d.py:257-316). `temb` is cat(time_embed, class_embed) when class-conditioned (:304-306)."""
    ce = F.linear(F.silu(temb), _p(sd, prefix + ".to_cond_embedding.1.weight"), _p(sd, prefix + ".to_cond_embedding.1.bias"))
    scale, shift = ce[:, :, None].chunk(2, dim=1)
    h = conv_block(sd, prefix + ".block1", x, groups)
    h = conv_block(sd, prefix + ".block2", h, groups, scale_shift=(scale, shift))
    if ("unet." + prefix + ".to_out.weight") in sd:
        x = F.conv1d(x, _p(sd, prefix + ".to_out.weight"), _p(sd, prefix + ".to_out.bias"))
    return h + x


def layer_norm_1d(x, g, eps=1e-5):
    """LayerNorm1d with bias=False (unet1d.py:32-45): over the channel dim of [b c l]."""
    var = torch.var(x, dim=1, unbiased=False, keepdim=True)
    mean = torch.mean(x, dim=1, keepdim=True)
    return (x - mean) * (var + eps).rsqrt() * g


def attention(sd, prefix, x, heads):
    """Self-attention branch of Attention.forward (attention_utils.py:117, :157, :163-184). x: [b n c]."""
    B, N, C = x.shape
    d = C // heads
    q = F.linear(x, _p(sd, prefix + ".to_q.weight"))
    k, v = F.linear(x, _p(sd, prefix + ".to_kv.weight")).chunk(2, dim=-1)
    q, k, v = (a.reshape(B, N, heads, d).transpose(1, 2) for a in (q, k, v))
    sim = torch.einsum("bhnd,bhmd->bhnm", q, k) * (d ** -0.5)
    attn = sim.softmax(dim=-1, dtype=torch.float32).to(sim.dtype)
    out = torch.einsum("bhnm,bhmd->bhnd", attn, v).transpose(1, 2).reshape(B, N, C)
    return F.linear(out, _p(sd, prefix + ".to_out.weight"))


def transformer_block(sd, prefix, x, heads):
    """TransformerBlock1d.forward with context=None (unet1d.py:106-122)."""
    C = x.shape[1]
    xt = x.transpose(1, 2)
    n = F.layer_norm(xt, (C,), _p(sd, prefix + ".norm.weight"), _p(sd, prefix + ".norm.bias"), eps=1e-5)
    xt = attention(sd, prefix + ".attention", n, heads) + xt
    x = xt.transpose(1, 2)
    h = layer_norm_1d(x, _p(sd, prefix + ".feed_forward.0.g"))
    h = F.conv1d(h, _p(sd, prefix + ".feed_forward.1.weight"))
    h = F.gelu(h)
    h = layer_norm_1d(h, _p(sd, prefix + ".feed_forward.3.g"))
    h = F.conv1d(h, _p(sd, prefix + ".feed_forward.4.weight"))
    return h + x


def unet1d_forward(sd, cfg, x, t, classes=None, cond_drop_prob=0.0):
    """UNet1dBase.forward -> UNet1d.forward (unet1d.py:856-893, :769-816). x: [B, in_channels, L], t: [B];
    classes: integer labels [B] when cfg["class_cond"] (unet1d.py:877)."""
    groups, heads = cfg["resnet_groups"], cfg["attention_heads"]
    factors, num_blocks, attentions = cfg["factors"], cfg["num_blocks"], cfg["attentions"]
    W, S = cfg["window_length"], cfg["stride"]
    n_levels = len(cfg["multipliers"]) - 1
    skip_scale = 2 ** -0.5 if cfg.get("use_skip_scale", False) else 1.0

    x = F.conv1d(x, _p(sd, "to_in.to_in.weight"), stride=S, padding=W // 2 - S // 2)
    temb = time_embedding(sd, t)
    if classes is not None:
        temb = torch.cat((temb, label_embedding(sd, classes, cond_drop_prob)), dim=-1)
    skips_list = []
    for i in range(n_levels):
        pre = f"downsamples.{i}"
        f = factors[i]
        km = cfg["kernel_multiplier_downsample"]
        x = F.conv1d(x, _p(sd, pre + ".downsample.weight"), _p(sd, pre + ".downsample.bias"), stride=f,
                     padding=f * (km // 2))
        skips = []
        for j in range(num_blocks[i]):
            x = resnet_block(sd, f"{pre}.blocks.{j}", x, temb, groups)
            skips.append(x)
        if attentions[i]:
            x = transformer_block(sd, pre + ".transformer", x, heads)
            skips.append(x)
        skips_list.append(skips)

    x = resnet_block(sd, "bottleneck.pre_block", x, temb, groups)
    if cfg.get("use_attention_bottleneck", False):
        x = transformer_block(sd, "bottleneck.transformer", x, heads)
    x = resnet_block(sd, "bottleneck.post_block", x, temb, groups)

    for u, i in enumerate(reversed(range(n_levels))):
        pre = f"upsamples.{u}"
        skips = skips_list.pop()
        f = factors[i]
        for j in range(num_blocks[i] + (1 if attentions[i] else 0)):
            x = torch.cat([x, skips.pop() * skip_scale], dim=1)
            x = resnet_block(sd, f"{pre}.blocks.{j}", x, temb, groups)
        if attentions[i]:
            x = transformer_block(sd, pre + ".transformer", x, heads)
        if f == 1:
            x = F.conv1d(x, _p(sd, pre + ".upsample.weight"), _p(sd, pre + ".upsample.bias"), padding=1)
        else:
            x = F.conv_transpose1d(x, _p(sd, pre + ".upsample.weight"), _p(sd, pre + ".upsample.bias"), stride=f,
                                   padding=f // 2 + f % 2, output_padding=f % 2)
    return F.conv_transpose1d(x, _p(sd, "to_out.to_out.weight"), stride=S, padding=W // 2 - S // 2)


def make_net_fn(sd, cfg):
    """net(x [B,C,L], c_noise [B], **kw) for oracle.edm.denoise (unconditional: kwargs ignored)."""
    return lambda x, t, classes=None, cond_drop_prob=0.0, **kw: unet1d_forward(sd, cfg, x, t, classes, cond_drop_prob)
