"""Deterministic weight factory for the DiffWave backbone (test infrastructure).

Produces a state_dict with exactly the reference's parameter names and shapes
(`src/models/backbones/wavenet.py:153-168`; names probed from `WaveNetNoise().state_dict()`:
`…conv.module.{bias,weight_g,weight_v}` for weight-normed convs, `output_projection.conv.*` for
the ZeroConv1d).  Values come from one seeded CPU generator walked in a fixed key order, so the
same (config, seed) gives bit-identical tensors in the build container (where the reference is
importable) and on the GPU box (where it is not).  The zero-initialised output projection
(`wavenet.py:57-66`) is deliberately given non-zero values: with the reference's own init every
network output is exactly 0 and parity would pass vacuously.
"""
from collections import OrderedDict
import math
import torch


def wavenet_param_shapes(residual_channels=256, residual_layers=36):
    C = residual_channels
    shapes = OrderedDict()

    def wn_conv(prefix, cout, cin, k):
        shapes[prefix + ".conv.module.bias"] = (cout,)
        shapes[prefix + ".conv.module.weight_g"] = ()
        shapes[prefix + ".conv.module.weight_v"] = (cout, cin, k)

    wn_conv("input_projection", C, 1, 1)
    shapes["residual_layer.fc_t1.weight"] = (512, 128)
    shapes["residual_layer.fc_t1.bias"] = (512,)
    shapes["residual_layer.fc_t2.weight"] = (512, 512)
    shapes["residual_layer.fc_t2.bias"] = (512,)
    for n in range(residual_layers):
        p = f"residual_layer.residual_blocks.{n}"
        wn_conv(p + ".dilated_conv", 2 * C, C, 3)
        shapes[p + ".diffusion_projection.weight"] = (C, 512)
        shapes[p + ".diffusion_projection.bias"] = (C,)
        wn_conv(p + ".output_projection", 2 * C, C, 1)
    wn_conv("skip_projection", C, C, 1)
    shapes["output_projection.conv.weight"] = (1, C, 1)
    shapes["output_projection.conv.bias"] = (1,)
    return shapes


def make_wavenet_state_dict(residual_channels=256, residual_layers=36, seed=0,
                            dtype=torch.float32):
    """Seeded state_dict with the reference's key set.

    Scale choices mimic the reference init so activations stay O(1): conv `weight_v` ~
    kaiming-normal (`wavenet.py:75`, std = sqrt(2 / fan_in)), `weight_g` = a value near
    ||v||_F but deliberately not equal to it (so the g/||v|| re-parameterisation of
    `wavenet.py:44-51` is exercised), biases and Linear weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    like torch's defaults.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd = OrderedDict()
    shapes = wavenet_param_shapes(residual_channels, residual_layers)
    for name, shape in shapes.items():
        if name.endswith("weight_v"):
            cout, cin, k = shape
            std = math.sqrt(2.0 / (cin * k))
            sd[name] = torch.randn(shape, generator=g, dtype=torch.float32) * std
        elif name.endswith("weight_g"):
            # weight_g precedes weight_v in key order: draw the factor now, fix up below
            sd[name] = torch.rand((), generator=g, dtype=torch.float32) * 0.4 + 0.8
        elif name == "output_projection.conv.weight":
            sd[name] = torch.randn(shape, generator=g, dtype=torch.float32) / math.sqrt(shape[1])
        elif name == "output_projection.conv.bias":
            sd[name] = torch.randn(shape, generator=g, dtype=torch.float32) * 0.05
        elif name.endswith(".weight"):          # nn.Linear weights
            bound = 1.0 / math.sqrt(shape[1])
            sd[name] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
        elif name.endswith(".bias"):
            bound = 0.1
            sd[name] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
        else:
            raise KeyError(name)
    # weight_g := factor * ||weight_v||_F  (factor in [0.8, 1.2])
    for name in list(sd.keys()):
        if name.endswith("weight_g"):
            v = sd[name[:-1] + "v"]
            sd[name] = (sd[name] * torch.linalg.vector_norm(v)).to(torch.float32)
    return OrderedDict((k, v.to(dtype)) for k, v in sd.items())


# --------------------------------------------------------------------------------------------------
# UNet1d (src/models/backbones/unet1d.py:624-893) — unconditional configuration
# --------------------------------------------------------------------------------------------------
UNET1D_CONFIG4 = dict(      # audio-diffusion-pytorch defaults in the reference's kwargs
    channels=128, cond_drop_prob=0.0, class_cond=False, text_cond=False, num_filters=128, window_length=32, stride=16,
    in_channels=2, resnet_groups=8, kernel_multiplier_downsample=2, multipliers=[1, 2, 4, 4, 4, 4, 4],
    factors=[4, 4, 4, 2, 2, 2], num_blocks=[2, 2, 2, 2, 2, 2], attentions=[False, False, False, True, True, True],
    attention_heads=8, attention_multiplier=2, use_nearest_upsample=False, use_skip_scale=True,
    use_attention_bottleneck=True)

UNET_SMALL = dict(channels=32, cond_drop_prob=0.0, class_cond=False, text_cond=False, num_filters=32, window_length=8, stride=4,
                  in_channels=2, resnet_groups=8, kernel_multiplier_downsample=2, multipliers=[1, 2, 2], factors=[4, 2],
                  num_blocks=[2, 1], attentions=[False, True], attention_heads=4, attention_multiplier=2,
                  use_nearest_upsample=False, use_skip_scale=True, use_attention_bottleneck=True)
UNET_MID = dict(channels=64, cond_drop_prob=0.0, class_cond=False, text_cond=False, num_filters=64, window_length=32, stride=16,
                in_channels=2, resnet_groups=8, kernel_multiplier_downsample=2, multipliers=[1, 2, 4], factors=[4, 2],
                num_blocks=[2, 1], attentions=[False, True], attention_heads=8, attention_multiplier=2,
                use_nearest_upsample=False, use_skip_scale=True, use_attention_bottleneck=True)
UNET_CLASS = dict(UNET_MID, class_cond=True, num_classes=10)     # label conditioning + classifier-free guidance
UNET_CASES = {   # name: (cfg, B, L, seed)
    "unet1d_small": (UNET_SMALL, 2, 256, 101),
    "unet1d_small_ragged": (UNET_SMALL, 3, 96, 102),         # L/stride/f0/f1 = 3 rows at the bottom: tiles far from full
    "unet1d_mid": (UNET_MID, 2, 16384, 103),
    "unet1d_cfg4_l65536": (UNET1D_CONFIG4, 1, 65536, 104),   # SURVEY §8(d) config 4 architecture at a quarter of the length
    "unet1d_cfg4_l262144": (UNET1D_CONFIG4, 1, 262144, 107),  # BASELINE.json configs[3] at its real length: 2 x 262144, B = 1
}




def unet1d_param_shapes(cfg):
    """Parameter names / shapes in the reference's state_dict order (probed from UNet1dBase(...).state_dict())."""
    ch, m = cfg["channels"], cfg["multipliers"]
    nf, W, cin = cfg["num_filters"], cfg["window_length"], cfg["in_channels"]
    cout = cfg.get("out_channels") or cin
    assert nf == ch * m[0], "to_in feeds the first down block: num_filters must equal channels * multipliers[0]"
    T = ch * 4
    am = cfg["attention_multiplier"]
    shapes = OrderedDict()
    cdim = ch * 4 if cfg.get("class_cond") else 0           # classes_channels (unet1d.py:843-850)
    if cfg.get("class_cond"):
        assert cfg.get("num_classes") is not None, "only the label (num_classes) conditioner is restated"
        shapes["label_conditioner.null_classes_emb"] = (1, ch)
        shapes["label_conditioner.label_emb.weight"] = (cfg["num_classes"], ch)
        shapes["label_conditioner.class_to_cond.0.weight"] = (ch,)
        shapes["label_conditioner.class_to_cond.0.bias"] = (ch,)
        shapes["label_conditioner.class_to_cond.1.weight"] = (cdim, ch)
        shapes["label_conditioner.class_to_cond.1.bias"] = (cdim,)
        shapes["label_conditioner.class_to_cond.3.weight"] = (cdim, cdim)
        shapes["label_conditioner.class_to_cond.3.bias"] = (cdim,)

    def resnet(p, ci, co):
        shapes[p + ".to_cond_embedding.1.weight"] = (2 * co, T + cdim)
        shapes[p + ".to_cond_embedding.1.bias"] = (2 * co,)
        shapes[p + ".block1.groupnorm.weight"] = (ci,)
        shapes[p + ".block1.groupnorm.bias"] = (ci,)
        shapes[p + ".block1.project.weight"] = (co, ci, 3)
        shapes[p + ".block1.project.bias"] = (co,)
        shapes[p + ".block2.groupnorm.weight"] = (co,)
        shapes[p + ".block2.groupnorm.bias"] = (co,)
        shapes[p + ".block2.project.weight"] = (co, co, 3)
        shapes[p + ".block2.project.bias"] = (co,)
        if ci != co:
            shapes[p + ".to_out.weight"] = (co, ci, 1)
            shapes[p + ".to_out.bias"] = (co,)

    def transformer(p, c):
        shapes[p + ".norm.weight"] = (c,)
        shapes[p + ".norm.bias"] = (c,)
        shapes[p + ".attention.to_q.weight"] = (c, c)
        shapes[p + ".attention.to_kv.weight"] = (2 * c, c)
        shapes[p + ".attention.to_out.weight"] = (c, c)
        shapes[p + ".feed_forward.0.g"] = (1, c, 1)
        shapes[p + ".feed_forward.1.weight"] = (c * am, c, 1)
        shapes[p + ".feed_forward.3.g"] = (1, c * am, 1)
        shapes[p + ".feed_forward.4.weight"] = (c, c * am, 1)

    shapes["unet.to_in.to_in.weight"] = (nf, cin, W)
    shapes["unet.to_out.to_out.weight"] = (nf, cout, W)
    shapes["unet.to_time.0.0.weights"] = (ch // 2,)
    shapes["unet.to_time.0.1.weight"] = (T, ch + 1)
    shapes["unet.to_time.0.1.bias"] = (T,)
    shapes["unet.to_time.2.weight"] = (T, T)
    shapes["unet.to_time.2.bias"] = (T,)
    n = len(m) - 1
    for i in range(n):
        ci, co, f = ch * m[i], ch * m[i + 1], cfg["factors"][i]
        p = f"unet.downsamples.{i}"
        shapes[p + ".downsample.weight"] = (co, ci, f * cfg["kernel_multiplier_downsample"] + 1)
        shapes[p + ".downsample.bias"] = (co,)
        for j in range(cfg["num_blocks"][i]):
            resnet(f"{p}.blocks.{j}", co, co)
        if cfg["attentions"][i]:
            transformer(p + ".transformer", co)
    cb = ch * m[-1]
    resnet("unet.bottleneck.pre_block", cb, cb)
    if cfg["use_attention_bottleneck"]:
        transformer("unet.bottleneck.transformer", cb)
    resnet("unet.bottleneck.post_block", cb, cb)
    for u, i in enumerate(reversed(range(n))):
        ci, co, f = ch * m[i + 1], ch * m[i], cfg["factors"][i]
        p = f"unet.upsamples.{u}"
        for j in range(cfg["num_blocks"][i] + (1 if cfg["attentions"][i] else 0)):
            resnet(f"{p}.blocks.{j}", 2 * ci, ci)
        if cfg["attentions"][i]:
            transformer(p + ".transformer", ci)
        if f == 1:
            shapes[p + ".upsample.weight"] = (co, ci, 3)
        else:
            shapes[p + ".upsample.weight"] = (ci, co, 2 * f)          # ConvTranspose1d weight is [Cin][Cout][k]
        shapes[p + ".upsample.bias"] = (co,)
    return shapes


def make_unet1d_state_dict(cfg, seed=0, dtype=torch.float32):
    """Seeded state_dict with the reference's key set. Conv / Linear weights ~ U(+-1/sqrt(fan_in)) (torch's
    default scale), norm gains near 1, biases small; the zero-initialised output transposed conv
    (unet1d.py:619) is given non-zero values so parity is not vacuous."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd = OrderedDict()

    def uni(shape, bound):
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound

    for name, shape in unet1d_param_shapes(cfg).items():
        if name.endswith("to_time.0.0.weights") or name.endswith("null_classes_emb") or name.endswith("label_emb.weight"):
            sd[name] = torch.randn(shape, generator=g, dtype=torch.float32)
        elif name.endswith("class_to_cond.0.weight"):
            sd[name] = 1.0 + uni(shape, 0.2)
        elif name.endswith(".g") or name.endswith("groupnorm.weight") or name.endswith("norm.weight"):
            sd[name] = 1.0 + uni(shape, 0.2)
        elif name.endswith(".bias"):
            sd[name] = uni(shape, 0.1)
        elif name.endswith("upsample.weight") and len(shape) == 3 and name.startswith("unet.upsamples") and \
                shape[2] != 3:
            sd[name] = uni(shape, 1.0 / math.sqrt(shape[0] * 2))      # each output sees 2 taps x Cin inputs
        elif name == "unet.to_out.to_out.weight":
            sd[name] = uni(shape, 1.0 / math.sqrt(shape[0] * 2))
        elif len(shape) == 3:
            sd[name] = uni(shape, 1.0 / math.sqrt(shape[1] * shape[2]))
        elif len(shape) == 2:
            sd[name] = uni(shape, 1.0 / math.sqrt(shape[1]))
        else:
            raise KeyError(name)
    return OrderedDict((k, v.to(dtype)) for k, v in sd.items())
