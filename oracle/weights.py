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
