"""Functional CPU restatement of the DiffWave backbone (test infrastructure, see oracle/__init__).

Follows `src/models/backbones/wavenet.py` of the reference:
  * weight re-parameterisation w = v * g / ||v||_F with a SCALAR g     (wavenet.py:44-51, :30)
  * sinusoidal step embedding + two swish MLP layers                  (wavenet.py:88-92, :139-141)
  * residual block: +Linear(512->C)(emb), dilated k=3 conv C->2C with zero "same" padding,
    sigmoid(first half) * tanh(second half), 1x1 conv C->2C, (x + first half)/sqrt(2), skip =
    second half                                                       (wavenet.py:107-115)
  * skip sum * sqrt(1/num_layers)                                      (wavenet.py:145-151)
  * input 1x1 conv + ReLU, skip 1x1 conv + ReLU, output 1x1 conv       (wavenet.py:170-180)
"""
import math
import torch
import torch.nn.functional as F


def fold_weight_norm(sd, prefix):
    """w = v * (g / ||v||_F); g is 0-dim (wavenet.py:30,50)."""
    v = sd[prefix + ".conv.module.weight_v"]
    g = sd[prefix + ".conv.module.weight_g"]
    return v * (g / torch.linalg.vector_norm(v)), sd[prefix + ".conv.module.bias"]


def step_embedding(t, dim_in=128):
    """[sin(t*f_j), cos(t*f_j)], f_j = exp(-4 j / (dim_in/2 - 1))  (wavenet.py:88-92)."""
    half = dim_in // 2
    j = torch.arange(half, dtype=t.dtype, device=t.device)
    arg = t[:, None] * torch.exp(-j * 4.0 / (half - 1))
    return torch.cat([arg.sin(), arg.cos()], dim=1)


def _swish(x):
    return x * torch.sigmoid(x)          # wavenet.py:84-86


def embed_mlp(sd, t):
    e = step_embedding(t)
    e = _swish(F.linear(e, sd["residual_layer.fc_t1.weight"], sd["residual_layer.fc_t1.bias"]))
    e = _swish(F.linear(e, sd["residual_layer.fc_t2.weight"], sd["residual_layer.fc_t2.bias"]))
    return e                              # [B, 512]


def residual_block(sd, n, dilation, x, emb):
    """One ResidualBlock (wavenet.py:107-115). x:[B,C,L], emb:[B,512] -> (h:[B,C,L], skip:[B,C,L])."""
    p = f"residual_layer.residual_blocks.{n}"
    C = x.shape[1]
    proj = F.linear(emb, sd[p + ".diffusion_projection.weight"], sd[p + ".diffusion_projection.bias"])
    y = x + proj[:, :, None]
    w1, b1 = fold_weight_norm(sd, p + ".dilated_conv")
    y = F.conv1d(y, w1, b1, dilation=dilation, padding=dilation)      # pad = d*(k-1)//2, wavenet.py:71
    gate, filt = y[:, :C], y[:, C:]                                   # torch.chunk order, :111
    z = torch.sigmoid(gate) * torch.tanh(filt)
    w2, b2 = fold_weight_norm(sd, p + ".output_projection")
    o = F.conv1d(z, w2, b2)
    return (x + o[:, :C]) / math.sqrt(2.0), o[:, C:]


def wavenet_forward(sd, audio, diffusion_step, dilation_cycle=12, return_intermediates=False):
    """WaveNetNoise.forward (wavenet.py:170-180): audio [B,L], diffusion_step [B] -> [B,1,L]."""
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("residual_layer.residual_blocks."))
    w_in, b_in = fold_weight_norm(sd, "input_projection")
    h = F.relu(F.conv1d(audio[:, None, :], w_in, b_in))
    emb = embed_mlp(sd, diffusion_step)
    skip = torch.zeros_like(h)
    inter = {}
    for n in range(n_layers):
        h, s = residual_block(sd, n, 2 ** (n % dilation_cycle), h, emb)
        skip = skip + s
        if return_intermediates:
            inter[f"h{n}"] = h
            inter[f"skip{n}"] = skip
    x = skip * math.sqrt(1.0 / n_layers)
    w_sp, b_sp = fold_weight_norm(sd, "skip_projection")
    x = F.relu(F.conv1d(x, w_sp, b_sp))
    x = F.conv1d(x, sd["output_projection.conv.weight"], sd["output_projection.conv.bias"])
    if return_intermediates:
        return x, inter
    return x


def make_net_fn(sd, dilation_cycle=12):
    """The adapter SURVEY.md §8(c) calls for: net(x[B,1,L], c_noise[B], **kw) -> [B,1,L]."""
    def net(x, t, **kwargs):
        return wavenet_forward(sd, x[:, 0, :], t, dilation_cycle=dilation_cycle)
    return net
