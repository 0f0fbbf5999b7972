"""EMA tracking of the network parameters behind the reference's class API (`src/models/phema.py`):
`PowerFunctionEMA` (:90-123, the power-function profiles of "Analyzing and Improving the Training Dynamics of
Diffusion Models") and `TraditionalEMA` (:126-160).

The reference updates every parameter tensor with its own `lerp_` (a Python loop of ~300 tiny launches per EMA per
step). Here each EMA copy keeps ONE flat fp32 vector and is updated with one `adb_ema_lerp` launch; `get()` writes
the flat vector back into a module copy with the reference's state_dict layout.
"""
import copy

import numpy as np
import torch

from . import _native as N


def std_to_exp(std):
    """Relative standard deviation -> power-function exponent (phema.py:28-33)."""
    std = np.float64(std)
    tmp = std.flatten() ** -2
    exp = [np.roots([1, 7, 16 - t, 12 - t]).real.max() for t in tmp]
    return np.float64(exp).reshape(std.shape)


def power_function_beta(std, t_next, t_delta):
    """phema.py:68-70."""
    return (1 - t_delta / t_next) ** (std_to_exp(std) + 1)


def _flat(net):
    return torch.cat([p.detach().reshape(-1).to(torch.float32) for p in net.parameters()]).contiguous()


def _flat_from_state_dict(net, state):
    """Flat fp32 vector in parameters() order from a module state_dict (what `state_dict()` of these classes holds)."""
    dev = next(net.parameters()).device
    return torch.cat([state[name].detach().reshape(-1).to(device=dev, dtype=torch.float32) for name, _ in net.named_parameters()]).contiguous()


def _lerp_(ema_flat, net, weight, flat_params=None):
    # `flat_params`: the live flat parameter vector when the trainer already keeps one (FusedTrainer.flat) — no re-concatenation
    src = flat_params if flat_params is not None else _flat(net)
    if not src.is_cuda:
        raise N.AdbError("EMA update needs the network on a B200 (no CPU path)")
    N.check(N.lib().adb_ema_lerp(N.ptr(ema_flat), N.ptr(src), float(weight), ema_flat.numel(), N.stream_ptr(src.device)))


def _materialise(net, ema_flat):
    out = copy.deepcopy(net)
    off = 0
    with torch.no_grad():
        for p in out.parameters():
            n = p.numel()
            p.copy_(ema_flat[off:off + n].view(p.shape))
            off += n
        for b_net, b_ema in zip(net.buffers(), out.buffers()):
            b_ema.copy_(b_net)
    return out


class PowerFunctionEMA:
    """phema.py:90-123: one EMA copy per relative standard deviation in `stds`."""

    @torch.no_grad()
    def __init__(self, net, stds=[0.050, 0.100], flat_params=None):
        self.net = net
        self.stds = stds
        self.flat_params = flat_params            # optional: the trainer's flat parameter vector, in parameters() order
        self.emas = [_flat(net) for _ in stds]

    @torch.no_grad()
    def reset(self):
        for e in self.emas:
            e.copy_(_flat(self.net))

    @torch.no_grad()
    def update(self, cur_nimg, batch_size):
        for std, e in zip(self.stds, self.emas):
            beta = power_function_beta(std=std, t_next=cur_nimg, t_delta=batch_size)
            _lerp_(e, self.net, 1 - float(beta), self.flat_params)

    @torch.no_grad()
    def get(self):
        return [(_materialise(self.net, e), f'-{std:.3f}') for std, e in zip(self.stds, self.emas)]

    def state_dict(self):
        return dict(stds=self.stds, emas=[m.state_dict() for m, _ in self.get()])

    @torch.no_grad()
    def load_state_dict(self, state):
        """phema.py:119-123: restore the profiles' relative standard deviations and every EMA copy (a checkpoint resume)."""
        self.stds = state['stds']
        flats = [_flat_from_state_dict(self.net, s_ema) for s_ema in state['emas']]
        if len(flats) != len(self.stds):
            raise ValueError(f"state holds {len(flats)} EMA copies for {len(self.stds)} stds")
        if len(flats) == len(self.emas):
            for e, f in zip(self.emas, flats):
                e.copy_(f)
        else:
            self.emas = flats


class TraditionalEMA:
    """phema.py:126-160: half-life EMA with ramp-up."""

    @torch.no_grad()
    def __init__(self, net, halflife_Mimg=float('inf'), rampup_ratio=0.09, flat_params=None):
        self.net = net
        self.halflife_Mimg = halflife_Mimg
        self.rampup_ratio = rampup_ratio
        self.flat_params = flat_params
        self.ema = _flat(net)

    @torch.no_grad()
    def reset(self):
        self.ema.copy_(_flat(self.net))

    @torch.no_grad()
    def update(self, cur_nimg, batch_size):
        halflife_Mimg = self.halflife_Mimg
        if self.rampup_ratio is not None:
            halflife_Mimg = min(halflife_Mimg, cur_nimg / 1e6 * self.rampup_ratio)
        beta = 0.5 ** (batch_size / max(halflife_Mimg * 1e6, 1e-8))
        _lerp_(self.ema, self.net, 1 - beta, self.flat_params)

    @torch.no_grad()
    def get(self):
        return _materialise(self.net, self.ema)

    def state_dict(self):
        return self.get().state_dict()

    @torch.no_grad()
    def load_state_dict(self, state):
        """phema.py:162-163."""
        self.ema.copy_(_flat_from_state_dict(self.net, state))
