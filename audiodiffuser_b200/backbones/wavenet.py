"""DiffWave backbone behind the reference's class API, executed by the sm_100a kernels.

Mirrors `src/models/backbones/wavenet.py:153-180` (`WaveNetNoise`): same constructor, same
parameter names / shapes (`…conv.module.{bias,weight_g,weight_v}`, probed from the reference's
state_dict), same initialisation (kaiming-normal conv weights re-parameterised with a SCALAR
weight-norm gain :25-41, zero-initialised output conv :57-66), so a reference checkpoint loads with
`load_state_dict(strict=True)`.

`forward` accepts both the reference signature `(audio [B,L], diffusion_step [B]) -> [B,1,L]`
(wavenet.py:170) and the denoiser protocol `net(x [B,1,L], c_noise [B], cond_drop_prob=..., **kw)`
(diffusion.py:50), which the reference class itself does not.

precision = "bf16": tcgen05 implicit-GEMM residual blocks (bf16 operands, fp32 accumulate);
precision = "fp32": CUDA-core fp32 path. Both are CUDA kernels; there is no PyTorch fallback.
"""
import ctypes
import math
from ctypes import c_int, c_int64, c_void_p

import torch
import torch.nn as nn
from torch import Tensor

from .. import _native as N


class _Params(nn.Module):
    """Parameter holder named like the reference's weight-normed nn.Conv1d (`module.*`)."""

    def __init__(self, cout, cin, k):
        super().__init__()
        w = torch.empty(cout, cin, k)
        nn.init.kaiming_normal_(w)                                   # wavenet.py:75
        bound = 1.0 / math.sqrt(cin * k)
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))   # nn.Conv1d default
        g = torch.norm(w)                                            # wavenet.py:30 (scalar)
        self.weight_g = nn.Parameter(g.clone())
        self.weight_v = nn.Parameter(w / g)


class _WeightNormed(nn.Module):
    def __init__(self, cout, cin, k):
        super().__init__()
        self.module = _Params(cout, cin, k)


class _Conv(nn.Module):
    def __init__(self, cin, cout, k=3, dilation=1):
        super().__init__()
        self.dilation = dilation
        self.conv = _WeightNormed(cout, cin, k)


class _ZeroConv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, kernel_size=1, padding=0)  # holder only; zero init (:61-62)
        self.conv.weight.data.zero_()
        self.conv.bias.data.zero_()


class _ResidualBlock(nn.Module):
    def __init__(self, C, dilation):
        super().__init__()
        self.dilated_conv = _Conv(C, 2 * C, 3, dilation)
        self.diffusion_projection = nn.Linear(512, C)
        self.output_projection = _Conv(C, 2 * C, 1)


class _ResidualGroup(nn.Module):
    def __init__(self, C, layers, cycle):
        super().__init__()
        self.fc_t1 = nn.Linear(128, 512)
        self.fc_t2 = nn.Linear(512, 512)
        self.residual_blocks = nn.ModuleList([_ResidualBlock(C, 2 ** (n % cycle)) for n in range(layers)])


class WaveNetNoise(nn.Module):
    _adb_unconditional = True           # no conditioning input: guidance / conditioning kwargs cannot change the output

    def __init__(self, residual_channels: int = 256, residual_layers: int = 36, dilation_cycle: int = 12,
                 precision: str = "bf16"):
        super().__init__()
        if precision not in N.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(N.PRECISIONS)}")
        self.residual_channels = residual_channels
        self.residual_layers = residual_layers
        self.dilation_cycle = dilation_cycle
        self.precision = precision
        self.input_projection = _Conv(1, residual_channels, 1)
        self.residual_layer = _ResidualGroup(residual_channels, residual_layers, dilation_cycle)
        self.skip_projection = _Conv(residual_channels, residual_channels, 1)
        self.output_projection = _ZeroConv(residual_channels, 1)
        self._handle = None
        self._handle_key = None
        self._handle_dev = None
        self._ws = {}

    # ---- native handle management ----------------------------------------------------------------
    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def flat_parameters(self) -> Tensor:
        """state_dict values concatenated in state_dict order — the layout adb_wavenet_create expects."""
        return torch.cat([v.detach().reshape(-1).to(torch.float32) for v in self.state_dict().values()])

    def _native(self):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise N.AdbError("WaveNetNoise parameters are on the CPU: adb200 has no CPU path; call .cuda() on a B200")
        N.ensure_device(dev)
        key = self._param_key()
        if self._handle is None or key != self._handle_key:
            flat = self.flat_parameters().contiguous()
            lib = N.lib()
            expect = lib.adb_wavenet_param_count(self.residual_channels, self.residual_layers)
            if flat.numel() != expect:
                raise N.AdbError(f"flat parameter vector has {flat.numel()} values, library expects {expect}")
            with torch.cuda.device(dev):
                if self._handle is not None and self._handle_dev == dev:
                    # same configuration, new values (optimizer step / load_state_dict): refresh in place
                    torch.cuda.current_stream(dev).synchronize()
                    N.check(lib.adb_wavenet_load_params(self._handle, N.ptr(flat), flat.numel(), 1))
                    torch.cuda.synchronize(dev)
                else:
                    self._free()
                    h = c_void_p()
                    N.check(lib.adb_wavenet_create(ctypes.byref(h), self.residual_channels, self.residual_layers,
                                                   self.dilation_cycle, N.ptr(flat), flat.numel(), 1))
                    self._handle, self._handle_dev = h, dev
            self._handle_key = key
        return self._handle

    def parameters_updated(self):
        """Tell the module its parameters were changed outside torch's version tracking: the packed device weights are
        rebuilt on the next call. REQUIRED after writes through `.data` (`p.data.copy_`, `p.data.mul_`, an EMA weight swap via
        `.data`, a fused optimizer step on the flat vector) — those do not bump `p._version`, which is what `_param_key`
        watches; ordinary in-place ops, `load_state_dict` and optimizers do and need nothing."""
        self._handle_key = None

    def _free(self):
        if self._handle is not None:
            N.lib().adb_wavenet_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    def __getstate__(self):
        """copy.deepcopy / pickle (the reference deep-copies and pickles the net for its EMA snapshots,
        diffunet_complex_module.py:162-167, phema.py:96) carry the parameters only: the native handle and the
        scratch buffers belong to THIS object and are rebuilt lazily by the copy."""
        state = self.__dict__.copy()
        state.update(_handle=None, _handle_key=None, _handle_dev=None, _ws={})
        state.pop("_last_flat_grad", None)
        return state

    def _workspace(self, B, L, prec, device):
        """(aligned pointer, byte count) of a cached scratch buffer large enough for (B, L, precision)."""
        need = N.lib().adb_wavenet_workspace_bytes(self._native(), B, L, prec)
        key = (device, prec)
        ent = self._ws.get(key)
        if ent is None or ent[2] < need:
            if ent is not None:
                # growing: release the old buffer first (the z stash alone is ~75 GB at 256-sample passes; two of them do
                # not fit) and hand its block back to the driver so the larger request is not a second allocation
                self._ws.pop(key)
                del ent
                torch.cuda.empty_cache()
            buf, p = N.alloc_workspace(need, device)
            self._ws[key] = ent = (buf, p, need)
        return ent[1], need

    def _prec(self):
        return N.PRECISIONS[self.precision]

    # ---- reference-facing forward -----------------------------------------------------------------
    @torch.no_grad()
    def forward(self, audio: Tensor, diffusion_step: Tensor, **kwargs) -> Tensor:
        """(audio [B,L] or [B,1,L], diffusion_step [B]) -> [B,1,L]   (wavenet.py:170-180).
        Extra kwargs of the denoiser protocol (cond_drop_prob, ...) are accepted and ignored: the
        backbone is unconditional."""
        x = N.require_cuda_f32(audio, "audio")
        if x.ndim == 3:
            if x.shape[1] != 1:
                raise N.AdbError(f"WaveNetNoise expects mono audio [B,1,L]; got {tuple(x.shape)}")
            x = x[:, 0, :]
        B, L = x.shape
        t = N.require_cuda_f32(diffusion_step, "diffusion_step").reshape(B)
        h = self._native()
        prec = self._prec()
        ws, nbytes = self._workspace(B, L, prec, x.device)
        out = torch.empty(B, 1, L, dtype=torch.float32, device=x.device)
        N.check(N.lib().adb_wavenet_forward(h, N.ptr(x.contiguous()), N.ptr(t), N.ptr(None), 0, N.ptr(out), B, L, prec,
                                            ws, nbytes, N.stream_ptr(x.device)))
        return out

    @torch.no_grad()
    def forward_debug(self, audio: Tensor, diffusion_step: Tensor, dump_layers: int):
        """forward + (h, skip) after each of the first `dump_layers` blocks as fp32 [n][B][L][C]."""
        x = N.require_cuda_f32(audio, "audio")
        B, L = x.shape
        t = N.require_cuda_f32(diffusion_step, "diffusion_step").reshape(B)
        h = self._native()
        prec = self._prec()
        ws, nbytes = self._workspace(B, L, prec, x.device)
        out = torch.empty(B, 1, L, dtype=torch.float32, device=x.device)
        C = self.residual_channels
        dh = torch.zeros(dump_layers, B, L, C, dtype=torch.float32, device=x.device)
        ds = torch.zeros_like(dh)
        N.check(N.lib().adb_wavenet_forward_debug(h, N.ptr(x), N.ptr(t), N.ptr(None), 0, N.ptr(out), B, L, prec,
                                                  ws, nbytes, N.ptr(dh), N.ptr(ds), dump_layers,
                                                  N.stream_ptr(x.device)))
        return out, dh, ds

    # ---- fused hooks used by EluDiffusion / EDMSampler ---------------------------------------------
    def _adb_fused_denoise(self, x: Tensor, sig: Tensor, stride: int, sigma_data: float) -> Tensor:
        if x.ndim != 3 or x.shape[1] != 1:
            raise N.AdbError(f"fused DiffWave denoiser expects x [B,1,L]; got {tuple(x.shape)}")
        B, _, L = x.shape
        h = self._native()
        prec = self._prec()
        ws, nbytes = self._workspace(B, L, prec, x.device)
        out = torch.empty_like(x)
        N.check(N.lib().adb_wavenet_denoise(h, N.ptr(x), N.ptr(sig), stride, sigma_data, N.ptr(out), B, L, prec,
                                            ws, nbytes, N.stream_ptr(x.device)))
        return out

    def _adb_fused_sample(self, noise, sig, num_steps, sigma_data, s_tmin, s_tmax, s_churn, s_noise, use_heun, alpha,
                          eps, churn_seed=0, sample_offset=0):
        if noise.ndim != 3 or noise.shape[1] != 1:
            raise N.AdbError(f"fused DiffWave sampler expects noise [B,1,L]; got {tuple(noise.shape)}")
        B, _, L = noise.shape
        h = self._native()
        prec = self._prec()
        ws, nbytes = self._workspace(B, L, prec, noise.device)
        out = torch.empty_like(noise)
        sig_arr = (ctypes.c_float * len(sig))(*sig)
        nfe = c_int(0)
        if eps is not None:
            eps = N.require_cuda_f32(eps, "eps")
            if eps.numel() != num_steps * noise.numel():
                raise N.AdbError(f"eps must have shape [num_steps, *noise.shape]; got {tuple(eps.shape)}")
        N.check(N.lib().adb_wavenet_sample_edm_seeded(h, N.ptr(noise), sig_arr, len(sig), num_steps, sigma_data, s_tmin, s_tmax,
                                                      s_churn, s_noise, int(use_heun), alpha, N.ptr(eps), int(churn_seed),
                                                      int(sample_offset), N.ptr(out), B, L, prec, ws, nbytes,
                                                      ctypes.byref(nfe), N.stream_ptr(noise.device)))
        return out, nfe.value

    # ---- training step (fused DSM loss forward / backward; fp32 kernels) -----------------------------
    def _adb_dsm_loss(self, x: Tensor, sigmas: Tensor, noise: Tensor, sigma_data: float) -> Tensor:
        """loss [B] of Diffusion.forward (diffusion.py:65-97) as an autograd node whose backward runs the CUDA
        backward pass and hands every parameter its gradient (so `loss.mean().backward(); optimizer.step()` and
        DDP's gradient hooks work unchanged)."""
        if x.ndim != 3 or x.shape[1] != 1:
            raise N.AdbError(f"fused DiffWave training step expects x [B,1,L]; got {tuple(x.shape)}")
        return _DsmLossFn.apply(x, sigmas, noise, float(sigma_data), self, *self.parameters())

    # ---- kernel-class timers (bench) ---------------------------------------------------------------
    def set_timing(self, enabled: bool):
        N.check(N.lib().adb_wavenet_set_timing(self._native(), int(enabled)))

    def timers(self):
        ms = (ctypes.c_double * len(N.TIMER_NAMES))()
        cnt = (c_int64 * len(N.TIMER_NAMES))()
        N.check(N.lib().adb_wavenet_timers(self._native(), ms, cnt))
        return {n: (ms[i], cnt[i]) for i, n in enumerate(N.TIMER_NAMES)}


class _DsmLossFn(torch.autograd.Function):
    """Diffusion.forward through the fused backbone: forward = adb_wavenet_dsm_forward_train (activations kept in a
    workspace tensor), backward = adb_wavenet_dsm_backward (flat gradient in state_dict order, split into views)."""

    @staticmethod
    def forward(ctx, x, sigmas, noise, sigma_data, net, *params):
        lib = N.lib()
        B, _, L = x.shape
        h = net._native()
        prec = net._prec()
        nbytes = lib.adb_wavenet_train_workspace_bytes(h, B, L, prec)
        ws_buf, ws = N.alloc_workspace(nbytes, x.device)
        loss = torch.empty(B, dtype=torch.float32, device=x.device)
        N.check(lib.adb_wavenet_dsm_forward_train(h, N.ptr(x), N.ptr(noise), N.ptr(sigmas), sigma_data, N.ptr(loss), B, L,
                                                  prec, ws, nbytes, N.stream_ptr(x.device)))
        ctx.net, ctx.ws, ctx.ws_buf, ctx.nbytes, ctx.sigma_data, ctx.prec = net, ws, ws_buf, nbytes, sigma_data, prec
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.save_for_backward(x, sigmas)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        x, sigmas = ctx.saved_tensors
        lib, net = N.lib(), ctx.net
        B, _, L = x.shape
        n_params = sum(math.prod(s) for s in ctx.shapes)
        flat = torch.empty(n_params, dtype=torch.float32, device=x.device)
        up = N.require_cuda_f32(grad_loss, "grad_loss").reshape(B)
        N.check(lib.adb_wavenet_dsm_backward(net._native(), N.ptr(x), N.ptr(sigmas), ctx.sigma_data, N.ptr(up), N.ptr(flat), B, L,
                                             ctx.prec, ctx.ws, ctx.nbytes, N.stream_ptr(x.device)))
        ctx.ws = ctx.ws_buf = None
        grads, off = [], 0
        for shp in ctx.shapes:
            n = math.prod(shp)
            grads.append(flat[off:off + n].view(shp))
            off += n
        net._last_flat_grad = flat
        return (None, None, None, None, None, *grads)


class EDMDenoiser(nn.Module):
    """The "(x, sigma) denoiser module" of BASELINE.json's north_star:
    D(x, sigma) == diffusion.denoise_fn(x, net=net, sigma=sigma, inference=True)   (diffusion.py:32-63)."""

    def __init__(self, net: nn.Module, diffusion: nn.Module):
        super().__init__()
        self.net = net
        self.diffusion = diffusion

    @torch.no_grad()
    def forward(self, x: Tensor, sigma, **kwargs) -> Tensor:
        if isinstance(sigma, Tensor) and sigma.ndim == 1 and sigma.numel() == x.shape[0] and sigma.numel() > 1:
            return self.diffusion.denoise_fn(x, net=self.net, sigmas=sigma, inference=True, **kwargs)
        return self.diffusion.denoise_fn(x, net=self.net, sigma=sigma, inference=True, **kwargs)
