"""DPM-Solver and UniPC samplers behind the reference's class API.

Mirrors `src/models/components/sampler_edm.py`: `DPMSampler` (:495-805 — single-step DPM-Solver-1/2/3 with the "fast"
order schedule, and multistep DPM-Solver / DPM-Solver++ of order 1-3) and `UniPCSampler` (:807-1053 — UniPC
predictor-corrector, variant B(h) = expm1(h)). Same constructors, same `forward(noise, fn, net, sigmas, **kwargs)`.

Every update of these solvers is `a x + sum_k c_k m_k`: the state plus up to four stored network outputs, with scalars
that depend only on the time grid. The reference evaluates them as chains of 6-15 elementwise torch kernels over full
tensors with the scalars living in 0-dim device tensors (one host sync whenever one is inspected); here the scalars are
folded on the host in double precision from the fp32 time grid and each update is ONE launch of `adb_edm_lincomb_n`
((K + 2) x 4 bytes per element), the final `clamp(-1, 1)` (:805, :1053) fused into the last one. The denoiser calls go
through `fn` (any denoiser; with this package's `EluDiffusion.denoise_fn` and a fused backbone they are the fused CUDA
path). Differences from the reference, on purpose: `**kwargs` (e.g. `classes`) reach EVERY denoiser call (the reference's
single-step solvers drop them for the intermediate evaluations, :578, :598-603), and UniPC accepts states of any rank
(the reference's einsum `k,bkchw->bchw`, :931, only 4-D ones).
"""
import ctypes
import math
from typing import Callable, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from .. import _native as N


def _f32(v: float) -> float:
    return ctypes.c_float(v).value


def lincomb(x: Tensor, a: float, terms: Sequence[Tuple[float, Tensor]], clamp: bool = False) -> Tensor:
    """a x + sum c_k m_k in one launch (<= 4 terms)."""
    k = len(terms)
    x = N.require_cuda_f32(x, "x")
    for _, m in terms:
        N.require_cuda_f32(m, "term")
        if m.shape != x.shape or not m.is_contiguous() or not x.is_contiguous():
            raise N.AdbError("lincomb operands must be contiguous and shaped like x")
    out = torch.empty_like(x)
    ptrs = (ctypes.c_void_p * max(k, 1))(*[m.data_ptr() for _, m in terms])
    coefs = (ctypes.c_float * max(k, 1))(*[float(c) for c, _ in terms])
    N.check(N.lib().adb_edm_lincomb_n(N.ptr(x), float(a), ptrs, coefs, k, int(clamp), N.ptr(out), x.numel(),
                                      N.stream_ptr(x.device)))
    return out


class _Grid:
    """Host copy of the time grid with the reference's three conversions (:521-551, :849-868)."""

    def __init__(self, sigmas: Tensor, points: int, log_time_spacing: bool):
        self.log = log_time_spacing
        s = sigmas.detach().to(torch.float32).cpu()
        if log_time_spacing:
            t = torch.linspace(-s[0].log(), -s[-1].log(), points + 1)           # fp32, like the reference (:544-546)
        else:
            t = s
        self.t: List[float] = [float(v) for v in t.tolist()]

    def lam(self, t: float) -> float:
        return t if self.log else -math.log(t)

    def sig(self, t: float) -> float:
        return math.exp(-t) if self.log else t

    def inv(self, v: float) -> float:
        return v if self.log else math.exp(-v)


class _SolverBase(nn.Module):
    def _model(self, fn, net, g: _Grid, x: Tensor, t: float, **kwargs) -> Tensor:
        """model_fn (:706-723, :831-847): D(x; sigma), or the noise prediction (x - D) / sigma in one launch."""
        sigma = _f32(g.sig(t))
        d = N.require_cuda_f32(fn(x, net=net, sigma=sigma, inference=True, cond_scale=self.cond_scale, **kwargs), "denoised")
        self.last_nfe += 1
        if self.x0_pred:
            return d.contiguous()
        return lincomb(x, 1.0 / sigma, [(-1.0 / sigma, d.contiguous())])

    def _start(self, noise: Tensor, sigmas: Tensor) -> Tensor:
        noise = N.require_cuda_f32(noise, "noise").contiguous()
        N.ensure_device(noise.device)
        self.last_nfe = 0
        return lincomb(noise, _f32(float(sigmas[0])), [])                       # x = sigmas[0] * noise


class DPMSampler(_SolverBase):
    """sampler_edm.py:495-805."""

    def __init__(self, cond_scale, order=1, num_steps=10, multisteps=False, x0_pred: bool = True,
                 log_time_spacing: bool = True):
        super().__init__()
        self.order = order
        self.cond_scale = cond_scale
        self.multisteps = multisteps
        self.x0_pred = x0_pred
        self.log_time_spacing = log_time_spacing
        self.num_steps = num_steps if self.log_time_spacing else num_steps - 1   # :514
        self.last_nfe = 0

    # ---- single-step solvers (:562-630) ------------------------------------------------------------------------------
    def _orders(self) -> List[int]:
        n = self.num_steps
        if self.order == 3:                                                        # :776-781
            k = n // 3 + 1
            return [3] * (k - 2) + [2, 1] if n % 3 == 0 else [3] * (k - 1) + [n % 3]
        if self.order == 2:                                                        # :783-789
            return [2] * (n // 2) if n % 2 == 0 else [2] * (n // 2) + [1]
        if self.order == 1:
            return [1] * n
        raise ValueError("'order' must be '1' or '2' or '3'.")

    def _single(self, fn, net, g, x, tc, tn, order, clamp, **kw):
        h = g.lam(tn) - g.lam(tc)
        sc, sn = g.sig(tc), g.sig(tn)
        x0 = self.x0_pred
        e = self._model(fn, net, g, x, tc, **kw)
        a_n = sn / sc if x0 else 1.0                                               # coefficient of x in the final update
        e_n = -math.expm1(-h) if x0 else -sn * math.expm1(h)                       # first-order coefficient of eps
        if order == 1:
            return lincomb(x, a_n, [(e_n, e)], clamp)
        r1 = 0.5 if order == 2 else 1.0 / 3.0
        s1 = g.inv(tc + r1 * h)                                                    # sic: tc is a SIGMA without log spacing (:592-594)
        if x0:
            u1 = lincomb(x, g.sig(s1) / sc, [(-math.expm1(-r1 * h), e)])
        else:
            u1 = lincomb(x, 1.0, [(-g.sig(s1) * math.expm1(r1 * h), e)])
        e1 = self._model(fn, net, g, u1, s1, **kw)
        if order == 2:
            k2 = e_n / (2 * r1)                                                    # -(1/2r1) expm1(-h) resp. -(sn/2r1) expm1(h)
            return lincomb(x, a_n, [(e_n - k2, e), (k2, e1)], clamp)
        r2 = 2.0 / 3.0
        s2 = g.inv(tc + r2 * h)
        if x0:
            k = (r2 / r1) * (math.expm1(-r2 * h) / (r2 * h) + 1)
            u2 = lincomb(x, g.sig(s2) / sc, [(-math.expm1(-r2 * h) - k, e), (k, e1)])
            k3 = (1 / r2) * (math.expm1(-h) / h + 1)
        else:
            k = -g.sig(s2) * (r2 / r1) * (math.expm1(r2 * h) / (r2 * h) - 1)
            u2 = lincomb(x, 1.0, [(-g.sig(s2) * math.expm1(r2 * h) - k, e), (k, e1)])
            k3 = -sn / r2 * (math.expm1(h) / h - 1)
        e2 = self._model(fn, net, g, u2, s2, **kw)
        return lincomb(x, a_n, [(e_n - k3, e), (k3, e2)], clamp)

    # ---- multistep solvers (:632-704) ----------------------------------------------------------------------------------
    def _multi(self, g, x, ms, ts, tc, order, clamp):
        m0, t0 = ms[-1], ts[-1]
        h = g.lam(tc) - g.lam(t0)
        sc = g.sig(tc)
        x0 = self.x0_pred
        a = sc / g.sig(t0) if x0 else 1.0
        w = 1.0 if x0 else sc                                                      # the eps form scales every phi by sigma(t)
        phi1 = math.expm1(-h) if x0 else math.expm1(h)
        if order == 1:
            return lincomb(x, a, [(-w * phi1, m0)], clamp)
        if order == 2:
            r0 = (g.lam(t0) - g.lam(ts[-2])) / h
            k = 0.5 * w * phi1 / r0
            return lincomb(x, a, [(-w * phi1 - k, m0), (k, ms[-2])], clamp)
        t2, t1, _ = ts
        m2, m1, _ = ms
        r0, r1 = (g.lam(t0) - g.lam(t1)) / h, (g.lam(t1) - g.lam(t2)) / h
        if x0:
            phi2 = phi1 / h + 1.0
            phi3 = phi2 / h - 0.5
            p1, pd1, pd2 = -phi1, phi2, -phi3
        else:
            phi2 = phi1 / h - 1.0
            phi3 = phi2 / h - 0.5
            p1, pd1, pd2 = -sc * phi1, -sc * phi2, -sc * phi3
        beta, gamma = r0 / (r0 + r1), 1.0 / (r0 + r1)
        q0 = pd1 * (1.0 + beta) + pd2 * gamma                                      # weight of D1_0 = (m0 - m1) / r0
        q1 = -pd1 * beta - pd2 * gamma                                             # weight of D1_1 = (m1 - m2) / r1
        return lincomb(x, a, [(p1 + q0 / r0, m0), (-q0 / r0 + q1 / r1, m1), (-q1 / r1, m2)], clamp)

    @torch.no_grad()
    def forward(self, noise: Tensor, fn: Callable, net: nn.Module, sigmas: Tensor, **kwargs) -> Tensor:
        x = self._start(noise, sigmas)
        n = self.num_steps
        if not self.multisteps:
            orders = self._orders()
            g = _Grid(sigmas, len(orders), self.log_time_spacing)
            if len(g.t) < len(orders) + 1:
                raise IndexError(f"schedule has {len(g.t)} sigmas, the solver needs {len(orders) + 1}")
            for i, o in enumerate(orders):
                x = self._single(fn, net, g, x, g.t[i], g.t[i + 1], o, i == len(orders) - 1, **kwargs)
            return x
        assert n >= self.order                                                     # :735
        if self.order not in (1, 2, 3):
            raise ValueError("'order' must be '1' or '2' or '3'.")
        g = _Grid(sigmas, n, self.log_time_spacing)
        if len(g.t) < n + 1:
            raise IndexError(f"schedule has {len(g.t)} sigmas, the solver needs {n + 1}")
        ms, ts = [self._model(fn, net, g, x, g.t[0], **kwargs)], [g.t[0]]
        for step in range(1, self.order):                                          # warm-up with increasing order (:746-760)
            x = self._multi(g, x, ms, ts, g.t[step], step, False)
            ts.append(g.t[step])
            ms.append(self._model(fn, net, g, x, g.t[step], **kwargs))
        for step in range(self.order, n + 1):                                      # :763-784
            x = self._multi(g, x, ms, ts, g.t[step], min(self.order, n + 1 - step), step == n)
            ts = ts[1:] + [g.t[step]]
            if step < n:                                                           # the final model value is not needed
                ms = ms[1:] + [self._model(fn, net, g, x, g.t[step], **kwargs)]
        return x


class UniPCSampler(_SolverBase):
    """sampler_edm.py:807-1053 (variant 'bh2'); `num_steps` denoiser calls."""

    def __init__(self, num_steps: int = 20, order: int = 2, cond_scale: float = 1.0, x0_pred: bool = True,
                 log_time_spacing: bool = True):
        super().__init__()
        self.order = order
        self.cond_scale = cond_scale
        self.x0_pred = x0_pred
        self.log_time_spacing = log_time_spacing
        self.num_steps = num_steps if self.log_time_spacing else num_steps - 1     # :830
        self.last_nfe = 0

    def _update(self, fn, net, g, x, ms, ts, tc, order, use_corrector, clamp, **kw):
        """multistep_uni_pc_update (:870-987): predictor, one denoiser call at the predicted point, corrector."""
        assert order <= len(ms)
        t0, m0 = ts[-1], ms[-1]
        h = g.lam(tc) - g.lam(t0)
        rks = [(g.lam(ts[-(i + 1)]) - g.lam(t0)) / h for i in range(1, order)]
        hist = [ms[-(i + 1)] for i in range(1, order)]
        hh = -h if self.x0_pred else h
        h_phi_1 = math.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        b_h = math.expm1(hh)                                                       # variant 'bh2'
        rk = np.array(rks + [1.0], dtype=np.float64)
        R, b, fact = [], [], 1
        for i in range(1, order + 1):
            R.append(rk ** (i - 1))
            b.append(h_phi_k * fact / b_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        R, b = np.stack(R), np.array(b)
        sc = g.sig(tc)
        a = sc / g.sig(t0) if self.x0_pred else 1.0
        w = 1.0 if self.x0_pred else sc
        scale = w * b_h

        def combo(rhos, extra):
            """base - scale (sum_k rho_k (m_k - m0) / rk_k + extra_rho (m_t - m0)) as coefficients of m0, m_k[, m_t]."""
            c0 = -w * h_phi_1 + scale * (sum(r / q for r, q in zip(rhos, rks)) + (extra[0] if extra else 0.0))
            terms = [(c0, m0)] + [(-scale * r / q, m) for r, q, m in zip(rhos, rks, hist)]
            if extra:
                terms.append((-scale * extra[0], extra[1]))
            return terms

        rhos_p = [] if order == 1 else ([0.5] if order == 2 else list(np.linalg.solve(R[:-1, :-1], b[:-1])))
        x_t = lincomb(x, a, combo(rhos_p, None), clamp and not use_corrector)
        if not use_corrector:
            return x_t, None
        rhos_c = [0.5] if order == 1 else list(np.linalg.solve(R, b))
        model_t = self._model(fn, net, g, x_t, tc, **kw)
        return lincomb(x, a, combo(rhos_c[:-1], (rhos_c[-1], model_t)), clamp), model_t

    @torch.no_grad()
    def forward(self, noise: Tensor, fn: Callable, net: nn.Module, sigmas: Tensor, **kwargs) -> Tensor:
        n = self.num_steps
        assert n >= self.order                                                     # :996
        if self.order not in (1, 2, 3):
            raise ValueError("UniPC order must be 1, 2 or 3 (one launch combines at most four network outputs)")
        x = self._start(noise, sigmas)
        g = _Grid(sigmas, n, self.log_time_spacing)
        if len(g.t) < n + 1:
            raise IndexError(f"schedule has {len(g.t)} sigmas, the solver needs {n + 1}")
        ms, ts = [self._model(fn, net, g, x, g.t[0], **kwargs)], [g.t[0]]
        for step in range(1, self.order):                                          # :1008-1018
            x, m = self._update(fn, net, g, x, ms, ts, g.t[step], step, True, False, **kwargs)
            ts.append(g.t[step])
            ms.append(m)
        for step in range(self.order, n + 1):                                      # :1021-1051
            x, m = self._update(fn, net, g, x, ms, ts, g.t[step], min(self.order, n + 1 - step), step != n, step == n,
                                **kwargs)
            ts = ts[1:] + [g.t[step]]
            if step < n:
                ms = ms[1:] + [m]
        return x
