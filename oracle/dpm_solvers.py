"""CPU restatement of the reference's DPM-Solver and UniPC samplers (test infrastructure only, see oracle/__init__).

Follows `src/models/components/sampler_edm.py`:
  DPMSampler      :495-805   single-step DPM-Solver-1/2/3 ("fast" order schedule) and multistep DPM-Solver(++) of order 1-3,
                             data-prediction (x0_pred) or noise-prediction form, log-SNR or given-sigma time spacing
  UniPCSampler    :807-1053  UniPC predictor-corrector (variant B(h) = expm1(h)) on the multistep history

In the EDM parameterisation alpha = 1, lambda = -log sigma. With `log_time_spacing` the time grid is a linspace in lambda;
without it the grid is the caller's sigma list, and the reference's single-step solvers then add a lambda-space increment to
a SIGMA-space time (:592-594, :612-616) — restated as is, because parity is against the reference's results.

`denoise_fn(x, sigma)` is D(x; sigma) (diffusion.denoise_fn with inference=True); sigma is passed as a 0-dim fp32 tensor like the
reference does. Tensors may have any rank (the reference's UniPC einsum `k,bkchw->bchw` (:931) only accepts 4-D states; the
weighted sum below is the same contraction for every rank). Pinned by tests/golden/dpm_unipc_small.npz, generated from
the reference's classes by oracle/make_golden_dpm.py.
"""
import torch


class _Grid:
    """Time grid + the three conversions of the reference (DPMSampler.lambd/sigma/inv_lambd :521-538,
    UniPCSampler.sigma_to_lambd/lambda_to_sigma :849-859)."""

    def __init__(self, sigmas, points, log_time_spacing):
        self.log = log_time_spacing
        if log_time_spacing:
            self.t = torch.linspace(-sigmas[0].log(), -sigmas[-1].log(), points + 1)     # :544-546
        else:
            self.t = sigmas                                                                # :550

    def lam(self, t):
        return t if self.log else -t.log()

    def sig(self, t):
        return t.neg().exp() if self.log else t

    def inv(self, lam_like):
        return lam_like if self.log else lam_like.neg().exp()


def _model(denoise_fn, grid, x, t, x0_pred):
    """model_fn (:706-723, :831-847): the denoised sample, or the noise prediction (x - D) / sigma."""
    d = denoise_fn(x, grid.sig(t))
    return d if x0_pred else (x - d) / grid.sig(t)


def _singlestep_orders(order, num_steps):
    """The "DPM-Solver-fast" order schedule (:776-791)."""
    if order == 3:
        k = num_steps // 3 + 1
        return [3] * (k - 2) + [2, 1] if num_steps % 3 == 0 else [3] * (k - 1) + [num_steps % 3]
    if order == 2:
        return [2] * (num_steps // 2) if num_steps % 2 == 0 else [2] * (num_steps // 2) + [1]
    if order == 1:
        return [1] * num_steps
    raise ValueError("'order' must be '1' or '2' or '3'.")


def _single(denoise_fn, g, x, tc, tn, order, x0):
    """dpm_solver_{1,2,3}_step (:562-630)."""
    h = g.lam(tn) - g.lam(tc)
    e = _model(denoise_fn, g, x, tc, x0)
    sc, sn = g.sig(tc), g.sig(tn)
    if order == 1:
        return sn / sc * x - torch.expm1(-h) * e if x0 else x - sn * h.expm1() * e
    if order == 2:
        r1 = 0.5
        s1 = g.inv(tc + r1 * h)
        if x0:
            u1 = g.sig(s1) / sc * x - torch.expm1(-r1 * h) * e
            e1 = _model(denoise_fn, g, u1, s1, x0)
            return sn / sc * x - torch.expm1(-h) * e - 1 / (2 * r1) * torch.expm1(-h) * (e1 - e)
        u1 = x - g.sig(s1) * (r1 * h).expm1() * e
        e1 = _model(denoise_fn, g, u1, s1, x0)
        return x - sn * h.expm1() * e - sn / (2 * r1) * h.expm1() * (e1 - e)
    r1, r2 = 1 / 3, 2 / 3
    s1, s2 = g.inv(tc + r1 * h), g.inv(tc + r2 * h)
    if x0:
        u1 = g.sig(s1) / sc * x - (-r1 * h).expm1() * e
        e1 = _model(denoise_fn, g, u1, s1, x0)
        u2 = g.sig(s2) / sc * x - (-r2 * h).expm1() * e + (r2 / r1) * ((-r2 * h).expm1() / (r2 * h) + 1) * (e1 - e)
        e2 = _model(denoise_fn, g, u2, s2, x0)
        return sn / sc * x - torch.expm1(-h) * e + 1 / r2 * (torch.expm1(-h) / h + 1) * (e2 - e)
    u1 = x - g.sig(s1) * (r1 * h).expm1() * e
    e1 = _model(denoise_fn, g, u1, s1, x0)
    u2 = x - g.sig(s2) * (r2 * h).expm1() * e - g.sig(s2) * (r2 / r1) * ((r2 * h).expm1() / (r2 * h) - 1) * (e1 - e)
    e2 = _model(denoise_fn, g, u2, s2, x0)
    return x - sn * h.expm1() * e - sn / r2 * (h.expm1() / h - 1) * (e2 - e)


def _multi(g, x, ms, ts, tc, order, x0):
    """multistep_dpm_solver_{1,2,3}_step (:632-704) on the history ms / ts (oldest first)."""
    m0, t0 = ms[-1], ts[-1]
    h = g.lam(tc) - g.lam(t0)
    sc = g.sig(tc)
    phi1 = torch.expm1(-h) if x0 else torch.expm1(h)
    if order == 1:
        return sc / g.sig(t0) * x - phi1 * m0 if x0 else x - sc * phi1 * m0
    if order == 2:
        r0 = (g.lam(t0) - g.lam(ts[-2])) / h
        d10 = (1.0 / r0) * (m0 - ms[-2])
        if x0:
            return sc / g.sig(t0) * x - phi1 * m0 - 0.5 * phi1 * d10
        return x - (sc * phi1) * m0 - 0.5 * (sc * phi1) * d10
    t2, t1, _ = ts
    m2, m1, _ = ms
    r0, r1 = (g.lam(t0) - g.lam(t1)) / h, (g.lam(t1) - g.lam(t2)) / h
    d10, d11 = (1.0 / r0) * (m0 - m1), (1.0 / r1) * (m1 - m2)
    d1 = d10 + (r0 / (r0 + r1)) * (d10 - d11)
    d2 = (1.0 / (r0 + r1)) * (d10 - d11)
    if x0:
        phi2 = phi1 / h + 1.0
        phi3 = phi2 / h - 0.5
        return sc / g.sig(t0) * x - phi1 * m0 + phi2 * d1 - phi3 * d2
    phi2 = phi1 / h - 1.0
    phi3 = phi2 / h - 0.5
    return x - (sc * phi1) * m0 - (sc * phi2) * d1 - (sc * phi3) * d2


def dpm_sampler(noise, denoise_fn, sigmas, order=1, num_steps=10, multisteps=False, x0_pred=True, log_time_spacing=True):
    """DPMSampler.forward (:725-805). `num_steps` is the constructor argument (:514 subtracts one without log spacing)."""
    n = num_steps if log_time_spacing else num_steps - 1
    x = sigmas[0] * noise
    if not multisteps:
        orders = _singlestep_orders(order, n)
        g = _Grid(sigmas, len(orders), log_time_spacing)
        for i, o in enumerate(orders):
            x = _single(denoise_fn, g, x, g.t[i], g.t[i + 1], o, x0_pred)
        return x.clamp(-1.0, 1.0)
    assert n >= order
    g = _Grid(sigmas, n, log_time_spacing)
    ms, ts = [_model(denoise_fn, g, x, g.t[0], x0_pred)], [g.t[0]]
    for step in range(1, order):                                                    # warm-up with increasing order
        x = _multi(g, x, ms, ts, g.t[step], step, x0_pred)
        ts.append(g.t[step])
        ms.append(_model(denoise_fn, g, x, g.t[step], x0_pred))
    for step in range(order, n + 1):
        x = _multi(g, x, ms, ts, g.t[step], min(order, n + 1 - step), x0_pred)       # lower order for the last steps
        ts = ts[1:] + [g.t[step]]
        if step < n:
            ms = ms[1:] + [_model(denoise_fn, g, x, g.t[step], x0_pred)]
        else:
            ms = ms[1:] + [ms[-1]]
    return x.clamp(-1.0, 1.0)


def _unipc_update(denoise_fn, g, x, ms, ts, tc, order, x0, use_corrector):
    """multistep_uni_pc_update with variant 'bh2' and x_t=None (:870-987)."""
    t0, m0 = ts[-1], ms[-1]
    h = g.lam(tc) - g.lam(t0)
    rks, d1s = [], []
    for i in range(1, order):
        rk = (g.lam(ts[-(i + 1)]) - g.lam(t0)) / h
        rks.append(rk)
        d1s.append((ms[-(i + 1)] - m0) / rk)
    rks = torch.stack([torch.as_tensor(r, dtype=torch.float32) for r in rks + [1.0]])
    hh = -h if x0 else h
    h_phi_1 = torch.expm1(hh)
    h_phi_k = h_phi_1 / hh - 1
    b_h = torch.expm1(hh)
    rows, b, fact = [], [], 1
    for i in range(1, order + 1):
        rows.append(torch.pow(rks, i - 1))
        b.append(h_phi_k * fact / b_h)
        fact *= i + 1
        h_phi_k = h_phi_k / hh - 1 / fact
    R, b = torch.stack(rows), torch.stack(b)
    pred = 0
    if d1s:
        rhos_p = torch.tensor([0.5]) if order == 2 else torch.linalg.solve(R[:-1, :-1], b[:-1])
        pred = sum(r * d for r, d in zip(rhos_p, d1s))
    scale = b_h if x0 else g.sig(tc) * b_h
    base = g.sig(tc) / g.sig(t0) * x - h_phi_1 * m0 if x0 else x - g.sig(tc) * h_phi_1 * m0
    x_t, model_t = base - scale * pred, None
    if use_corrector:
        rhos_c = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(R, b)
        model_t = _model(denoise_fn, g, x_t, tc, x0)
        corr = sum(r * d for r, d in zip(rhos_c[:-1], d1s)) if d1s else 0
        x_t = base - scale * (corr + rhos_c[-1] * (model_t - m0))
    return x_t, model_t


def unipc_sampler(noise, denoise_fn, sigmas, num_steps=20, order=2, x0_pred=True, log_time_spacing=True):
    """UniPCSampler.forward (:989-1053): num_steps network evaluations."""
    n = num_steps if log_time_spacing else num_steps - 1
    assert n >= order
    x = sigmas[0] * noise
    g = _Grid(sigmas, n, log_time_spacing)
    ms, ts = [_model(denoise_fn, g, x, g.t[0], x0_pred)], [g.t[0]]
    for step in range(1, order):
        x, m = _unipc_update(denoise_fn, g, x, ms, ts, g.t[step], step, x0_pred, True)
        ts.append(g.t[step])
        ms.append(m)
    for step in range(order, n + 1):
        x, m = _unipc_update(denoise_fn, g, x, ms, ts, g.t[step], min(order, n + 1 - step), x0_pred, step != n)
        ts = ts[1:] + [g.t[step]]
        ms = ms[1:] + [m if step < n else ms[-1]]
    return x.clamp(-1.0, 1.0)
