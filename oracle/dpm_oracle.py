"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's DPM-Solver++ sampling loop as `scripts/inference.py`
drives it (`DPMS(model.forward_with_dpmsolver, condition, uncondition, cfg_scale, model_kwargs).sample(z, steps, order=2,
skip_type="time_uniform", method="multistep")`, scripts/inference.py:102-118).  Only tests, `__graft_entry__.smoke()` and
bench.py's CPU arms may import this module; the product (`pixart_sigma_b200/sampler.py`) must not.

Pinned against the unmodified reference (`diffusion/dpm_solver.py::DPMS`, imported through oracle/refshim.py) by
`oracle/gen_golden_dpm.py` -> `tests/golden/dpm_*.pt`, checked in tests/test_oracle.py.

What is restated (file:line in /root/reference):
  * linear beta schedule, float64                     diffusion/model/gaussian_diffusion.py:82,107-116
  * NoiseScheduleVP('discrete'): log-alpha table, t table, clip near T, piecewise-linear lookups
                                                      diffusion/model/dpm_solver.py:97-106,114-125,127-155,1285-1324
  * model_wrapper: t_input = (t - 1/N) * 1000, CFG batch [uncond, cond], eps = eu + s (ec - eu)
                                                      diffusion/model/dpm_solver.py:273-282,284-291,326-332
  * data prediction x0 = (x - sigma eps) / alpha      diffusion/model/dpm_solver.py:435-444
  * time_uniform steps linspace(T, 1/N, steps+1)      diffusion/model/dpm_solver.py:475-476,1180-1181
  * first-order (DDIM-like) and second-order multistep DPM-Solver++ updates
                                                      diffusion/model/dpm_solver.py:565-577,822-840
  * multistep driver with lower_order_final           diffusion/model/dpm_solver.py:1196-1241
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
import torch


def linear_betas(num_steps: int = 1000) -> np.ndarray:
    scale = 1000 / num_steps
    return np.linspace(scale * 0.0001, scale * 0.02, num_steps, dtype=np.float64)


class DiscreteSchedule:
    """log(alpha_t) tabulated at t_i = i / N, i = 1..N (float32 like the reference's default dtype), looked up by
    piecewise-linear interpolation; beyond the table the outermost segment is extended."""

    def __init__(self, betas: Optional[np.ndarray] = None, clipped_lambda: float = -5.1):
        betas = linear_betas() if betas is None else betas
        log_alphas = 0.5 * torch.log(1 - torch.tensor(betas)).cumsum(dim=0)          # float64
        log_sigmas = 0.5 * torch.log(1. - torch.exp(2. * log_alphas))
        lambs = log_alphas - log_sigmas
        idx = int(torch.searchsorted(torch.flip(lambs, [0]), torch.tensor(clipped_lambda, dtype=lambs.dtype)))
        if idx > 0:
            log_alphas = log_alphas[:-idx]
        self.log_alpha = log_alphas.to(torch.float32)
        self.total_N = self.log_alpha.numel()
        self.t = torch.linspace(0., 1., self.total_N + 1)[1:].to(torch.float32)
        self.T = 1.0

    def _interp(self, x: torch.Tensor, xp: torch.Tensor, yp: torch.Tensor) -> torch.Tensor:
        x = x.reshape(-1).to(torch.float32)
        k = xp.numel()
        seg = torch.searchsorted(xp.contiguous(), x.contiguous(), right=False) - 1     # xp[seg] < x <= xp[seg+1]
        seg = seg.clamp(0, k - 2)
        x0, x1, y0, y1 = xp[seg], xp[seg + 1], yp[seg], yp[seg + 1]
        return y0 + (x - x0) * (y1 - y0) / (x1 - x0)

    def log_alpha_t(self, t: torch.Tensor) -> torch.Tensor:
        return self._interp(t, self.t, self.log_alpha)

    def alpha(self, t):
        return torch.exp(self.log_alpha_t(t))

    def sigma(self, t):
        return torch.sqrt(1. - torch.exp(2. * self.log_alpha_t(t)))

    def lam(self, t):
        la = self.log_alpha_t(t)
        return la - 0.5 * torch.log(1. - torch.exp(2. * la))


def time_steps(schedule: DiscreteSchedule, steps: int, t_start: Optional[float] = None, t_end: Optional[float] = None):
    t_0 = 1. / schedule.total_N if t_end is None else t_end
    t_T = schedule.T if t_start is None else t_start
    return torch.linspace(t_T, t_0, steps + 1)


def model_input_time(schedule: DiscreteSchedule, t: torch.Tensor) -> torch.Tensor:
    return (t - 1. / schedule.total_N) * 1000.


def step_coefficients(schedule: DiscreteSchedule, ts: torch.Tensor, order: int = 2, lower_order_final: bool = True):
    """Per update i (x at ts[i] -> ts[i+1]): dict(sigma_s, alpha_s, a, b, c) with
    x0 = (x - sigma_s eps) / alpha_s;  x_next = a x - b x0 - c (x0 - x0_prev)   (c = 0 for first-order updates)."""
    steps = ts.numel() - 1
    out = []
    for i in range(steps):
        s, t = ts[i:i + 1], ts[i + 1:i + 2]
        step = i + 1
        step_order = 1 if step < order else (min(order, steps + 1 - step) if lower_order_final else order)
        lam_s, lam_t = schedule.lam(s), schedule.lam(t)
        h = lam_t - lam_s
        sigma_s, sigma_t = schedule.sigma(s), schedule.sigma(t)
        alpha_t = torch.exp(schedule.log_alpha_t(t))
        phi_1 = torch.expm1(-h)
        a = sigma_t / sigma_s
        b = alpha_t * phi_1
        c = torch.zeros_like(b)
        if step_order == 2:
            lam_p = schedule.lam(ts[i - 1:i])
            r0 = (lam_s - lam_p) / h
            c = 0.5 * b * (1. / r0)
        out.append(dict(sigma_s=float(sigma_s), alpha_s=float(schedule.alpha(s)), a=float(a), b=float(b), c=float(c),
                        order=step_order, t_input=float(model_input_time(schedule, s))))
    return out


def sample(model: Callable, z: torch.Tensor, condition: torch.Tensor, uncondition: torch.Tensor, cfg_scale: float,
           steps: int = 20, order: int = 2, model_kwargs: Optional[dict] = None, lower_order_final: bool = True,
           return_intermediate: bool = False):
    """`model(x (2n,4,h,w), t (2n,), cond (2n,...), **model_kwargs) -> eps (2n,4,h,w)`; returns x at t = 1/N."""
    model_kwargs = model_kwargs or {}
    sch = DiscreteSchedule()
    ts = time_steps(sch, steps)
    x = z.to(torch.float32)
    prev_t, prev_m = [], []
    inter = []

    def data_pred(x, t):
        t_c = t.expand(x.shape[0])
        t_in = model_input_time(sch, torch.cat([t_c] * 2))
        out = model(torch.cat([x] * 2), t_in, torch.cat([uncondition, condition]), **model_kwargs)
        eu, ec = out.to(torch.float32).chunk(2)
        eps = eu + cfg_scale * (ec - eu)
        return (x - sch.sigma(t) * eps) / sch.alpha(t)

    def first(x, s, t, m_s):
        h = sch.lam(t) - sch.lam(s)
        return sch.sigma(t) / sch.sigma(s) * x - torch.exp(sch.log_alpha_t(t)) * torch.expm1(-h) * m_s

    def second(x, t):
        m1, m0 = prev_m[-2], prev_m[-1]
        t1, t0 = prev_t[-2], prev_t[-1]
        l1, l0, lt = sch.lam(t1), sch.lam(t0), sch.lam(t)
        h0, h = l0 - l1, lt - l0
        r0 = h0 / h
        d1 = (1. / r0) * (m0 - m1)
        alpha_t = torch.exp(sch.log_alpha_t(t))
        phi_1 = torch.expm1(-h)
        return (sch.sigma(t) / sch.sigma(t0)) * x - (alpha_t * phi_1) * m0 - 0.5 * (alpha_t * phi_1) * d1

    t = ts[0:1]
    prev_t, prev_m = [t], [data_pred(x, t)]
    for step in range(1, order):
        t = ts[step:step + 1]
        x = first(x, prev_t[-1], t, prev_m[-1])
        inter.append(x)
        prev_t.append(t)
        prev_m.append(data_pred(x, t))
    for step in range(order, steps + 1):
        t = ts[step:step + 1]
        step_order = min(order, steps + 1 - step) if lower_order_final else order
        x = first(x, prev_t[-1], t, prev_m[-1]) if step_order == 1 else second(x, t)
        inter.append(x)
        prev_t = prev_t[1:] + [t]
        if step < steps:
            prev_m = prev_m[1:] + [data_pred(x, t)]
    return (x, inter) if return_intermediate else x


def toy_model(x: torch.Tensor, t: torch.Tensor, cond: torch.Tensor, **kwargs) -> torch.Tensor:
    """Deterministic stand-in denoiser for pinning the sampler arithmetic (smooth in x, t and the condition)."""
    c = cond.to(torch.float32).reshape(cond.shape[0], -1).mean(dim=1).reshape(-1, 1, 1, 1)
    tt = (t.to(torch.float32) / 1000.).reshape(-1, 1, 1, 1)
    return torch.tanh(0.7 * x.to(torch.float32) + 0.4 * tt + c) + 0.1 * torch.roll(x.to(torch.float32), 1, dims=-1) * tt
