"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's diffusion schedules and samplers.

Functional fp32 torch-CPU code: schedule tables, classifier-free guidance, the DDPM step with
RePaint, the DDIM step, and the legacy DDPM step.  Noise is *injected* (``noise_fn(shape)``) so a
test can feed the same tape to this oracle and to the CUDA samplers.

Reference followed (relative to /root/reference/polyffusion/):
  stable_diffusion/latent_diffusion.py:90-103 (beta schedule), stable_diffusion/sampler/__init__.py:63-80
  (CFG), sampler_sdf.py:52-78 (tables), 121-171 (p_sample), 192 (q_sample), 289-341 (paint/RePaint),
  sampler_ddim.py:63-102 (tau + tables), 141-145 (loop indices), 233-272 (step), 296-299 (q_sample),
  336-359 (paint), ddpm/__init__.py:25-34 and 66-88 (legacy schedule + p_sample).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch


# ------------------------------------------------------------------ schedules
def ldm_schedule(n_steps: int = 1000, linear_start: float = 0.00085, linear_end: float = 0.012):
    """latent_diffusion.py:90-103 -> fp32 (alpha, beta, alpha_bar)."""
    beta = torch.linspace(linear_start**0.5, linear_end**0.5, n_steps, dtype=torch.float64) ** 2
    alpha = 1.0 - beta
    alpha_bar = torch.cumprod(alpha, dim=0)
    return alpha.float(), beta.float(), alpha_bar.float()


def ddpm_tables(alpha_bar: torch.Tensor, beta: torch.Tensor) -> dict:
    """sampler_sdf.py:52-78."""
    ab_prev = torch.cat([alpha_bar.new_tensor([1.0]), alpha_bar[:-1]])
    var = beta * (1.0 - ab_prev) / (1.0 - alpha_bar)
    return dict(
        sqrt_ab=alpha_bar**0.5,
        sqrt_1m_ab=(1.0 - alpha_bar) ** 0.5,
        sqrt_recip_ab=alpha_bar**-0.5,
        sqrt_recip_m1_ab=(1 / alpha_bar - 1) ** 0.5,
        log_var=torch.log(torch.clamp(var, min=1e-20)),
        mean_x0=beta * (ab_prev**0.5) / (1.0 - alpha_bar),
        mean_xt=(1.0 - ab_prev) * ((1 - beta) ** 0.5) / (1.0 - alpha_bar),
    )


def ddim_time_steps(n_total: int, n_steps: int, discretize: str = "uniform") -> np.ndarray:
    """sampler_ddim.py:63-73 (integer arithmetic; bit-exact path)."""
    if discretize == "uniform":
        c = n_total // n_steps
        return np.asarray(list(range(0, n_total, c))) + 1
    if discretize == "quad":
        return ((np.linspace(0, np.sqrt(n_total * 0.8), n_steps)) ** 2).astype(int) + 1
    raise NotImplementedError(discretize)


def ddim_tables(alpha_bar: torch.Tensor, tau: np.ndarray, eta: float) -> dict:
    """sampler_ddim.py:75-102."""
    a = alpha_bar[tau].clone().float()
    a_prev = torch.cat([alpha_bar[0:1], alpha_bar[tau[:-1]]])
    sigma = eta * ((1 - a_prev) / (1 - a) * (1 - a / a_prev)) ** 0.5
    return dict(alpha=a, alpha_sqrt=torch.sqrt(a), alpha_prev=a_prev, sigma=sigma,
                sqrt_1m_alpha=(1.0 - a) ** 0.5)


# ------------------------------------------------------------------ eps with guidance
def guided_eps(eps_fn: Callable, x, t, c, uncond_scale: float, uncond_cond):
    """sampler/__init__.py:63-80."""
    if uncond_cond is None or uncond_scale == 1.0:
        return eps_fn(x, t, c)
    if uncond_scale == 0.0:
        return eps_fn(x, t, uncond_cond)
    e_u, e_c = eps_fn(torch.cat([x] * 2), torch.cat([t] * 2), torch.cat([uncond_cond, c])).chunk(2)
    return e_u + uncond_scale * (e_c - e_u)


# ------------------------------------------------------------------ DDPM (SDFSampler)
def ddpm_p_sample(tb: dict, eps_fn, x, c, step: int, noise_fn, *, temperature=1.0, uncond_scale=1.0,
                  uncond_cond=None, repeat_noise=False):
    """sampler_sdf.py:117-171 -> (x_prev, x0, e_t)."""
    bs = x.shape[0]
    t = x.new_full((bs,), step, dtype=torch.long)
    e_t = guided_eps(eps_fn, x, t, c, uncond_scale, uncond_cond)
    full = lambda v: x.new_full((bs, 1, 1, 1), float(v))
    x0 = full(tb["sqrt_recip_ab"][step]) * x - full(tb["sqrt_recip_m1_ab"][step]) * e_t
    mean = full(tb["mean_x0"][step]) * x0 + full(tb["mean_xt"][step]) * x
    if step == 0:
        noise = 0
    elif repeat_noise:
        noise = noise_fn((1, *x.shape[1:]))
    else:
        noise = noise_fn(tuple(x.shape))
    noise = noise * temperature
    x_prev = mean + (0.5 * full(tb["log_var"][step])).exp() * noise
    return x_prev, x0, e_t


def ddpm_sample(alpha_bar, beta, eps_fn, x, cond, noise_fn, *, t_start=0, **kw):
    """sampler_sdf.py:229-255 starting from x (= x_last)."""
    tb = ddpm_tables(alpha_bar, beta)
    for step in np.flip(np.arange(len(alpha_bar)))[t_start:]:
        x, _, _ = ddpm_p_sample(tb, eps_fn, x, cond, int(step), noise_fn, **kw)
    return x


def ddpm_paint(alpha_bar, beta, eps_fn, x, cond, t_start: int, noise_fn, *, orig=None, mask=None,
               uncond_scale=1.0, uncond_cond=None, repaint_n=1):
    """sampler_sdf.py:289-341 (RePaint; known region re-noised every step; beta, not sqrt(beta), in
    the resampling jump)."""
    tb = ddpm_tables(alpha_bar, beta)
    for step in np.flip(np.arange(len(alpha_bar))[: t_start + 1]):
        step = int(step)
        if orig is None:
            x, _, _ = ddpm_p_sample(tb, eps_fn, x, cond, step, noise_fn, uncond_scale=uncond_scale,
                                    uncond_cond=uncond_cond)
            continue
        x_t = x
        for u in range(repaint_n):
            noise = noise_fn(tuple(orig.shape)) if step > 0 else torch.zeros_like(orig)
            x_kn = tb["sqrt_ab"][step] * orig + tb["sqrt_1m_ab"][step] * noise
            x_unkn, _, _ = ddpm_p_sample(tb, eps_fn, x_t, cond, step, noise_fn,
                                         uncond_scale=uncond_scale, uncond_cond=uncond_cond)
            x = x_kn * mask + x_unkn * (1 - mask)
            if u < repaint_n - 1 and step > 0:
                noise = noise_fn(tuple(orig.shape))
                x_t = (1 - beta[step - 1]) ** 0.5 * x + beta[step - 1] * noise
    return x


# ------------------------------------------------------------------ DDIM
def ddim_step(tb: dict, e_t, index: int, x, noise_fn, *, temperature=1.0, repeat_noise=False):
    """sampler_ddim.py:233-272 -> (x_prev, pred_x0)."""
    alpha, alpha_prev, sigma = tb["alpha"][index], tb["alpha_prev"][index], tb["sigma"][index]
    pred_x0 = (x - tb["sqrt_1m_alpha"][index] * e_t) / (alpha**0.5)
    dir_xt = (1.0 - alpha_prev - sigma**2).sqrt() * e_t
    if sigma == 0.0:
        noise = 0.0
    elif repeat_noise:
        noise = noise_fn((1, *x.shape[1:]))
    else:
        noise = noise_fn(tuple(x.shape))
    noise = noise * temperature
    return (alpha_prev**0.5) * pred_x0 + dir_xt + sigma * noise, pred_x0


def ddim_run(alpha_bar, eps_fn, x, cond, noise_fn, *, n_steps, discretize="uniform", eta=0.0,
             t_start=None, paint=False, orig=None, mask=None, orig_noise=None, uncond_scale=1.0,
             uncond_cond=None, temperature=1.0):
    """sampler_ddim.py:104-166 (sample: t_start skips from the top) and 301-362 (paint: t_start is
    the first index)."""
    tau = ddim_time_steps(len(alpha_bar), n_steps, discretize)
    tb = ddim_tables(alpha_bar, tau, eta)
    if paint:
        steps = np.flip(tau[: t_start + 1])
    else:
        steps = np.flip(tau)[(t_start or 0):]
    for i, step in enumerate(steps):
        index = len(steps) - i - 1
        t = x.new_full((x.shape[0],), int(step), dtype=torch.long)
        e_t = guided_eps(eps_fn, x, t, cond, uncond_scale, uncond_cond)
        x, _ = ddim_step(tb, e_t, index, x, noise_fn, temperature=temperature)
        if paint and orig is not None:
            n = orig_noise if orig_noise is not None else noise_fn(tuple(orig.shape))
            orig_t = tb["alpha_sqrt"][index] * orig + tb["sqrt_1m_alpha"][index] * n
            x = orig_t * mask + x * (1 - mask)
    return x


# ------------------------------------------------------------------ legacy DDPM
def legacy_schedule(n_steps: int):
    """ddpm/__init__.py:25-34 (fp32 throughout)."""
    beta = torch.linspace(0.0001, 0.02, n_steps)
    alpha = 1.0 - beta
    return alpha, beta, torch.cumprod(alpha, dim=0)


def legacy_p_sample(alpha, beta, alpha_bar, eps_fn, xt, t, noise_fn):
    """ddpm/__init__.py:66-88 (noise added at every t, including 0)."""
    g = lambda c: c.gather(-1, t).reshape(-1, 1, 1, 1)
    eps_theta = eps_fn(xt, t)
    eps_coef = (1 - g(alpha)) / (1 - g(alpha_bar)) ** 0.5
    mean = 1 / (g(alpha) ** 0.5) * (xt - eps_coef * eps_theta)
    return mean + (g(beta) ** 0.5) * noise_fn(tuple(xt.shape))
