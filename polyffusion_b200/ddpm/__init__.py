"""``DenoiseDiffusion`` -- drop-in for ``ddpm/__init__.py`` of the reference (legacy unconditional DDPM).

Same constructor and attributes (ddpm/__init__.py:16-34: fp32 ``linspace(1e-4, 0.02, T)`` betas as
a buffer; ``alpha`` / ``alpha_bar`` / ``sigma2`` as plain attributes).  ``p_sample`` evaluates
``eps_model(xt, t)`` (any CUDA ``nn.Module``) and applies the reverse step
``(xt - (1-alpha)/sqrt(1-alpha_bar) eps)/sqrt(alpha) + sqrt(beta) noise`` (:66-88, noise is added even
at t = 0) in one ``pf_sample_step_ddpm_legacy`` kernel.  ``beta`` is a registered buffer and follows
``.to(device)``; ``alpha`` / ``alpha_bar`` / ``sigma2`` are plain attributes exactly as in the reference
 -- the step kernels read host copies of the per-step scalars made once here.
``loss`` (training, train/train_ddpm.py) is the reference's expression over the differentiable path of
the eps-model (``ddpm.unet.UNet`` switches to its PyTorch graph when gradients are enabled).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from polyffusion_b200._step import fused_q_sample, fused_step


class DenoiseDiffusion(nn.Module):
    def __init__(self, eps_model: nn.Module, n_steps: int):
        super().__init__()
        self.eps_model = eps_model
        self.register_buffer("beta", torch.linspace(0.0001, 0.02, n_steps))
        self.alpha = 1.0 - self.beta
        self.alpha_bar = torch.cumprod(self.alpha, dim=0)
        self.n_steps = n_steps
        self.sigma2 = self.beta
        with torch.no_grad():
            self._h = {
                "c0": ((1 - self.alpha) / (1 - self.alpha_bar) ** 0.5).tolist(),
                "c1": (1 / (self.alpha**0.5)).tolist(),
                "c2": (self.sigma2**0.5).tolist(),
                "qa": (self.alpha_bar**0.5).tolist(),
                "qb": ((1 - self.alpha_bar) ** 0.5).tolist(),
            }

    @staticmethod
    def _uniform_t(t: torch.Tensor) -> Optional[int]:
        """The common timestep of the batch, or None if the entries differ.  This reads ``t`` back to the
        host (one device->host copy per call): the legacy API passes the step only as a tensor
        (ddpm/__init__.py:66), and the per-step scalars of the fused kernel live on the host."""
        tl = t.tolist()
        return int(tl[0]) if all(v == tl[0] for v in tl) else None

    def q_xt_x0(self, x0: torch.Tensor, t: torch.Tensor):
        """Mean and variance of q(x_t | x_0) (ddpm/__init__.py:36-48)."""
        from polyffusion_b200.ddpm.utils import gather

        ab = gather(self.alpha_bar.to(t.device), t)
        return ab**0.5 * x0, 1 - ab

    @torch.no_grad()
    def q_sample(self, x0: torch.Tensor, t: torch.Tensor, eps: Optional[torch.Tensor] = None):
        if eps is None:
            eps = torch.randn_like(x0)
        ti = self._uniform_t(t)
        if ti is not None:
            return fused_q_sample(x0, eps, self._h["qa"][ti], self._h["qb"][ti])
        out = [fused_q_sample(x0[i : i + 1], eps[i : i + 1], self._h["qa"][int(v)], self._h["qb"][int(v)])
               for i, v in enumerate(t.tolist())]
        return torch.cat(out)

    @torch.no_grad()
    def p_sample(self, xt: torch.Tensor, t: torch.Tensor):
        eps_theta = self.eps_model(xt, t)
        noise = torch.randn(xt.shape, device=xt.device)
        ti = self._uniform_t(t)
        h = self._h
        if ti is not None:
            x_prev, _, _ = fused_step("legacy", xt, eps_theta, None, noise,
                                      (h["c0"][ti], h["c1"][ti], h["c2"][ti], 0.0, 0.0),
                                      want_x0=False, want_eps=False)
            return x_prev
        outs = []
        for i, v in enumerate(t.tolist()):
            v = int(v)
            xp, _, _ = fused_step("legacy", xt[i : i + 1], eps_theta[i : i + 1], None, noise[i : i + 1],
                                  (h["c0"][v], h["c1"][v], h["c2"][v], 0.0, 0.0), want_x0=False,
                                  want_eps=False)
            outs.append(xp)
        return torch.cat(outs)

    def loss(self, x0: torch.Tensor, noise: Optional[torch.Tensor] = None):
        """Simplified DDPM loss (ddpm/__init__.py:90-110): uniform t per sample, MSE(noise, eps_theta)."""
        batch_size = x0.shape[0]
        t = torch.randint(0, self.n_steps, (batch_size,), device=x0.device, dtype=torch.long)
        if noise is None:
            noise = torch.randn_like(x0)
        mean, var = self.q_xt_x0(x0, t)
        xt = mean + (var**0.5) * noise
        eps_theta = self.eps_model(xt, t)
        return torch.nn.functional.mse_loss(noise, eps_theta)
