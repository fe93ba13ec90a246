"""``LatentDiffusion`` -- drop-in for ``stable_diffusion/latent_diffusion.py`` of the reference.

Holds the eps-model (a ``polyffusion_b200`` ``UNetModel``), the beta / alpha / alpha_bar tables built
exactly as latent_diffusion.py:90-103 builds them (fp64 linspace of sqrt(beta), squared, cumprod in
fp64, cast to fp32 non-trainable Parameters; ``sigma2`` aliases ``beta``), and forwards
``__call__(x, t, context)`` to the CUDA UNet (latent_diffusion.py:138-147).  ``q_sample`` runs on the
``pf_q_sample`` kernel.  The autoencoder hooks are kept (identity when ``autoencoder`` is None, which
is the case in every Polyffusion config, inference_sdf.py:537).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from polyffusion_b200._lib import check, current_stream, lib, ptr
from polyffusion_b200.stable_diffusion.model.unet import UNetModel


class LatentDiffusion(nn.Module):
    eps_model: UNetModel

    def __init__(
        self,
        unet_model: UNetModel,
        autoencoder,
        latent_scaling_factor: float,
        n_steps: int,
        linear_start: float,
        linear_end: float,
    ):
        super().__init__()
        self.eps_model = unet_model
        self.first_stage_model = autoencoder
        if self.first_stage_model is not None:
            for param in self.first_stage_model.parameters():
                param.requires_grad = False
        self.latent_scaling_factor = latent_scaling_factor
        self.n_steps = n_steps
        beta = torch.linspace(linear_start**0.5, linear_end**0.5, n_steps, dtype=torch.float64) ** 2
        alpha = 1.0 - beta
        alpha_bar = torch.cumprod(alpha, dim=0)
        self.alpha = nn.Parameter(alpha.to(torch.float32), requires_grad=False)
        self.beta = nn.Parameter(beta.to(torch.float32), requires_grad=False)
        self.alpha_bar = nn.Parameter(alpha_bar.to(torch.float32), requires_grad=False)
        self.sigma2 = self.beta

    @property
    def device(self):
        return next(iter(self.eps_model.parameters())).device

    def autoencoder_encode(self, image: torch.Tensor):
        if self.first_stage_model is not None:
            return self.latent_scaling_factor * self.first_stage_model.encode(image).sample()
        return image

    def autoencoder_decode(self, z: torch.Tensor):
        if self.first_stage_model is not None:
            return self.first_stage_model.decode(z / self.latent_scaling_factor)
        return z

    def forward(self, x: torch.Tensor, t: torch.Tensor, context: torch.Tensor):
        return self.eps_model(x, t, context)

    @torch.no_grad()
    def q_sample(self, x0: torch.Tensor, t: torch.Tensor, eps: Optional[torch.Tensor] = None):
        """sqrt(alpha_bar_t) x0 + sqrt(1 - alpha_bar_t) eps with a per-sample t (latent_diffusion.py:149-177)."""
        if eps is None:
            eps = torch.randn_like(x0)
        ab = self.alpha_bar.gather(-1, t).reshape(-1, 1, 1, 1)
        if (t == t[0]).all():
            # one timestep for the whole batch: fused kernel
            a = float(ab[0] ** 0.5)
            b = float((1 - ab[0]) ** 0.5)
            x0c, epsc = x0.contiguous().float(), eps.contiguous().float()
            out = torch.empty_like(x0c)
            check(lib().pf_q_sample(ptr(x0c), ptr(epsc), ptr(out), out.numel(), a, b, current_stream()))
            return out
        return ab**0.5 * x0 + ((1 - ab) ** 0.5) * eps

    def q_xt_x0(self, x0: torch.Tensor, t: torch.Tensor):
        """Mean and variance of q(x_t | x_0) (latent_diffusion.py:149-162)."""
        ab = self.alpha_bar.gather(-1, t).reshape(-1, 1, 1, 1)
        return ab**0.5 * x0, 1 - ab

    def loss(self, x0: torch.Tensor, cond: torch.Tensor, noise: Optional[torch.Tensor] = None,
             cond_concat: Optional[torch.Tensor] = None):
        """Simplified DDPM loss, statement for statement latent_diffusion.py:203-240: one uniform
        t per sample, x_t = q_sample(x0, t, noise), MSE between the noise and eps_theta(x_t, t, cond).
        In training mode the eps-model runs its differentiable PyTorch graph (unet_torch.py); under
        ``no_grad`` / ``eval`` (validation) it runs the CUDA plan."""
        batch_size = x0.shape[0]
        t = torch.randint(0, self.n_steps, (batch_size,), device=x0.device, dtype=torch.long)
        if self.first_stage_model is not None:
            x0 = self.autoencoder_encode(x0)
        if noise is None:
            noise = torch.randn_like(x0)
        mean, var = self.q_xt_x0(x0, t)
        xt = mean + (var**0.5) * noise
        if cond_concat is not None:
            xt = torch.concat([xt, cond_concat], dim=1)
        eps_theta = self.eps_model(xt, t, cond)
        return torch.nn.functional.mse_loss(noise, eps_theta)
