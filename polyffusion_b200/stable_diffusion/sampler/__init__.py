"""``DiffusionSampler`` -- drop-in base class for ``stable_diffusion/sampler/__init__.py``.

``get_eps`` keeps the reference's classifier-free-guidance dispatch
(sampler/__init__.py:63-80): ``s == 1`` or no unconditional embedding -> one UNet evaluation on
``c``; ``s == 0`` -> one evaluation on ``uncond_cond``; otherwise one evaluation on the doubled batch
``cat([x, x]), cat([t, t]), cat([uncond_cond, c])``.  In the fused samplers the combination
``e_u + s (e_c - e_u)`` is not materialised here but inside the step kernel; ``get_eps`` itself
remains available (and materialises it) for callers that want eps only.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from polyffusion_b200.stable_diffusion.latent_diffusion import LatentDiffusion


class DiffusionSampler:
    model: LatentDiffusion

    def __init__(self, model: LatentDiffusion):
        super().__init__()
        self.model = model
        self.n_steps = model.n_steps

    # returns (e_cond, e_uncond_or_None) without combining
    def _eps_pair(self, x, t, c, *, uncond_scale: float, uncond_cond: Optional[torch.Tensor]):
        if uncond_cond is None or uncond_scale == 1.0:
            return self.model(x, t, c), None
        if uncond_scale == 0.0:
            return self.model(x, t, uncond_cond), None
        x_in = torch.cat([x] * 2)
        t_in = torch.cat([t] * 2)
        c_in = torch.cat([uncond_cond, c])
        e_u, e_c = self.model(x_in, t_in, c_in).chunk(2)
        return e_c, e_u

    def get_eps(self, x, t, c, *, uncond_scale: float, uncond_cond: Optional[torch.Tensor]):
        e_c, e_u = self._eps_pair(x, t, c, uncond_scale=uncond_scale, uncond_cond=uncond_cond)
        if e_u is None:
            return e_c
        return e_u + uncond_scale * (e_c - e_u)

    def sample(self, shape: List[int], cond: torch.Tensor, repeat_noise: bool = False,
               temperature: float = 1.0, x_last: Optional[torch.Tensor] = None,
               uncond_scale: float = 1.0, uncond_cond: Optional[torch.Tensor] = None,
               skip_steps: int = 0):
        raise NotImplementedError()

    def paint(self, x: torch.Tensor, cond: torch.Tensor, t_start: int, *,
              orig: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
              orig_noise: Optional[torch.Tensor] = None, uncond_scale: float = 1.0,
              uncond_cond: Optional[torch.Tensor] = None):
        raise NotImplementedError()

    def q_sample(self, x0: torch.Tensor, index: int, noise: Optional[torch.Tensor] = None):
        raise NotImplementedError()
