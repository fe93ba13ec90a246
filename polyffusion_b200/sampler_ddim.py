"""``DDIMSampler`` -- drop-in for ``sampler_ddim.py`` of the reference.

Constructor, tau discretisation (integer arithmetic, bit-exact: sampler_ddim.py:63-73), tables
(:75-102) and method signatures match the reference.  The per-step update
(get_x_prev_and_pred_x0, :233-272) plus classifier-free guidance and the inpainting blend (:355-359)
run as one ``pf_sample_step_ddim`` kernel; the ``sigma == 0`` test is done on a host copy of the
sigma table, so a step has no device->host synchronisation.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

import os

from polyffusion_b200._loop import FusedLoop
from polyffusion_b200._step import fused_q_sample, fused_step
from polyffusion_b200.stable_diffusion.latent_diffusion import LatentDiffusion
from polyffusion_b200.stable_diffusion.sampler import DiffusionSampler


class DDIMSampler(DiffusionSampler):
    model: LatentDiffusion

    def __init__(self, model: LatentDiffusion, n_steps: int, ddim_discretize: str = "uniform",
                 ddim_eta: float = 0.0, is_show_image=False):
        super().__init__(model)
        self.is_show_image = is_show_image
        self.n_steps = model.n_steps
        if ddim_discretize == "uniform":
            c = self.n_steps // n_steps
            self.time_steps = np.asarray(list(range(0, self.n_steps, c))) + 1
        elif ddim_discretize == "quad":
            self.time_steps = ((np.linspace(0, np.sqrt(self.n_steps * 0.8), n_steps)) ** 2).astype(int) + 1
        else:
            raise NotImplementedError(ddim_discretize)
        with torch.no_grad():
            alpha_bar = self.model.alpha_bar
            self.ddim_alpha = alpha_bar[self.time_steps].clone().to(torch.float32)
            self.ddim_alpha_sqrt = torch.sqrt(self.ddim_alpha)
            self.ddim_alpha_prev = torch.cat([alpha_bar[0:1], alpha_bar[self.time_steps[:-1]]])
            self.ddim_sigma = (
                ddim_eta
                * ((1 - self.ddim_alpha_prev) / (1 - self.ddim_alpha)
                   * (1 - self.ddim_alpha / self.ddim_alpha_prev)) ** 0.5
            )
            self.ddim_sqrt_one_minus_alpha = (1.0 - self.ddim_alpha) ** 0.5
            self._h = {
                "c0": self.ddim_sqrt_one_minus_alpha.tolist(),
                "c1": (self.ddim_alpha**0.5).tolist(),
                "c2": (self.ddim_alpha_prev**0.5).tolist(),
                "c3": (1.0 - self.ddim_alpha_prev - self.ddim_sigma**2).sqrt().tolist(),
                "c4": self.ddim_sigma.tolist(),
                "qa": self.ddim_alpha_sqrt.tolist(),
                "qb": self.ddim_sqrt_one_minus_alpha.tolist(),
            }
        # whole-step graph loop / noise source: as in SDFSampler (sampler_sdf.py, _loop.py)
        self.fused_loop = os.environ.get("PF_FUSED_LOOP", "1") != "0"
        self.noise = os.environ.get("PF_NOISE", "torch")
        self.seed = 0
        self.sample0 = 0
        h = self._h
        self._loop = FusedLoop(self.model.eps_model, 2,
                               [(h["c0"][i], h["c1"][i], h["c2"][i], h["c3"][i], h["c4"][i], h["qa"][i], h["qb"][i])
                                for i in range(len(self.time_steps))], [int(v) for v in self.time_steps])

    def _can_fuse(self, x, uncond_scale, uncond_cond, cond_concat):
        from polyffusion_b200.stable_diffusion.model.unet import UNetModel

        guided = uncond_cond is not None and uncond_scale not in (0.0, 1.0)
        return (self.fused_loop and x.is_cuda and not guided and cond_concat is None
                and isinstance(self.model.eps_model, UNetModel) and self.model.first_stage_model is None)

    def _run_fused(self, x, cond, start, n, *, orig, mask, orig_noise, temperature, repeat_noise, uncond_scale,
                   uncond_cond):
        if uncond_cond is not None and uncond_scale == 0.0:
            cond = uncond_cond
        shape = tuple(x.shape)
        sigma = self._h["c4"]

        def draw(index):
            # reference order inside a step: step noise if sigma != 0 (sampler_ddim.py:255-262), then the
            # known-region noise of q_sample when no orig_noise was given (:293-294)
            nz = None
            if sigma[index] != 0.0:
                nz = torch.randn((1, *shape[1:]) if repeat_noise else shape, device=x.device)
            nk = torch.randn_like(orig) if (orig is not None and orig_noise is None) else None
            return nk, nz

        return self._loop.run(x, cond, start, n, orig=orig, mask=mask, noise_mode=self.noise,
                              fixed_noise_kn=orig_noise if orig is not None else None, temperature=temperature,
                              seed=self.seed, sample0=self.sample0, draw=draw)

    def _coefs(self, index: int):
        h = self._h
        return h["c0"][index], h["c1"][index], h["c2"][index], h["c3"][index], h["c4"][index]

    def _draw_noise(self, x, index, repeat_noise):
        if self._h["c4"][index] == 0.0:
            return None, 0
        if repeat_noise:
            n = torch.randn((1, *x.shape[1:]), device=x.device)
            return n, n.numel()
        return torch.randn(x.shape, device=x.device), 0

    def _step(self, x, c, t, index, *, repeat_noise=False, temperature=1.0, uncond_scale=1.0,
              uncond_cond=None, cond_concat=None, orig=None, mask=None, orig_noise=None, want_aux=True):
        index = int(index)
        x_in = x if cond_concat is None else torch.concat([x, cond_concat], dim=1)
        e_c, e_u = self._eps_pair(x_in, t, c, uncond_scale=uncond_scale, uncond_cond=uncond_cond)
        noise, bcast = self._draw_noise(x, index, repeat_noise)
        noise_kn = None
        if orig is not None:
            # q_sample(orig, index, noise=orig_noise) draws fresh noise when none is given (:293-294)
            noise_kn = orig_noise if orig_noise is not None else torch.randn_like(orig)
        kn = (self._h["qa"][index], self._h["qb"][index])
        return fused_step("ddim", x, e_c, e_u, noise, self._coefs(index), uncond_scale=uncond_scale,
                          temperature=temperature, orig=orig, mask=mask, noise_kn=noise_kn, kn=kn,
                          want_x0=want_aux, want_eps=want_aux, noise_bcast=bcast)

    @torch.no_grad()
    def sample(self, shape: List[int], cond: torch.Tensor, repeat_noise: bool = False,
               temperature: float = 1.0, x_last: Optional[torch.Tensor] = None,
               uncond_scale: float = 1.0, uncond_cond: Optional[torch.Tensor] = None, t_start: int = 0):
        """sampler_ddim.py:104-166."""
        device = self.model.device
        bs = shape[0]
        x = x_last if x_last is not None else torch.randn(shape, device=device)
        time_steps = np.flip(self.time_steps)[t_start:]
        if len(time_steps) and self._can_fuse(x, uncond_scale, uncond_cond, None):
            return self._run_fused(x, cond, len(time_steps) - 1, len(time_steps), orig=None, mask=None,
                                   orig_noise=None, temperature=temperature, repeat_noise=repeat_noise,
                                   uncond_scale=uncond_scale, uncond_cond=uncond_cond)
        for i, step in enumerate(time_steps):
            index = len(time_steps) - i - 1
            ts = x.new_full((bs,), int(step), dtype=torch.long)
            x, _, _ = self._step(x, cond, ts, index, repeat_noise=repeat_noise, temperature=temperature,
                                 uncond_scale=uncond_scale, uncond_cond=uncond_cond, want_aux=False)
        return x

    @torch.no_grad()
    def advance(self, x: torch.Tensor, cond: torch.Tensor, start_index: int, n_steps: int, *,
                orig: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                orig_noise: Optional[torch.Tensor] = None, uncond_scale: float = 1.0,
                uncond_cond: Optional[torch.Tensor] = None, cond_concat=None):
        """``n_steps`` consecutive iterations of the sample / paint loop body (sampler_ddim.py:145-163,
        343-359) starting at schedule index ``start_index``."""
        n_steps = min(int(n_steps), int(start_index) + 1)
        if self._can_fuse(x, uncond_scale, uncond_cond, cond_concat):
            return self._run_fused(x, cond, int(start_index), n_steps, orig=orig, mask=mask, orig_noise=orig_noise,
                                   temperature=1.0, repeat_noise=False, uncond_scale=uncond_scale,
                                   uncond_cond=uncond_cond)
        bs = x.shape[0]
        for index in range(int(start_index), int(start_index) - n_steps, -1):
            ts = x.new_full((bs,), int(self.time_steps[index]), dtype=torch.long)
            x, _, _ = self._step(x, cond, ts, index, uncond_scale=uncond_scale, uncond_cond=uncond_cond,
                                 cond_concat=cond_concat, orig=orig, mask=mask, orig_noise=orig_noise,
                                 want_aux=False)
        return x

    @torch.no_grad()
    def p_sample(self, x: torch.Tensor, c: torch.Tensor, t: torch.Tensor, step: int, index: int, *,
                 repeat_noise: bool = False, temperature: float = 1.0, uncond_scale: float = 1.0,
                 uncond_cond: Optional[torch.Tensor] = None, cond_concat=None):
        """sampler_ddim.py:168-218: returns (x_prev, pred_x0, e_t)."""
        return self._step(x, c, t, index, repeat_noise=repeat_noise, temperature=temperature,
                          uncond_scale=uncond_scale, uncond_cond=uncond_cond, cond_concat=cond_concat)

    def get_x_prev_and_pred_x0(self, e_t: torch.Tensor, index: int, x: torch.Tensor, *,
                               temperature: float, repeat_noise: bool):
        """sampler_ddim.py:220-272, given an already computed eps."""
        index = int(index)
        noise, bcast = self._draw_noise(x, index, repeat_noise)
        x_prev, x0, _ = fused_step("ddim", x, e_t, None, noise, self._coefs(index),
                                   temperature=temperature, want_eps=False, noise_bcast=bcast)
        return x_prev, x0

    @torch.no_grad()
    def q_sample(self, x0: torch.Tensor, index: int, noise: Optional[torch.Tensor] = None):
        """sampler_ddim.py:274-299 (indexes the S-length tables)."""
        if noise is None:
            noise = torch.randn_like(x0)
        index = int(index)
        return fused_q_sample(x0, noise, self._h["qa"][index], self._h["qb"][index])

    @torch.no_grad()
    def paint(self, x: torch.Tensor, cond: torch.Tensor, t_start: int, *,
              orig: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
              orig_noise: Optional[torch.Tensor] = None, uncond_scale: float = 1.0,
              uncond_cond: Optional[torch.Tensor] = None, cond_concat=None, repaint_n=1):
        """sampler_ddim.py:301-362: after every step the known region is replaced by
        q_sample(orig, index, orig_noise) (the *current* index, fixed noise)."""
        bs = x.shape[0]
        time_steps = np.flip(self.time_steps[: t_start + 1])
        if repaint_n == 1 and self._can_fuse(x, uncond_scale, uncond_cond, cond_concat):
            return self._run_fused(x, cond, len(time_steps) - 1, len(time_steps), orig=orig, mask=mask,
                                   orig_noise=orig_noise, temperature=1.0, repeat_noise=False,
                                   uncond_scale=uncond_scale, uncond_cond=uncond_cond)
        for i, step in enumerate(time_steps):
            index = len(time_steps) - i - 1
            ts = x.new_full((bs,), int(step), dtype=torch.long)
            x, _, _ = self._step(x, cond, ts, index, uncond_scale=uncond_scale, uncond_cond=uncond_cond,
                                 cond_concat=cond_concat, orig=orig, mask=mask, orig_noise=orig_noise,
                                 want_aux=False)
        return x
