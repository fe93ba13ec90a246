"""``SDFSampler`` -- drop-in for ``sampler_sdf.py`` of the reference (DDPM sampling + RePaint).

Same constructor, attributes (``time_steps``, ``model``, the coefficient tables) and method
signatures as sampler_sdf.py:37-350.  What differs is where the work happens:

* eps comes from the CUDA UNet plan (one ``pf_unet_forward`` per step, batch doubled for CFG);
* everything after eps -- classifier-free-guidance combine, x0, posterior mean, noise add and the
  RePaint blend with the re-noised known region -- is ONE kernel launch (``pf_sample_step_ddpm``)
  instead of the reference's 31-37 ATen ops;
* the per-step scalars are read from host-side float lists prepared once in ``__init__`` (the
  tables themselves are computed with the very same torch expressions as sampler_sdf.py:52-78), so
  a step performs no device->host synchronisation (the reference does five per step).

Random numbers are drawn with ``torch.randn`` in the reference's order (known-region noise first,
then step noise), so a seeded run consumes the generator exactly like the reference does.

``sample`` / ``paint`` loops without classifier-free guidance run as replays of ONE CUDA graph per step
(``_loop.FusedLoop``: UNet + in-kernel step + device-side step counter; the time embedding comes from a
table, the cross-attention vectors are computed once per loop).  ``sampler.noise = "philox"`` switches
those loops to in-kernel Philox noise keyed by the global sample index (``sampler.seed``,
``sampler.sample0``), which makes a batch sharded over ranks reproduce the single-GPU result.
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


class SDFSampler(DiffusionSampler):
    model: LatentDiffusion

    def __init__(self, model: LatentDiffusion, is_show_image=False):
        super().__init__(model)
        self.time_steps = np.asarray(list(range(self.n_steps)), dtype=np.int32)
        self.is_show_image = is_show_image
        with torch.no_grad():
            alpha_bar = self.model.alpha_bar
            beta = self.model.beta
            alpha_bar_prev = torch.cat([alpha_bar.new_tensor([1.0]), alpha_bar[:-1]])
            self.sqrt_alpha_bar = alpha_bar**0.5
            self.sqrt_1m_alpha_bar = (1.0 - alpha_bar) ** 0.5
            self.sqrt_recip_alpha_bar = alpha_bar**-0.5
            self.sqrt_recip_m1_alpha_bar = (1 / alpha_bar - 1) ** 0.5
            variance = beta * (1.0 - alpha_bar_prev) / (1.0 - alpha_bar)
            self.log_var = torch.log(torch.clamp(variance, min=1e-20))
            self.mean_x0_coef = beta * (alpha_bar_prev**0.5) / (1.0 - alpha_bar)
            self.mean_xt_coef = (1.0 - alpha_bar_prev) * ((1 - beta) ** 0.5) / (1.0 - alpha_bar)
            # host copies of every per-step scalar (one transfer here, none inside the loop)
            std = (0.5 * self.log_var).exp()
            self._h = {
                "c0": self.sqrt_recip_alpha_bar.tolist(),
                "c1": self.sqrt_recip_m1_alpha_bar.tolist(),
                "c2": self.mean_x0_coef.tolist(),
                "c3": self.mean_xt_coef.tolist(),
                "c4": std.tolist(),
                "qa": self.sqrt_alpha_bar.tolist(),
                "qb": self.sqrt_1m_alpha_bar.tolist(),
                "rp_a": ((1 - beta) ** 0.5).tolist(),
                "rp_b": beta.tolist(),
            }
        # whole-step graph loop (see module docstring); PF_FUSED_LOOP=0 keeps the per-step launches
        self.fused_loop = os.environ.get("PF_FUSED_LOOP", "1") != "0"
        self.noise = os.environ.get("PF_NOISE", "torch")  # "torch" (reference order) | "philox" (in-kernel)
        self.seed = 0
        self.sample0 = 0  # global index of this shard's first sample (Philox mode)
        h = self._h
        self._loop = FusedLoop(self.model.eps_model, 1,
                               [(h["c0"][i], h["c1"][i], h["c2"][i], h["c3"][i], h["c4"][i], h["qa"][i], h["qb"][i])
                                for i in range(self.n_steps)], list(range(self.n_steps)))

    def _can_fuse(self, x, uncond_scale, uncond_cond, cond_concat, repaint_n=1):
        from polyffusion_b200.stable_diffusion.model.unet import UNetModel

        guided = uncond_cond is not None and uncond_scale not in (0.0, 1.0)
        return (self.fused_loop and x.is_cuda and not guided and cond_concat is None and repaint_n == 1
                and isinstance(self.model.eps_model, UNetModel) and self.model.first_stage_model is None)

    def _run_fused(self, x, cond, start, n, *, orig, mask, temperature, repeat_noise, uncond_scale, uncond_cond):
        if uncond_cond is not None and uncond_scale == 0.0:
            cond = uncond_cond  # sampler/__init__.py:66-67
        shape = tuple(x.shape)

        def draw(step):
            # the reference's order: known-region noise (sampler_sdf.py:318), then the step noise (:157-160)
            nk = torch.randn_like(orig, device=orig.device) if (orig is not None and step > 0) else None
            if step == 0:
                return nk, None
            if repeat_noise:
                return nk, torch.randn((1, *shape[1:]), device=x.device)
            return nk, torch.randn(shape, device=x.device)

        return self._loop.run(x, cond, start, n, orig=orig, mask=mask, noise_mode=self.noise,
                              temperature=temperature, seed=self.seed, sample0=self.sample0, draw=draw)

    def _coefs(self, step: int):
        h = self._h
        return h["c0"][step], h["c1"][step], h["c2"][step], h["c3"][step], h["c4"][step]

    def _step(self, x, c, t, step, *, repeat_noise=False, temperature=1.0, uncond_scale=1.0,
              uncond_cond=None, cond_concat=None, orig=None, mask=None, noise_kn=None, want_aux=True):
        step = int(step)
        x_in = x if cond_concat is None else torch.concat([x, cond_concat], dim=1)
        e_c, e_u = self._eps_pair(x_in, t, c, uncond_scale=uncond_scale, uncond_cond=uncond_cond)
        noise, bcast = None, 0
        if step != 0:
            if repeat_noise:
                noise = torch.randn((1, *x.shape[1:]), device=x.device)
                bcast = noise.numel()
            else:
                noise = torch.randn(x.shape, device=x.device)
        kn = (self._h["qa"][step], self._h["qb"][step])
        return fused_step("ddpm", x, e_c, e_u, noise, self._coefs(step), uncond_scale=uncond_scale,
                          temperature=temperature, orig=orig, mask=mask, noise_kn=noise_kn, kn=kn,
                          want_x0=want_aux, want_eps=want_aux, noise_bcast=bcast)

    @torch.no_grad()
    def p_sample(self, x: torch.Tensor, c: torch.Tensor, t: torch.Tensor, step: int,
                 repeat_noise: bool = False, temperature: float = 1.0, uncond_scale: float = 1.0,
                 uncond_cond: Optional[torch.Tensor] = None, cond_concat=None):
        """One reverse step (sampler_sdf.py:80-171): returns (x_{t-1}, predicted x0, eps)."""
        return self._step(x, c, t, step, repeat_noise=repeat_noise, temperature=temperature,
                          uncond_scale=uncond_scale, uncond_cond=uncond_cond, cond_concat=cond_concat)

    @torch.no_grad()
    def q_sample(self, x0: torch.Tensor, index: int, noise: Optional[torch.Tensor] = None):
        """sqrt(alpha_bar_i) x0 + sqrt(1 - alpha_bar_i) noise (sampler_sdf.py:173-192)."""
        if noise is None:
            noise = torch.randn_like(x0, device=x0.device)
        index = int(index)
        return fused_q_sample(x0, noise, self._h["qa"][index], self._h["qb"][index])

    @torch.no_grad()
    def sample(self, shape: List[int], cond: torch.Tensor, repeat_noise: bool = False,
               temperature: float = 1.0, x_last: Optional[torch.Tensor] = None,
               uncond_scale: float = 1.0, uncond_cond: Optional[torch.Tensor] = None, t_start: int = 0):
        """Reverse loop T-1 ... 0 from x_T (sampler_sdf.py:194-255)."""
        bs = shape[0]
        x = x_last if x_last is not None else torch.randn(shape, device=cond.device)
        time_steps = np.flip(self.time_steps)[t_start:]
        if len(time_steps) and self._can_fuse(x, uncond_scale, uncond_cond, None):
            return self._run_fused(x, cond, int(time_steps[0]), len(time_steps), orig=None, mask=None,
                                   temperature=temperature, repeat_noise=repeat_noise,
                                   uncond_scale=uncond_scale, uncond_cond=uncond_cond)
        for step in time_steps:
            ts = x.new_full((bs,), int(step), dtype=torch.long)
            x, _, _ = self._step(x, cond, ts, step, repeat_noise=repeat_noise, temperature=temperature,
                                 uncond_scale=uncond_scale, uncond_cond=uncond_cond, want_aux=False)
        return x

    @torch.no_grad()
    def advance(self, x: torch.Tensor, cond: torch.Tensor, start_step: int, n_steps: int, *,
                orig: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                uncond_scale: float = 1.0, uncond_cond: Optional[torch.Tensor] = None, cond_concat=None):
        """``n_steps`` consecutive iterations of ``paint``'s loop body (sampler_sdf.py:292-341) starting at
        ``start_step``: ``paint(x, cond, t_start, ...) == advance(x, cond, t_start, t_start + 1, ...)``.
        Lets a caller (bench.py, a progress UI) run the chain in slices."""
        n_steps = min(int(n_steps), int(start_step) + 1)
        if self._can_fuse(x, uncond_scale, uncond_cond, cond_concat):
            return self._run_fused(x, cond, int(start_step), n_steps, orig=orig, mask=mask, temperature=1.0,
                                   repeat_noise=False, uncond_scale=uncond_scale, uncond_cond=uncond_cond)
        bs = x.shape[0]
        for step in range(int(start_step), int(start_step) - n_steps, -1):
            ts = x.new_full((bs,), step, dtype=torch.long)
            noise_kn = torch.randn_like(orig, device=orig.device) if (orig is not None and step > 0) else None
            x, _, _ = self._step(x, cond, ts, step, uncond_scale=uncond_scale, uncond_cond=uncond_cond,
                                 cond_concat=cond_concat, orig=orig, mask=mask, noise_kn=noise_kn, want_aux=False)
        return x

    @torch.no_grad()
    def paint(self, x: torch.Tensor, cond: torch.Tensor, t_start: int,
              orig: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
              orig_noise: Optional[torch.Tensor] = None, uncond_scale: float = 1.0,
              uncond_cond: Optional[torch.Tensor] = None, cond_concat=None, repaint_n=1):
        """RePaint-style inpainting loop (sampler_sdf.py:257-350).  As in the reference,
        ``orig_noise`` is accepted and ignored: the known region is re-noised with fresh noise at
        every step, and resampling (repaint_n > 1) re-noises with beta (not sqrt(beta))."""
        bs = x.shape[0]
        time_steps = np.flip(self.time_steps[: t_start + 1])
        print(f"RePainting: sampling steps = {repaint_n}")
        if self._can_fuse(x, uncond_scale, uncond_cond, cond_concat, repaint_n):
            assert orig is None or mask is not None
            return self._run_fused(x, cond, int(t_start), len(time_steps), orig=orig, mask=mask, temperature=1.0,
                                   repeat_noise=False, uncond_scale=uncond_scale, uncond_cond=uncond_cond)
        for step in time_steps:
            step = int(step)
            ts = x.new_full((bs,), step, dtype=torch.long)
            if orig is None:
                x, _, _ = self._step(x, cond, ts, step, uncond_scale=uncond_scale,
                                     uncond_cond=uncond_cond, cond_concat=cond_concat, want_aux=False)
                continue
            assert mask is not None
            x_t = x
            for u in range(repaint_n):
                noise_kn = torch.randn_like(orig, device=orig.device) if step > 0 else None
                x, _, _ = self._step(x_t, cond, ts, step, uncond_scale=uncond_scale,
                                     uncond_cond=uncond_cond, cond_concat=cond_concat, orig=orig,
                                     mask=mask, noise_kn=noise_kn, want_aux=False)
                if u < repaint_n - 1 and step > 0:
                    noise = torch.randn_like(orig, device=orig.device)
                    x_t = fused_q_sample(x, noise, self._h["rp_a"][step - 1], self._h["rp_b"][step - 1])
        return x
