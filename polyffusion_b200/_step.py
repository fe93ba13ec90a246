"""Shared launcher for the fused sampler-step kernels (pf_sample_step_* in include/pf_b200.h)."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from ._lib import StepArgs, check, current_stream, lib


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _c(t: Optional[torch.Tensor], like: torch.Tensor) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        t = t.to(like.device)
    return t.contiguous().float()


def fused_step(kind: str, x, e_cond, e_uncond, noise, coefs, *, uncond_scale=1.0, temperature=1.0,
               orig=None, mask=None, noise_kn=None, kn=(0.0, 0.0), want_x0=True, want_eps=True,
               noise_bcast=0):
    """One reverse-diffusion step epilogue on the GPU: CFG combine + x0 + x_{t-1} (+ RePaint blend)."""
    if not x.is_cuda:
        raise RuntimeError("polyffusion_b200 samplers run on CUDA tensors only (no CPU fallback)")
    x = x.contiguous().float()
    e_cond = _c(e_cond, x)
    e_uncond = _c(e_uncond, x)
    noise = _c(noise, x)
    if orig is not None:
        orig = _c(orig.expand_as(x) if orig.shape != x.shape else orig, x)
        mask = _c(mask.expand_as(x) if mask.shape != x.shape else mask, x)
        noise_kn = _c(noise_kn, x)
    x_prev = torch.empty_like(x)
    x0 = torch.empty_like(x) if want_x0 else None
    e_t = (torch.empty_like(x) if e_uncond is not None else None) if want_eps else None
    a = StepArgs()
    a.x, a.e_cond, a.e_uncond, a.noise = _p(x), _p(e_cond), _p(e_uncond), _p(noise)
    a.orig, a.mask, a.noise_kn = _p(orig), _p(mask), _p(noise_kn)
    a.x_prev, a.x0, a.e_t = _p(x_prev), _p(x0), _p(e_t)
    a.n = x.numel()
    a.noise_bcast = int(noise_bcast)
    a.uncond_scale = float(uncond_scale)
    a.c0, a.c1, a.c2, a.c3, a.c4 = (float(v) for v in coefs)
    a.temperature = float(temperature)
    a.kn_a, a.kn_b = float(kn[0]), float(kn[1])
    fn = {"ddpm": lib().pf_sample_step_ddpm, "ddim": lib().pf_sample_step_ddim,
          "legacy": lib().pf_sample_step_ddpm_legacy}[kind]
    with torch.cuda.device(x.device):
        check(fn(ctypes.byref(a), current_stream()))
    if want_eps and e_t is None:
        e_t = e_cond
    return x_prev, x0, e_t


def fused_q_sample(x0: torch.Tensor, noise: torch.Tensor, a: float, b: float) -> torch.Tensor:
    if not x0.is_cuda:
        raise RuntimeError("polyffusion_b200 samplers run on CUDA tensors only (no CPU fallback)")
    x0 = x0.contiguous().float()
    noise = _c(noise, x0)
    out = torch.empty_like(x0)
    from ._lib import ptr

    with torch.cuda.device(x0.device):
        check(lib().pf_q_sample(ptr(x0), ptr(noise), ptr(out), out.numel(), float(a), float(b),
                                current_stream()))
    return out
