"""``UNetModel`` -- drop-in for ``stable_diffusion/model/unet.py`` of the reference.

Same constructor signature (unet.py:35-47), same parameter names / shapes / construction order (so
reference checkpoints load with ``load_state_dict`` and a fixed seed gives the same random
initialisation), same ``forward(x, time_steps, cond)`` contract (unet.py:171-196).  The arithmetic
is not PyTorch: ``forward`` hands the three tensors to ``libpf_b200.so`` (``pf_unet_forward``), which
replays a static plan of hand-written sm_100a kernels (CUDA tensors only, no CPU fallback).  When
gradients are required (training: ``LatentDiffusion.loss`` under the reference's ``learner.py``), the
same parameters are evaluated by the differentiable PyTorch graph in ``unet_torch.py`` instead.
"""
from __future__ import annotations

from typing import List

import torch
from torch import nn

from polyffusion_b200.engine import UNetEngine
from polyffusion_b200.stable_diffusion.model.unet_attention import SpatialTransformer, _PlanOnly


class GroupNorm32(nn.GroupNorm):
    """Parameter holder for GroupNorm(32, C), eps 1e-5 (unet.py:321-336)."""


def normalization(channels):
    return GroupNorm32(32, channels)


class UpSample(_PlanOnly):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)


class DownSample(_PlanOnly):
    def __init__(self, channels: int):
        super().__init__()
        self.op = nn.Conv2d(channels, channels, 3, stride=2, padding=1)


class ResBlock(_PlanOnly):
    def __init__(self, channels: int, d_t_emb: int, *, out_channels=None):
        super().__init__()
        if out_channels is None:
            out_channels = channels
        self.in_layers = nn.Sequential(
            normalization(channels), nn.SiLU(), nn.Conv2d(channels, out_channels, 3, padding=1)
        )
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(d_t_emb, out_channels))
        self.out_layers = nn.Sequential(
            normalization(out_channels), nn.SiLU(), nn.Dropout(0.0),
            nn.Conv2d(out_channels, out_channels, 3, padding=1),
        )
        if out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = nn.Conv2d(channels, out_channels, 1)


class TimestepEmbedSequential(nn.Sequential):
    """Container only; dispatch by layer type happens inside the CUDA plan (unet.py:199-215)."""

    def forward(self, *args, **kwargs):  # pragma: no cover - defensive
        raise RuntimeError("TimestepEmbedSequential is evaluated inside UNetModel's fused CUDA plan")


class UNetModel(nn.Module):
    def __init__(
        self,
        *,
        in_channels: int,
        out_channels: int,
        channels: int,
        n_res_blocks: int,
        attention_levels: List[int],
        channel_multipliers: List[int],
        n_heads: int,
        tf_layers: int = 1,
        d_cond: int = 768,
    ):
        super().__init__()
        self.channels = channels
        self._cfg = dict(
            in_channels=int(in_channels), out_channels=int(out_channels), channels=int(channels),
            n_res_blocks=int(n_res_blocks), attention_levels=[int(a) for a in attention_levels],
            channel_multipliers=[int(m) for m in channel_multipliers], n_heads=int(n_heads),
            tf_layers=int(tf_layers), d_cond=int(d_cond),
        )
        levels = len(channel_multipliers)
        d_time_emb = channels * 4
        self.time_embed = nn.Sequential(
            nn.Linear(channels, d_time_emb), nn.SiLU(), nn.Linear(d_time_emb, d_time_emb)
        )
        self.input_blocks = nn.ModuleList()
        self.input_blocks.append(TimestepEmbedSequential(nn.Conv2d(in_channels, channels, 3, padding=1)))
        input_block_channels = [channels]
        channels_list = [channels * m for m in channel_multipliers]
        for i in range(levels):
            for _ in range(n_res_blocks):
                layers = [ResBlock(channels, d_time_emb, out_channels=channels_list[i])]
                channels = channels_list[i]
                if i in attention_levels:
                    layers.append(SpatialTransformer(channels, n_heads, tf_layers, d_cond))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                input_block_channels.append(channels)
            if i != levels - 1:
                self.input_blocks.append(TimestepEmbedSequential(DownSample(channels)))
                input_block_channels.append(channels)
        self.middle_block = TimestepEmbedSequential(
            ResBlock(channels, d_time_emb),
            SpatialTransformer(channels, n_heads, tf_layers, d_cond),
            ResBlock(channels, d_time_emb),
        )
        self.output_blocks = nn.ModuleList([])
        for i in reversed(range(levels)):
            for j in range(n_res_blocks + 1):
                layers = [
                    ResBlock(channels + input_block_channels.pop(), d_time_emb, out_channels=channels_list[i])
                ]
                channels = channels_list[i]
                if i in attention_levels:
                    layers.append(SpatialTransformer(channels, n_heads, tf_layers, d_cond))
                if i != 0 and j == n_res_blocks:
                    layers.append(UpSample(channels))
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(
            normalization(channels), nn.SiLU(), nn.Conv2d(channels, out_channels, 3, padding=1)
        )
        self._engine = None

    # -- kept for API compatibility (unet.py:151-169); used by nothing on the CUDA path
    def time_step_embedding(self, time_steps: torch.Tensor, max_period: int = 10000):
        import math

        half = self.channels // 2
        frequencies = torch.exp(
            -math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half
        ).to(device=time_steps.device)
        args = time_steps[:, None].float() * frequencies[None]
        return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)

    @property
    def engine(self) -> UNetEngine:
        if self._engine is None:
            object.__setattr__(self, "_engine", UNetEngine(self, self._cfg))
        return self._engine

    def forward(self, x: torch.Tensor, time_steps: torch.Tensor, cond: torch.Tensor):
        """eps_theta(x_t, t, c): x [B,C,H,W], time_steps [B] (long), cond [B,n_cond,d_cond]."""
        if self._wants_autograd(x, cond):
            from polyffusion_b200.stable_diffusion.model.unet_torch import unet_forward_torch

            return unet_forward_torch(self, x, time_steps, cond)
        return self.engine.forward(x, time_steps, cond)

    def _wants_autograd(self, x: torch.Tensor, cond: torch.Tensor) -> bool:
        """True when the caller needs a differentiable result: grad mode is on and either the module is
        in training mode with trainable parameters, or an input requires grad.  Sampling (the samplers
        are ``@torch.no_grad()``, inference_sdf.py calls ``model.eval()``) never satisfies this."""
        if not torch.is_grad_enabled():
            return False
        if x.requires_grad or (cond is not None and cond.requires_grad):
            return True
        return self.training and any(p.requires_grad for p in self.parameters())

    # the engine holds a ctypes handle and device buffers: never copied / pickled with the module
    # (copy.deepcopy for EMA, torch.save(model)); the copy builds its own engine lazily
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None
        return state

    def __deepcopy__(self, memo):
        import copy

        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            object.__setattr__(new, k, None if k == "_engine" else copy.deepcopy(v, memo))
        return new
