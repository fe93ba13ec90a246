"""Parameter containers of the transformer part of the UNet (drop-in names for
``stable_diffusion/model/unet_attention.py`` of the reference).

The classes below own exactly the parameters (same attribute names, same construction order, hence
the same ``state_dict`` keys and the same default initialisation stream) as the reference's
``SpatialTransformer`` (unet_attention.py:26-59), ``BasicTransformerBlock`` (:89-110),
``CrossAttention`` (:127-184), ``FeedForward`` (:296-311) and ``GeGLU`` (:317-327).  They carry no
arithmetic of their own: the whole block is evaluated by the fused CUDA plan behind
``UNetModel.forward`` (LayerNorm folded into the operand transform, fused QKV projection, tcgen05
QK^T / PV, GeGLU, residuals in the GEMM epilogues; n_cond == 1 cross-attention collapsed to a
per-sample vector).
"""
from __future__ import annotations

from torch import nn


class _PlanOnly(nn.Module):
    def forward(self, *args, **kwargs):  # pragma: no cover - defensive
        raise RuntimeError(
            f"{type(self).__name__} is evaluated inside UNetModel's fused CUDA plan; "
            "call UNetModel.forward instead (polyffusion_b200 has no per-module PyTorch path)"
        )


class GeGLU(_PlanOnly):
    def __init__(self, d_in: int, d_out: int):
        super().__init__()
        self.proj = nn.Linear(d_in, d_out * 2)


class FeedForward(_PlanOnly):
    def __init__(self, d_model: int, d_mult: int = 4):
        super().__init__()
        self.net = nn.Sequential(
            GeGLU(d_model, d_model * d_mult), nn.Dropout(0.0), nn.Linear(d_model * d_mult, d_model)
        )


class CrossAttention(_PlanOnly):
    use_flash_attention: bool = False  # kept for attribute compatibility; the CUDA plan ignores it

    def __init__(self, d_model: int, d_cond: int, n_heads: int, d_head: int, is_inplace: bool = True):
        super().__init__()
        self.is_inplace = is_inplace
        self.n_heads = n_heads
        self.d_head = d_head
        self.scale = d_head**-0.5
        d_attn = d_head * n_heads
        self.to_q = nn.Linear(d_model, d_attn, bias=False)
        self.to_k = nn.Linear(d_cond, d_attn, bias=False)
        self.to_v = nn.Linear(d_cond, d_attn, bias=False)
        self.to_out = nn.Sequential(nn.Linear(d_attn, d_model))


class BasicTransformerBlock(_PlanOnly):
    def __init__(self, d_model: int, n_heads: int, d_head: int, d_cond: int):
        super().__init__()
        self.attn1 = CrossAttention(d_model, d_model, n_heads, d_head)
        self.norm1 = nn.LayerNorm(d_model)
        self.attn2 = CrossAttention(d_model, d_cond, n_heads, d_head)
        self.norm2 = nn.LayerNorm(d_model)
        self.ff = FeedForward(d_model)
        self.norm3 = nn.LayerNorm(d_model)


class SpatialTransformer(_PlanOnly):
    def __init__(self, channels: int, n_heads: int, n_layers: int, d_cond: int):
        super().__init__()
        self.norm = nn.GroupNorm(num_groups=32, num_channels=channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(channels, channels, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(channels, n_heads, channels // n_heads, d_cond=d_cond) for _ in range(n_layers)]
        )
        self.proj_out = nn.Conv2d(channels, channels, kernel_size=1, stride=1, padding=0)
