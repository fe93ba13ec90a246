"""TEST INFRASTRUCTURE ONLY -- functional fp32 CPU restatement of the reference UNet forward.

``unet_forward(sd, cfg, x, t, cond)`` evaluates eps_theta(x_t, t, c) from a plain ``state_dict``
(the reference's 556 checkpoint keys) with stock ``torch.nn.functional`` ops on the CPU.  It is a
restatement, not an import: it is pinned against the real reference modules in
tests/test_oracle_vs_reference.py (build container) and against tests/golden/*.npz.

Reference followed (paths relative to /root/reference/polyffusion/):
  stable_diffusion/model/unet.py:151-169 (time_step_embedding), 171-196 (UNetModel.forward),
  207-215 (TimestepEmbedSequential), 231-238 (UpSample), 254-259 (DownSample), 304-318 (ResBlock),
  321-336 (GroupNorm32); stable_diffusion/model/unet_attention.py:61-86 (SpatialTransformer),
  112-124 (BasicTransformerBlock), 186-212 + 261-293 (CrossAttention), 313-333 (FeedForward/GeGLU).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List

import torch
import torch.nn.functional as F


@dataclass
class UNetCfg:
    in_channels: int = 2
    out_channels: int = 2
    channels: int = 64
    n_res_blocks: int = 2
    attention_levels: List[int] = field(default_factory=lambda: [2, 3])
    channel_multipliers: List[int] = field(default_factory=lambda: [1, 2, 4, 4])
    n_heads: int = 4
    tf_layers: int = 1
    d_cond: int = 512


SDF_CHD8BAR = UNetCfg(d_cond=512)   # params/sdf_chd8bar.yaml:9-23
SDF_TXT = UNetCfg(d_cond=1024)      # params/sdf_txt.yaml:9-23
SDF_TXTVNL = UNetCfg(d_cond=128)    # params/sdf_txtvnl.yaml (n_cond = 128)


def block_layout(cfg: UNetCfg):
    """Module layout implied by UNetModel.__init__ (unet.py:70-149): for every input / middle /
    output block, the list of (kind, state_dict prefix) in execution order."""
    levels = len(cfg.channel_multipliers)
    inputs = [[("conv", "input_blocks.0.0")]]
    for lvl in range(levels):
        for _ in range(cfg.n_res_blocks):
            base = f"input_blocks.{len(inputs)}"
            blk = [("res", base + ".0")]
            if lvl in cfg.attention_levels:
                blk.append(("st", base + ".1"))
            inputs.append(blk)
        if lvl != levels - 1:
            inputs.append([("down", f"input_blocks.{len(inputs)}.0")])
    middle = [("res", "middle_block.0"), ("st", "middle_block.1"), ("res", "middle_block.2")]
    outputs = []
    for lvl in reversed(range(levels)):
        for j in range(cfg.n_res_blocks + 1):
            base = f"output_blocks.{len(outputs)}"
            blk = [("res", base + ".0")]
            if lvl in cfg.attention_levels:
                blk.append(("st", f"{base}.{len(blk)}"))
            if lvl != 0 and j == cfg.n_res_blocks:
                blk.append(("up", f"{base}.{len(blk)}"))
            outputs.append(blk)
    return inputs, middle, outputs


def time_freqs(channels: int, max_period: int = 10000) -> torch.Tensor:
    """unet.py:158-164: exp(-ln(max_period) * arange(half) / half), fp32."""
    half = channels // 2
    return torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)


def _gn(x, sd, prefix, eps):
    return F.group_norm(x.float(), 32, sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _res_block(sd, p, x, t_emb):
    # unet.py:304-318
    h = F.conv2d(F.silu(_gn(x, sd, p + ".in_layers.0", 1e-5)), sd[p + ".in_layers.2.weight"],
                 sd[p + ".in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(t_emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    h = h + e[:, :, None, None]
    h = F.conv2d(F.silu(_gn(h, sd, p + ".out_layers.0", 1e-5)), sd[p + ".out_layers.3.weight"],
                 sd[p + ".out_layers.3.bias"], padding=1)
    if p + ".skip_connection.weight" in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def _attention(sd, p, x, ctx, n_heads):
    # unet_attention.py:186-212, 261-293 (normal_attention; softmax over keys, scale d_head**-0.5)
    q = F.linear(x, sd[p + ".to_q.weight"])
    k = F.linear(ctx, sd[p + ".to_k.weight"])
    v = F.linear(ctx, sd[p + ".to_v.weight"])
    b, n, _ = q.shape
    d_head = q.shape[-1] // n_heads
    q = q.view(b, n, n_heads, d_head)
    k = k.view(b, -1, n_heads, d_head)
    v = v.view(b, -1, n_heads, d_head)
    attn = torch.einsum("bihd,bjhd->bhij", q, k) * d_head ** -0.5
    attn = attn.softmax(dim=-1)
    out = torch.einsum("bhij,bjhd->bihd", attn, v).reshape(b, n, -1)
    return F.linear(out, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def _transformer_block(sd, p, x, cond, n_heads):
    # unet_attention.py:112-124 and 313-333
    ln = lambda t, q: F.layer_norm(t, (t.shape[-1],), sd[q + ".weight"], sd[q + ".bias"], 1e-5)
    h = ln(x, p + ".norm1")
    x = _attention(sd, p + ".attn1", h, h, n_heads) + x
    x = _attention(sd, p + ".attn2", ln(x, p + ".norm2"), cond, n_heads) + x
    g = F.linear(ln(x, p + ".norm3"), sd[p + ".ff.net.0.proj.weight"], sd[p + ".ff.net.0.proj.bias"])
    val, gate = g.chunk(2, dim=-1)
    x = F.linear(val * F.gelu(gate), sd[p + ".ff.net.2.weight"], sd[p + ".ff.net.2.bias"]) + x
    return x


def _spatial_transformer(sd, p, x, cond, cfg):
    # unet_attention.py:61-86 (GroupNorm eps 1e-6)
    b, c, h, w = x.shape
    t = F.conv2d(_gn(x, sd, p + ".norm", 1e-6), sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    t = t.permute(0, 2, 3, 1).reshape(b, h * w, c)
    for i in range(cfg.tf_layers):
        t = _transformer_block(sd, f"{p}.transformer_blocks.{i}", t, cond, cfg.n_heads)
    t = t.view(b, h, w, c).permute(0, 3, 1, 2)
    return F.conv2d(t, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"]) + x


def _run_block(sd, blk, x, t_emb, cond, cfg):
    for kind, p in blk:
        if kind == "conv":
            x = F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif kind == "res":
            x = _res_block(sd, p, x, t_emb)
        elif kind == "st":
            x = _spatial_transformer(sd, p, x, cond, cfg)
        elif kind == "down":
            x = F.conv2d(x, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)
        elif kind == "up":
            x = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), sd[p + ".conv.weight"],
                         sd[p + ".conv.bias"], padding=1)
    return x


@torch.no_grad()
def unet_forward(sd: Dict[str, torch.Tensor], cfg: UNetCfg, x: torch.Tensor, t: torch.Tensor,
                 cond: torch.Tensor) -> torch.Tensor:
    """eps_theta(x, t, cond); x [B,Cin,H,W] fp32, t [B] int64, cond [B,n_cond,d_cond]."""
    sd = {k: v.detach().float().cpu() for k, v in sd.items()}
    x, cond = x.float().cpu(), cond.float().cpu()
    args = t.cpu()[:, None].float() * time_freqs(cfg.channels)[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    emb = F.linear(F.silu(F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])),
                   sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    inputs, middle, outputs = block_layout(cfg)
    skips = []
    for blk in inputs:
        x = _run_block(sd, blk, x, emb, cond, cfg)
        skips.append(x)
    x = _run_block(sd, middle, x, emb, cond, cfg)
    for blk in outputs:
        x = _run_block(sd, blk, torch.cat([x, skips.pop()], dim=1), emb, cond, cfg)
    x = F.silu(_gn(x, sd, "out.0", 1e-5))
    return F.conv2d(x, sd["out.2.weight"], sd["out.2.bias"], padding=1)
