"""Autograd-capable PyTorch evaluation of ``UNetModel`` over the drop-in's own parameter containers.

Used ONLY when gradients are required (training: ``LatentDiffusion.loss``, the reference's
``learner.py`` / ``lightning_learner.py`` drive the model unchanged).  The
sampling hot path (``torch.no_grad()``, CUDA tensors) never comes here -- it runs the CUDA plan behind
``pf_unet_forward`` and raises when the library or a GPU is missing.

Follows stable_diffusion/model/unet.py:171-196 (forward), :207-215 (TimestepEmbedSequential dispatch),
:231-238 (UpSample), :254-259 (DownSample), :304-318 (ResBlock), :321-336 (GroupNorm32: fp32
statistics) and stable_diffusion/model/unet_attention.py:61-86, 112-124, 186-212, 261-293, 313-333 of
the reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _group_norm32(norm, x):
    # GroupNorm32.forward (unet.py:334-336): statistics in fp32, result cast back
    return F.group_norm(x.float(), norm.num_groups, norm.weight, norm.bias, norm.eps).type(x.dtype)


def _res_block(blk, x, t_emb):
    gn1, _, conv1 = blk.in_layers
    h = conv1(F.silu(_group_norm32(gn1, x)))
    e = blk.emb_layers[1](F.silu(t_emb)).type(h.dtype)
    h = h + e[:, :, None, None]
    gn2, _, drop, conv2 = blk.out_layers
    h = conv2(drop(F.silu(_group_norm32(gn2, h))))
    return blk.skip_connection(x) + h


def _attention(attn, x, cond):
    # CrossAttention.forward / normal_attention (unet_attention.py:186-212, 261-293)
    ctx = x if cond is None else cond
    q, k, v = attn.to_q(x), attn.to_k(ctx), attn.to_v(ctx)
    b, n, _ = q.shape
    h = attn.n_heads
    q = q.view(b, n, h, -1)
    k = k.view(b, k.shape[1], h, -1)
    v = v.view(b, v.shape[1], h, -1)
    scores = torch.einsum("bihd,bjhd->bhij", q, k) * attn.scale
    out = torch.einsum("bhij,bjhd->bihd", scores.softmax(dim=-1), v).reshape(b, n, -1)
    return attn.to_out(out)


def _transformer_block(tb, x, cond):
    x = _attention(tb.attn1, tb.norm1(x), None) + x
    x = _attention(tb.attn2, tb.norm2(x), cond) + x
    geglu, drop, lin = tb.ff.net
    val, gate = geglu.proj(tb.norm3(x)).chunk(2, dim=-1)
    return lin(drop(val * F.gelu(gate))) + x


def _spatial_transformer(st, x, cond):
    b, c, h, w = x.shape
    t = st.proj_in(F.group_norm(x, st.norm.num_groups, st.norm.weight, st.norm.bias, st.norm.eps))
    t = t.permute(0, 2, 3, 1).reshape(b, h * w, c)
    for tb in st.transformer_blocks:
        t = _transformer_block(tb, t, cond)
    t = t.view(b, h, w, c).permute(0, 3, 1, 2)
    return st.proj_out(t) + x


def _sequential(seq, x, t_emb, cond):
    # TimestepEmbedSequential.forward (unet.py:207-215): dispatch on the layer type
    from polyffusion_b200.stable_diffusion.model.unet import DownSample, ResBlock, UpSample
    from polyffusion_b200.stable_diffusion.model.unet_attention import SpatialTransformer

    for layer in seq:
        if isinstance(layer, ResBlock):
            x = _res_block(layer, x, t_emb)
        elif isinstance(layer, SpatialTransformer):
            x = _spatial_transformer(layer, x, cond)
        elif isinstance(layer, UpSample):
            x = layer.conv(F.interpolate(x, scale_factor=2, mode="nearest"))
        elif isinstance(layer, DownSample):
            x = layer.op(x)
        else:
            x = layer(x)
    return x


def unet_forward_torch(model, x: torch.Tensor, time_steps: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
    """eps_theta(x_t, t, c) as a differentiable PyTorch graph (unet.py:171-196)."""
    t_emb = model.time_embed(model.time_step_embedding(time_steps))
    skips = []
    for module in model.input_blocks:
        x = _sequential(module, x, t_emb, cond)
        skips.append(x)
    x = _sequential(model.middle_block, x, t_emb, cond)
    for module in model.output_blocks:
        x = _sequential(module, torch.cat([x, skips.pop()], dim=1), t_emb, cond)
    gn, _, conv = model.out
    return conv(F.silu(_group_norm32(gn, x)))
