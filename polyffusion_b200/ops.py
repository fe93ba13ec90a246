"""Thin Python wrappers over the building-block entry points of libpf_b200.so (pf_op_*).

These exist for the parity tests: each runs one hot-path building block (tcgen05 split-bf16
conv/linear GEMM, attention core, GroupNorm operand transform) on CUDA tensors through the C ABI.
"""
from __future__ import annotations

import torch

from ._lib import check, current_stream, lib, ptr


def _f32c(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda, "polyffusion_b200 ops need CUDA tensors (no CPU fallback)"
    return t.contiguous().float()


def conv2d_nhwc(x, w, bias=None, resid=None, stride=1, upsample=False, force_bn=0):
    """x [B,H,W,Cin] fp32 NHWC, w [Cout,Cin,k,k] -> [B,Ho,Wo,Cout] (unet.py:229,236,252,282,295,302)."""
    x, w = _f32c(x), _f32c(w)
    B, H, W, Cin = x.shape
    Cout, _, k, _ = w.shape
    Ho, Wo = (2 * H, 2 * W) if upsample else (H // stride, W // stride)
    out = torch.empty((B, Ho, Wo, Cout), device=x.device, dtype=torch.float32)
    bias = None if bias is None else _f32c(bias)
    resid = None if resid is None else _f32c(resid)
    check(lib().pf_op_conv2d_nhwc(ptr(x), B, H, W, Cin, ptr(w), Cout, k, stride, int(upsample),
                                  ptr(bias), ptr(resid), ptr(out), force_bn, current_stream()))
    return out


def attention(q, k, v, heads):
    """softmax(q k^T / 8) v per 64-wide head; q [B,N,heads*64], k/v [B,Nk,heads*64]."""
    q, k, v = _f32c(q), _f32c(k), _f32c(v)
    B, N, C = q.shape
    Nk = k.shape[1]
    out = torch.empty_like(q)
    check(lib().pf_op_attention(ptr(q), ptr(k), ptr(v), B, N, Nk, heads, ptr(out), current_stream()))
    return out


def groupnorm_nhwc(x, gamma, beta, eps=1e-5, silu=False):
    """GroupNorm(32) [+SiLU] over an NHWC tensor [B,HW,C] -> fp32 (hi+lo of the split operand)."""
    x, gamma, beta = _f32c(x), _f32c(gamma), _f32c(beta)
    B, HW, C = x.shape
    out = torch.empty_like(x)
    check(lib().pf_op_groupnorm_nhwc(ptr(x), B, HW, C, ptr(gamma), ptr(beta), eps, int(silu), ptr(out),
                                     current_stream()))
    return out
