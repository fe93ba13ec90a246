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


# ---------------------------------------------------------------------------------------------------
# generic blocks used by the legacy ddpm.unet.UNet drop-in (polyffusion_b200/ddpm/unet.py)
def conv2d_nhwc_vec(x, w, bias=None, resid=None, stride=1):
    """conv2d_nhwc with a per-sample epilogue vector: bias is [Cout] or [B, Cout]."""
    x, w = _f32c(x), _f32c(w)
    B, H, W, Cin = x.shape
    Cout, _, k, _ = w.shape
    out = torch.empty((B, H // stride, W // stride, Cout), device=x.device, dtype=torch.float32)
    bias = None if bias is None else _f32c(bias)
    ld = 0 if bias is None or bias.dim() == 1 else bias.shape[1]
    resid = None if resid is None else _f32c(resid)
    check(lib().pf_op_conv2d_nhwc_ex(ptr(x), B, H, W, Cin, ptr(w), Cout, k, stride, 0, ptr(bias), ld,
                                     ptr(resid), ptr(out), 0, current_stream()))
    return out


def groupnorm_generic(x, groups, gamma, beta, eps=1e-5, silu=False):
    """GroupNorm(groups) [+ Swish] over an NHWC tensor [B,HW,C], any C % groups == 0."""
    x, gamma, beta = _f32c(x), _f32c(gamma), _f32c(beta)
    B, HW, C = x.shape
    out = torch.empty_like(x)
    check(lib().pf_op_groupnorm_generic(ptr(x), B, HW, C, groups, ptr(gamma), ptr(beta), eps, int(silu),
                                        ptr(out), current_stream()))
    return out


def softmax_rows(s, scale=1.0):
    s = _f32c(s)
    out = torch.empty_like(s)
    n = s.shape[-1]
    check(lib().pf_op_softmax_rows(ptr(s), float(scale), ptr(out), s.numel() // n, n, current_stream()))
    return out


def time_sincos(t, freqs):
    """[sin(t f) | cos(t f)] (legacy TimeEmbedding order), t int64 [B], freqs fp32 [half]."""
    assert t.is_cuda
    t = t.contiguous().to(torch.int64)
    freqs = _f32c(freqs)
    out = torch.empty((t.shape[0], 2 * freqs.shape[0]), device=t.device, dtype=torch.float32)
    check(lib().pf_op_time_sincos(ptr(t), ptr(freqs), ptr(out), t.shape[0], freqs.shape[0], current_stream()))
    return out


def conv3x3_direct(x, w, bias, in_nchw: bool, out_nchw: bool):
    """fp32 direct 3x3 convolution (pad 1) for edge layers; layouts per the two flags."""
    x, w = _f32c(x), _f32c(w)
    if in_nchw:
        B, Cin, H, W = x.shape
    else:
        B, H, W, Cin = x.shape
    Cout = w.shape[0]
    shape = (B, Cout, H, W) if out_nchw else (B, H, W, Cout)
    out = torch.empty(shape, device=x.device, dtype=torch.float32)
    bias = None if bias is None else _f32c(bias)
    check(lib().pf_op_conv3x3_direct(ptr(x), ptr(w), ptr(bias), ptr(out), B, Cin, H, W, Cout, int(in_nchw),
                                     int(out_nchw), current_stream()))
    return out


def linear(x, weight, bias=None, act=0):
    """out = act(x W^T + b) on the small-linear kernel (pf_linear); act 0 none, 1 SiLU."""
    x, weight = _f32c(x), _f32c(weight)
    rows, n_in = x.shape
    n_out = weight.shape[0]
    out = torch.empty((rows, n_out), device=x.device, dtype=torch.float32)
    bias = None if bias is None else _f32c(bias)
    check(lib().pf_linear(ptr(x), n_in, ptr(weight), ptr(bias), ptr(out), n_out, rows, n_out, n_in, act,
                          current_stream()))
    return out
