"""GPU parity of the building blocks (through the C ABI) against fp64 torch on the CPU.

Tolerance: the split-bf16 ("bf16x3") GEMM keeps 16 mantissa bits per operand, so one GEMM is
accurate to ~2^-16 relative to the size of the accumulated terms; we require 1e-4 absolute on unit-scale data here,
far inside the end-to-end rtol 1e-3 / atol 1e-4 of BASELINE.json (measured: <= 5e-5 at K = 2304).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref_conv(x_nhwc, w, bias, resid, stride, upsample):
    x = x_nhwc.double().permute(0, 3, 1, 2)
    if upsample:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    y = F.conv2d(x, w.double(), None if bias is None else bias.double(), stride=stride,
                 padding=w.shape[-1] // 2)
    y = y.permute(0, 2, 3, 1)
    if resid is not None:
        y = y + resid.double()
    return y


CASES = [
    # B, H, W, Cin, Cout, k, stride, upsample, bias, resid, force_bn
    (1, 16, 16, 64, 64, 1, 1, False, False, False, 0),
    (2, 16, 16, 64, 64, 3, 1, False, True, True, 0),
    (2, 32, 32, 128, 128, 3, 1, False, True, False, 0),
    (1, 32, 32, 256, 256, 3, 1, False, True, True, 0),
    (1, 32, 32, 256, 256, 3, 1, False, True, True, 128),
    (1, 32, 32, 256, 256, 3, 1, False, True, True, 64),
    (1, 128, 128, 64, 64, 3, 1, False, True, False, 0),
    (1, 64, 64, 192, 128, 3, 1, False, True, False, 0),
    (2, 32, 32, 128, 128, 3, 2, False, True, False, 0),
    (1, 128, 128, 64, 64, 3, 2, False, True, False, 0),
    (2, 16, 16, 256, 256, 3, 1, True, True, False, 0),
    (1, 32, 32, 512, 256, 1, 1, False, True, False, 0),
    (3, 16, 16, 1024, 256, 1, 1, False, True, True, 0),
    (1, 16, 16, 256, 2048, 1, 1, False, True, False, 0),
]


@pytest.mark.parametrize("case", CASES)
def test_conv_gemm(case):
    from polyffusion_b200 import ops

    B, H, W, Cin, Cout, k, stride, up, use_bias, use_res, bn = case
    g = torch.Generator().manual_seed(hash(case) % (2**31))
    x = torch.randn(B, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    bias = torch.randn(Cout, generator=g) if use_bias else None
    Ho, Wo = (2 * H, 2 * W) if up else (H // stride, W // stride)
    resid = torch.randn(B, Ho, Wo, Cout, generator=g) if use_res else None
    ref = _ref_conv(x, w, bias, resid, stride, up)
    out = ops.conv2d_nhwc(x.cuda(), w.cuda(), None if bias is None else bias.cuda(),
                          None if resid is None else resid.cuda(), stride=stride, upsample=up,
                          force_bn=bn).cpu().double()
    err = (out - ref).abs().max().item()
    assert err < 1e-4, f"max abs err {err}"


@pytest.mark.parametrize("B,N,Nk,heads", [(1, 256, 256, 4), (2, 1024, 1024, 4), (2, 256, 128, 4), (1, 1024, 128, 2)])
def test_attention(B, N, Nk, heads):
    from polyffusion_b200 import ops

    g = torch.Generator().manual_seed(B * 1000 + N + Nk)
    C = heads * 64
    q = torch.randn(B, N, C, generator=g)
    k = torch.randn(B, Nk, C, generator=g)
    v = torch.randn(B, Nk, C, generator=g)
    qd, kd, vd = (t.double().view(B, -1, heads, 64) for t in (q, k, v))
    attn = torch.einsum("bihd,bjhd->bhij", qd, kd) * 0.125
    ref = torch.einsum("bhij,bjhd->bihd", attn.softmax(-1), vd).reshape(B, N, C)
    out = ops.attention(q.cuda(), k.cuda(), v.cuda(), heads).cpu().double()
    err = (out - ref).abs().max().item()
    assert err < 1e-4, f"max abs err {err}"


@pytest.mark.parametrize("B,HW,C,silu,eps", [(2, 256, 256, False, 1e-6), (1, 16384, 64, True, 1e-5), (3, 1024, 128, True, 1e-5)])
def test_groupnorm(B, HW, C, silu, eps):
    from polyffusion_b200 import ops

    g = torch.Generator().manual_seed(HW + C)
    x = torch.randn(B, HW, C, generator=g) * 2 + 0.5
    gamma = torch.randn(C, generator=g)
    beta = torch.randn(C, generator=g)
    ref = F.group_norm(x.double().permute(0, 2, 1), 32, gamma.double(), beta.double(), eps).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    out = ops.groupnorm_nhwc(x.cuda(), gamma.cuda(), beta.cuda(), eps, silu).cpu().double()
    err = (out - ref).abs().max().item()
    assert err < 2e-4, f"max abs err {err}"  # operand keeps 16 mantissa bits: ~1.5e-5 * |x| (<~8)
