"""``UNet`` -- drop-in for ``ddpm/unet.py`` of the reference (the legacy unconditional eps-model,
SURVEY.md section 8a row D2; BASELINE config 1 uses ``UNet(2, 64, [1, 2, 2, 4], [F, F, F, T])``).

Same constructor, module tree, parameter names / shapes and construction order as ddpm/unet.py:305-407
(so reference checkpoints load with ``load_state_dict`` and a fixed seed reproduces the reference's
initialisation), including its quirks: compounding channel multipliers (64 -> 64 -> 128 -> 256 -> 1024,
:360), sin-then-cos timestep embedding over ``half - 1`` (:62-67), the ResidualBlock time projection
WITHOUT an activation (:141), single-head attention with ``d_k = channels`` whose GroupNorm is
constructed but never applied (:163-205), ``ConvTranspose2d(4, 2, 1)`` upsampling (:288-301) and the
final ``GroupNorm(8, 64)`` (:404).

``forward(x, t)`` under ``torch.no_grad()`` on CUDA tensors composes the library's kernels through the C
ABI: every convolution / projection runs on the tcgen05 implicit-GEMM kernel (``pf_op_conv2d_nhwc_ex``;
the attention products q k^T and p v are per-sample 1x1 "convolutions"), GroupNorm + Swish, the row
softmax, the timestep embedding and the two edge convolutions on the generic kernels
(``pf_op_groupnorm_generic``, ``pf_op_softmax_rows``, ``pf_op_time_sincos``, ``pf_op_conv3x3_direct``,
``pf_linear``).  PyTorch only moves data (NCHW <-> NHWC views, channel concat, parity scatter of the
transposed convolution).  This is the plumbing path of config 1, not the optimised sdf path: every op
allocates its scratch and synchronises.  With gradients enabled (``train/train_ddpm.py``) the same
parameters are evaluated by the differentiable PyTorch graph in ``_forward_torch``.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple, Union

import torch
import torch.nn.functional as F
from torch import nn


class Swish(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


class TimeEmbedding(nn.Module):
    def __init__(self, n_channels: int):
        super().__init__()
        self.n_channels = n_channels
        self.lin1 = nn.Linear(self.n_channels // 4, self.n_channels)
        self.act = Swish()
        self.lin2 = nn.Linear(self.n_channels, self.n_channels)

    def frequencies(self, device) -> torch.Tensor:
        # ddpm/unet.py:63-65, evaluated with the same torch expression
        half_dim = self.n_channels // 8
        emb = math.log(10_000) / (half_dim - 1)
        return torch.exp(torch.arange(half_dim, device=device) * -emb)

    def forward(self, t: torch.Tensor):
        emb = t[:, None] * self.frequencies(t.device)[None, :]
        emb = torch.cat((emb.sin(), emb.cos()), dim=1)
        return self.lin2(self.act(self.lin1(emb)))


class ResidualBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, time_channels: int, n_groups: int = 32):
        super().__init__()
        self.norm1 = nn.GroupNorm(n_groups, in_channels)
        self.act1 = Swish()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=(3, 3), padding=(1, 1))
        self.norm2 = nn.GroupNorm(n_groups, out_channels)
        self.act2 = Swish()
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=(3, 3), padding=(1, 1))
        if in_channels != out_channels:
            self.shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=(1, 1))
        else:
            self.shortcut = nn.Identity()
        self.time_emb = nn.Linear(time_channels, out_channels)

    def forward(self, x: torch.Tensor, t: torch.Tensor):
        h = self.conv1(self.act1(self.norm1(x)))
        h = h + self.time_emb(t)[:, :, None, None]
        h = self.conv2(self.act2(self.norm2(h)))
        return h + self.shortcut(x)


class AttentionBlock(nn.Module):
    def __init__(self, n_channels: int, n_heads: int = 1, d_k: int = None, n_groups: int = 32):
        super().__init__()
        if d_k is None:
            d_k = n_channels
        self.norm = nn.GroupNorm(n_groups, n_channels)  # owned but never applied (reference quirk)
        self.projection = nn.Linear(n_channels, n_heads * d_k * 3)
        self.output = nn.Linear(n_heads * d_k, n_channels)
        self.scale = d_k**-0.5
        self.n_heads = n_heads
        self.d_k = d_k

    def forward(self, x: torch.Tensor, t: Optional[torch.Tensor] = None):
        b, c, h, w = x.shape
        x = x.view(b, c, -1).permute(0, 2, 1)
        qkv = self.projection(x).view(b, -1, self.n_heads, 3 * self.d_k)
        q, k, v = torch.chunk(qkv, 3, dim=-1)
        attn = (torch.einsum("bihd,bjhd->bijh", q, k) * self.scale).softmax(dim=2)
        res = torch.einsum("bijh,bjhd->bihd", attn, v).reshape(b, -1, self.n_heads * self.d_k)
        res = self.output(res) + x
        return res.permute(0, 2, 1).view(b, c, h, w)


class DownBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, time_channels: int, has_attn: bool):
        super().__init__()
        self.res = ResidualBlock(in_channels, out_channels, time_channels)
        self.attn = AttentionBlock(out_channels) if has_attn else nn.Identity()

    def forward(self, x: torch.Tensor, t: torch.Tensor):
        return self.attn(self.res(x, t))


class UpBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, time_channels: int, has_attn: bool):
        super().__init__()
        self.res = ResidualBlock(in_channels + out_channels, out_channels, time_channels)
        self.attn = AttentionBlock(out_channels) if has_attn else nn.Identity()

    def forward(self, x: torch.Tensor, t: torch.Tensor):
        return self.attn(self.res(x, t))


class MiddleBlock(nn.Module):
    def __init__(self, n_channels: int, time_channels: int):
        super().__init__()
        self.res1 = ResidualBlock(n_channels, n_channels, time_channels)
        self.attn = AttentionBlock(n_channels)
        self.res2 = ResidualBlock(n_channels, n_channels, time_channels)

    def forward(self, x: torch.Tensor, t: torch.Tensor):
        return self.res2(self.attn(self.res1(x, t)), t)


class Upsample(nn.Module):
    def __init__(self, n_channels):
        super().__init__()
        self.conv = nn.ConvTranspose2d(n_channels, n_channels, (4, 4), (2, 2), (1, 1))

    def forward(self, x: torch.Tensor, t: torch.Tensor):
        return self.conv(x)


class Downsample(nn.Module):
    def __init__(self, n_channels):
        super().__init__()
        self.conv = nn.Conv2d(n_channels, n_channels, (3, 3), (2, 2), (1, 1))

    def forward(self, x: torch.Tensor, t: torch.Tensor):
        return self.conv(x)


class UNet(nn.Module):
    def __init__(self, image_channels: int = 3, n_channels: int = 64,
                 ch_mults: Union[Tuple[int, ...], List[int]] = (1, 2, 2, 4),
                 is_attn: Union[Tuple[bool, ...], List[int]] = (False, False, True, True),
                 n_blocks: int = 2):
        super().__init__()
        n_resolutions = len(ch_mults)
        self.image_proj = nn.Conv2d(image_channels, n_channels, kernel_size=(3, 3), padding=(1, 1))
        self.time_emb = TimeEmbedding(n_channels * 4)
        down = []
        out_channels = in_channels = n_channels
        for i in range(n_resolutions):
            out_channels = in_channels * ch_mults[i]
            for _ in range(n_blocks):
                down.append(DownBlock(in_channels, out_channels, n_channels * 4, is_attn[i]))
                in_channels = out_channels
            if i < n_resolutions - 1:
                down.append(Downsample(in_channels))
        self.down = nn.ModuleList(down)
        self.middle = MiddleBlock(out_channels, n_channels * 4)
        up = []
        in_channels = out_channels
        for i in reversed(range(n_resolutions)):
            out_channels = in_channels
            for _ in range(n_blocks):
                up.append(UpBlock(in_channels, out_channels, n_channels * 4, is_attn[i]))
            out_channels = in_channels // ch_mults[i]
            up.append(UpBlock(in_channels, out_channels, n_channels * 4, is_attn[i]))
            in_channels = out_channels
            if i > 0:
                up.append(Upsample(in_channels))
        self.up = nn.ModuleList(up)
        self.norm = nn.GroupNorm(8, n_channels)
        self.act = Swish()
        self.final = nn.Conv2d(in_channels, image_channels, kernel_size=(3, 3), padding=(1, 1))

    # ------------------------------------------------------------------ dispatch
    def forward(self, x: torch.Tensor, t: torch.Tensor):
        """eps_theta(x_t, t): x [B, C, H, W], t [B] (ddpm/unet.py:410-444)."""
        if torch.is_grad_enabled() and (x.requires_grad or (
                self.training and any(p.requires_grad for p in self.parameters()))):
            return self._forward_torch(x, t)
        if not x.is_cuda:
            raise RuntimeError("polyffusion_b200.ddpm.unet.UNet evaluates on CUDA tensors only under "
                               "torch.no_grad() (no CPU fallback); autograd runs its PyTorch graph")
        return self._forward_cuda(x, t)

    def _forward_torch(self, x: torch.Tensor, t: torch.Tensor):
        t = self.time_emb(t)
        x = self.image_proj(x)
        h = [x]
        for m in self.down:
            x = m(x, t)
            h.append(x)
        x = self.middle(x, t)
        for m in self.up:
            if isinstance(m, Upsample):
                x = m(x, t)
            else:
                x = m(torch.cat((x, h.pop()), dim=1), t)
        return self.final(self.act(self.norm(x)))

    # ------------------------------------------------------------------ CUDA composition (NHWC)
    @torch.no_grad()
    def _forward_cuda(self, x: torch.Tensor, t: torch.Tensor):
        from polyffusion_b200 import ops

        def res_block(rb: ResidualBlock, xin, temb):
            B, H, W, C = xin.shape
            a = ops.groupnorm_generic(xin.view(B, H * W, C), rb.norm1.num_groups, rb.norm1.weight,
                                      rb.norm1.bias, rb.norm1.eps, True).view(B, H, W, C)
            # conv1 bias + time_emb(t): one per-sample epilogue vector (the Linear's bias absorbs conv1's)
            vec = ops.linear(temb, rb.time_emb.weight, rb.time_emb.bias + rb.conv1.bias)
            hmid = ops.conv2d_nhwc_vec(a, rb.conv1.weight, vec)
            Co = hmid.shape[-1]
            a2 = ops.groupnorm_generic(hmid.view(B, H * W, Co), rb.norm2.num_groups, rb.norm2.weight,
                                       rb.norm2.bias, rb.norm2.eps, True).view(B, H, W, Co)
            if isinstance(rb.shortcut, nn.Identity):
                sc = xin
            else:
                sc = ops.conv2d_nhwc_vec(xin, rb.shortcut.weight, rb.shortcut.bias)
            return ops.conv2d_nhwc_vec(a2, rb.conv2.weight, rb.conv2.bias, resid=sc)

        def attention(ab: AttentionBlock, xin):
            B, H, W, C = xin.shape
            if ab.n_heads != 1:
                raise NotImplementedError("legacy AttentionBlock: only n_heads == 1 (the reference default)")
            N, d = H * W, ab.d_k
            qkv = ops.conv2d_nhwc_vec(xin, ab.projection.weight[:, :, None, None], ab.projection.bias)
            qkv = qkv.view(B, N, 3 * d)
            outs = []
            for b in range(B):
                q, k, v = qkv[b, :, :d], qkv[b, :, d:2 * d], qkv[b, :, 2 * d:]
                # scores[i, j] = q_i . k_j: a 1x1 "convolution" of the N query pixels with k as weights
                s = ops.conv2d_nhwc_vec(q.reshape(1, 1, N, d), k.reshape(N, d, 1, 1)).view(N, N)
                p = ops.softmax_rows(s, ab.scale)  # softmax over j (dim=2 of 'bijh')
                o = ops.conv2d_nhwc_vec(p.view(1, 1, N, N), v.t().reshape(d, N, 1, 1))
                outs.append(o.view(1, N, d))
            res = torch.cat(outs).view(B, H, W, d)
            return ops.conv2d_nhwc_vec(res, ab.output.weight[:, :, None, None], ab.output.bias, resid=xin)

        def upsample(up: Upsample, xin):
            # ConvTranspose2d(4, 2, 1): out[2y+py, 2x+px] = sum over a 2x2 window of the input; each
            # output parity is a 3x3 convolution whose other taps are zero.  w: [Cin, Cout, 4, 4]
            B, H, W, C = xin.shape
            wt = up.conv.weight
            Co = wt.shape[1]
            out = torch.empty((B, 2 * H, 2 * W, Co), device=xin.device, dtype=torch.float32)
            for py in range(2):
                for px in range(2):
                    w3 = torch.zeros((Co, C, 3, 3), device=xin.device, dtype=torch.float32)
                    for dy, ky in (((-1, 3), (0, 1)) if py == 0 else ((0, 2), (1, 0))):
                        for dx, kx in (((-1, 3), (0, 1)) if px == 0 else ((0, 2), (1, 0))):
                            w3[:, :, dy + 1, dx + 1] = wt[:, :, ky, kx].t()
                    out[:, py::2, px::2, :] = ops.conv2d_nhwc_vec(xin, w3, up.conv.bias)
            return out

        def block(m, xin, temb):
            if isinstance(m, Upsample):
                return upsample(m, xin)
            if isinstance(m, Downsample):
                return ops.conv2d_nhwc_vec(xin, m.conv.weight, m.conv.bias, stride=2)
            y = res_block(m.res, xin, temb)
            return y if isinstance(m.attn, nn.Identity) else attention(m.attn, y)

        te = self.time_emb
        emb = ops.time_sincos(t, te.frequencies(x.device))
        temb = ops.linear(ops.linear(emb, te.lin1.weight, te.lin1.bias, act=1), te.lin2.weight, te.lin2.bias)
        h = ops.conv3x3_direct(x, self.image_proj.weight, self.image_proj.bias, in_nchw=True, out_nchw=False)
        skips = [h]
        for m in self.down:
            h = block(m, h, temb)
            skips.append(h)
        h = res_block(self.middle.res1, h, temb)
        h = attention(self.middle.attn, h)
        h = res_block(self.middle.res2, h, temb)
        for m in self.up:
            if isinstance(m, Upsample):
                h = upsample(m, h)
            else:
                h = block(m, torch.cat((h, skips.pop()), dim=3), temb)
        B, H, W, C = h.shape
        a = ops.groupnorm_generic(h.view(B, H * W, C), self.norm.num_groups, self.norm.weight, self.norm.bias,
                                  self.norm.eps, True).view(B, H, W, C)
        return ops.conv3x3_direct(a, self.final.weight, self.final.bias, in_nchw=False, out_nchw=True)
