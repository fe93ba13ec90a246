"""ctypes glue shared by the encoder modules (C ABI: include/pf_b200.h, "Condition encoders")."""
from __future__ import annotations

import ctypes

import torch

from .._lib import check, current_stream, lib, ptr


def _dev32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def require_cuda(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"polyffusion_b200.{what} runs on CUDA tensors only (no CPU fallback)")


def linear(x: torch.Tensor, lin: torch.nn.Linear, act: int = 0) -> torch.Tensor:
    """act(x @ W^T + b) over the last dim; act 0 none, 2 exp."""
    x2 = x.reshape(-1, x.shape[-1]).contiguous()
    w, b = _dev32(lin.weight, x.device), _dev32(lin.bias, x.device)
    out = torch.empty((x2.shape[0], w.shape[0]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().pf_linear(ptr(x2), x2.shape[1], ptr(w), ptr(b), ptr(out), out.shape[1], x2.shape[0],
                              w.shape[0], w.shape[1], act, current_stream()))
    return out.reshape(*x.shape[:-1], w.shape[0])


def gru_bidir_last(x: torch.Tensor, gru: torch.nn.GRU) -> torch.Tensor:
    """``gru(x)[-1].transpose(0, 1).reshape(B, 2H)`` for a 1-layer bidirectional batch_first GRU."""
    if not (gru.num_layers == 1 and gru.bidirectional and gru.batch_first and gru.bias):
        raise NotImplementedError("only the reference's GRU(batch_first=True, bidirectional=True) is supported")
    B, T, I = x.shape
    H = gru.hidden_size
    dev = x.device
    names = [("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"),
             ("weight_ih_l0_reverse", "weight_hh_l0_reverse", "bias_ih_l0_reverse", "bias_hh_l0_reverse")]
    keep = [[_dev32(getattr(gru, n), dev) for n in group] for group in names]
    arr = lambda k: (ctypes.c_void_p * 2)(keep[0][k].data_ptr(), keep[1][k].data_ptr())
    out = torch.empty((B, 2 * H), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib().pf_gru_workspace_bytes(B, T, H)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        xc = x.contiguous().float()
        check(lib().pf_gru_bidir_last(ptr(xc), B, T, I, H, arr(0), arr(1), arr(2), arr(3), ptr(out), ptr(ws),
                                      nbytes, current_stream()))
    return out


def txt_cnn(pr: torch.Tensor, conv: torch.nn.Conv2d) -> torch.Tensor:
    B, T, P = pr.shape
    C = conv.out_channels
    if conv.kernel_size != (4, 12) or conv.stride != (4, 1) or conv.in_channels != 1:
        raise NotImplementedError("TextureEncoder.cnn geometry other than Conv2d(1, C, (4, 12), stride (4, 1))")
    w, b = _dev32(conv.weight, pr.device), _dev32(conv.bias, pr.device)
    out = torch.empty((B, C, T // 4, (P - 11) // 4), dtype=torch.float32, device=pr.device)
    with torch.cuda.device(pr.device):
        check(lib().pf_txt_cnn(ptr(pr.contiguous().float()), ptr(w), ptr(b), ptr(out), B, C, T, P, current_stream()))
    return out
