"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's autoregressive driver pieces.

Reference followed (relative to /root/reference/polyffusion/): inference_sdf.py:121-129
(get_autoreg_data), 132-180 (get_mask "below"/"above", incl. the Python-index wrap-around of the
forward fill), 227-283 (Experiments.predict, autoreg branch: 2B-1 sequential batch-1 paints).
"""
from __future__ import annotations

import torch


def get_autoreg_data(data: torch.Tensor, split_dim: int = 1) -> torch.Tensor:
    steps = data.shape[split_dim]
    half_1, half_2 = data.split(steps // 2, dim=split_dim)
    return torch.cat((half_2, half_1.roll(-1, dims=0)), dim=split_dim)


def get_mask(orig: torch.Tensor, inpaint_type: str) -> torch.Tensor:
    B, _, T, P = orig.shape
    onset = orig[:, 0].reshape(B * T, P)
    if inpaint_type == "below":
        val = onset.argmax(dim=1)
        empty = 0
    elif inpaint_type == "above":
        val = (P - 1) - onset.flip(1).argmax(dim=1)
        empty = P - 1
    else:
        raise NotImplementedError(inpaint_type)
    val = val.clone()
    first = int(val.nonzero()[0])  # IndexError when there is no onset, as in the reference
    val[:first] = val[first]
    for i in range(B * T):
        if val[i] == empty:
            val[i] = val[i - 1]  # i = 0 wraps to the last row (Python indexing), as in the reference
    pitches = torch.arange(P)[None, :]
    m = (pitches >= val[:, None]) if inpaint_type == "below" else (pitches <= val[:, None])
    return m.float().reshape(B, 1, T, P).expand(-1, 2, -1, -1).contiguous()


def predict_autoreg(paint_fn, q_sample_fn, cond, cond_mid, orig, mask, noise, t_idx):
    """One song.  paint_fn(xt, cond_seg, t_idx, orig_seg, mask_seg) -> x0; returns [2B, C, H/2, W]."""
    B, _, H, _ = orig.shape
    half = H // 2
    orig_mid, mask_mid, noise_mid = (get_autoreg_data(t, 2) for t in (orig, mask, noise))
    gen, new_half = [], None
    for idx in range(2 * B - 1):
        src = (cond_mid, orig_mid, mask_mid, noise_mid) if idx % 2 == 1 else (cond, orig, mask, noise)
        cond_seg, orig_seg, mask_seg, noise_seg = (t[idx // 2].unsqueeze(0).clone() for t in src)
        if idx != 0:
            orig_seg[:, :, :half, :] = new_half
            mask_seg[:, :, :half, :] = 1
        xt = q_sample_fn(orig_seg, t_idx, noise_seg)
        x0 = paint_fn(xt, cond_seg, t_idx, orig_seg, mask_seg)
        if idx == 0:
            gen.append(x0[:, :, :half, :])
        new_half = x0[:, :, half:, :]
        gen.append(new_half)
    return torch.cat(gen, dim=0)
