"""Condition glue of ``Polyffusion_SDF`` (reference ``models/model_sdf.py:92-104, 153-164``): latent
conditions for the UNet's cross-attention from chord matrices / piano rolls, batched."""
from __future__ import annotations

import torch


def encode_chord(chord_enc, chord: torch.Tensor) -> torch.Tensor:
    """``_encode_chord`` (model_sdf.py:92-104): chord [B, 32, 36] -> [B, 1, z]; without an encoder the
    chord matrix is flattened."""
    if chord_enc is not None:
        return chord_enc(chord).mean.unsqueeze(1)
    return torch.reshape(chord, (-1, 1, chord.shape[1] * chord.shape[2]))


def encode_txt(txt_enc, prmat: torch.Tensor) -> torch.Tensor:
    """``_encode_txt`` (model_sdf.py:153-164): prmat [B, 128, 128] -> [B, 1, 4 * z]: the reference
    encodes the four 32-step segments one after the other and concatenates; here they are one batch."""
    if txt_enc is None:
        return prmat
    B, T, P = prmat.shape
    seg = T // 32
    z = txt_enc(prmat.reshape(B * seg, 32, P)).mean          # [B * seg, z], segment-minor
    return z.reshape(B, seg * z.shape[-1]).unsqueeze(1)
