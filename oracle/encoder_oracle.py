"""TEST INFRASTRUCTURE ONLY -- CPU restatement (functional fp32 PyTorch) of the reference's condition
encoders.

Reference followed (relative to /root/reference/polyffusion/): dl_modules/chord_enc.py:5-22
(RnnEncoder: bidirectional GRU -> final hidden states -> linear_mu, exp(linear_var)),
dl_modules/txt_enc.py:5-35 (TextureEncoder: Conv2d(1,C,(4,12),(4,1)) -> ReLU -> MaxPool(1,4) ->
view(bs, 8, -1) -> fc1 -> fc2 -> bidirectional GRU -> heads), models/model_sdf.py:92-104, 153-164
(_encode_chord / _encode_txt).  The GRU is written out gate by gate (torch.nn.GRU semantics:
r, z, n order; n = tanh(W_in x + b_in + r * (W_hn h + b_hn))).

Pin status: pinned -- tests/test_encoders.py compares this restatement with golden outputs of the
real reference modules (loaded by file path by oracle/make_golden.py; the dl_modules package itself
needs pretty_midi) and, where /root/reference exists, with the modules directly.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def gru_bidir_last(x, sd, prefix="gru."):
    """x [B, T, I] -> [B, 2H] (forward final state | reverse final state)."""
    outs = []
    for suffix, order in (("", range(x.shape[1])), ("_reverse", range(x.shape[1] - 1, -1, -1))):
        w_ih, w_hh = sd[f"{prefix}weight_ih_l0{suffix}"], sd[f"{prefix}weight_hh_l0{suffix}"]
        b_ih, b_hh = sd[f"{prefix}bias_ih_l0{suffix}"], sd[f"{prefix}bias_hh_l0{suffix}"]
        H = w_hh.shape[1]
        h = x.new_zeros(x.shape[0], H)
        for t in order:
            gi = F.linear(x[:, t], w_ih, b_ih)
            gh = F.linear(h, w_hh, b_hh)
            r = torch.sigmoid(gi[:, :H] + gh[:, :H])
            z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
            n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
            h = (1 - z) * n + z * h
        outs.append(h)
    return torch.cat(outs, dim=1)


def heads(h, sd):
    mu = F.linear(h, sd["linear_mu.weight"], sd["linear_mu.bias"])
    var = F.linear(h, sd["linear_var.weight"], sd["linear_var.bias"]).exp()
    return mu, var


def chord_encoder(sd, x):
    """RnnEncoder.forward -> (mean, scale) of the returned Normal."""
    return heads(gru_bidir_last(x, sd), sd)


def texture_encoder(sd, pr):
    bs = pr.shape[0]
    f = F.conv2d(pr.unsqueeze(1), sd["cnn.0.weight"], sd["cnn.0.bias"], stride=(4, 1))
    f = F.max_pool2d(F.relu(f), kernel_size=(1, 4), stride=(1, 4)).reshape(bs, 8, -1)
    f = F.linear(F.linear(f, sd["fc1.weight"], sd["fc1.bias"]), sd["fc2.weight"], sd["fc2.bias"])
    return heads(gru_bidir_last(f, sd), sd)


def encode_chord(sd, chord):
    return chord_encoder(sd, chord)[0].unsqueeze(1)


def encode_txt(sd, prmat):
    z = [texture_encoder(sd, seg)[0] for seg in prmat.split(32, 1)]
    return torch.cat(z, dim=-1).unsqueeze(1)
