"""Drop-in for the reference's ``dl_modules/txt_enc.py:5-35`` (``TextureEncoder``): CNN over a
32-step piano roll, two linears, bidirectional GRU over the 8 beats, ``linear_mu`` / ``linear_var``.
The forward runs in libpf_b200 (pf_txt_cnn, pf_linear, pf_gru_bidir_last)."""
import torch
from torch import nn
from torch.distributions import Normal

from . import _ops


class TextureEncoder(nn.Module):
    def __init__(self, emb_size, hidden_dim, z_dim, num_channel=10):
        """input must be piano_mat: (B, 32, 128)"""
        super().__init__()
        self.cnn = nn.Sequential(
            nn.Conv2d(1, num_channel, kernel_size=(4, 12), stride=(4, 1), padding=0),
            nn.ReLU(),
            nn.MaxPool2d(kernel_size=(1, 4), stride=(1, 4)),
        )
        self.fc1 = nn.Linear(num_channel * 29, 1000)
        self.fc2 = nn.Linear(1000, emb_size)
        self.gru = nn.GRU(emb_size, hidden_dim, batch_first=True, bidirectional=True)
        self.linear_mu = nn.Linear(hidden_dim * 2, z_dim)
        self.linear_var = nn.Linear(hidden_dim * 2, z_dim)
        self.emb_size = emb_size
        self.hidden_dim = hidden_dim
        self.z_dim = z_dim

    @torch.no_grad()
    def forward(self, pr):
        _ops.require_cuda(pr, "dl_modules.TextureEncoder")
        bs = pr.size(0)
        # [B, C, 8, 29] viewed as (bs, 8, -1): the reference's .view mixes channel and beat (txt_enc.py:26)
        f = _ops.txt_cnn(pr, self.cnn[0]).view(bs, 8, -1)
        f = _ops.linear(_ops.linear(f, self.fc1), self.fc2)   # (bs, 8, emb_size), no activation between
        h = _ops.gru_bidir_last(f, self.gru)
        mu = _ops.linear(h, self.linear_mu)
        var = _ops.linear(h, self.linear_var, act=2)
        return Normal(mu, var)
