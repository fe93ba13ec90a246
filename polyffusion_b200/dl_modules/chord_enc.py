"""Drop-in for the reference's ``dl_modules/chord_enc.py:5-22`` (``RnnEncoder``, imported as
``ChordEncoder``): bidirectional GRU over the chord sequence, ``linear_mu`` / ``linear_var`` heads,
returns ``Normal(mu, exp(linear_var))``.  The forward runs in libpf_b200 (pf_gru_bidir_last,
pf_linear); the submodules exist for the parameter tree, ``state_dict`` keys and init RNG stream."""
import torch
from torch import nn
from torch.distributions import Normal

from . import _ops


class RnnEncoder(nn.Module):
    def __init__(self, input_dim, hidden_dim, z_dim):
        super().__init__()
        self.gru = nn.GRU(input_dim, hidden_dim, batch_first=True, bidirectional=True)
        self.linear_mu = nn.Linear(hidden_dim * 2, z_dim)
        self.linear_var = nn.Linear(hidden_dim * 2, z_dim)
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.z_dim = z_dim

    @torch.no_grad()
    def forward(self, x):
        _ops.require_cuda(x, "dl_modules.RnnEncoder")
        h = _ops.gru_bidir_last(x, self.gru)           # [B, 2H], forward half first (chord_enc.py:15-17)
        mu = _ops.linear(h, self.linear_mu)
        var = _ops.linear(h, self.linear_var, act=2)   # .exp_() (chord_enc.py:20)
        return Normal(mu, var)
