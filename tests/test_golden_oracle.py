"""CPU: the oracle restatement (oracle/) against the golden vectors generated from the REAL
reference (oracle/make_golden.py).  The oracle runs the same ATen CPU ops as the reference, so the
match is expected to be exact up to thread-count reorder noise (<= 2e-6)."""
import os

import numpy as np
import pytest
import torch

from _util import NoiseTape, SDF_KW, build_unet, oracle_cfg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(os.path.join(GOLD, name)).items()}


@pytest.fixture(scope="module")
def sd512():
    return build_unet(512).state_dict()


def eps_fn_for(sd, d_cond):
    from oracle.unet_oracle import unet_forward

    return lambda x, t, c: unet_forward(sd, oracle_cfg(d_cond), x, t, c)


def test_unet_golden_chd8bar(sd512):
    g = load("unet_chd8bar_b2.npz")
    out = eps_fn_for(sd512, 512)(g["x"], g["t"], g["cond"])
    assert (out - g["eps"]).abs().max().item() < 5e-6


def test_unet_golden_txtvnl():
    g = load("unet_txtvnl_b1.npz")
    sd = build_unet(128).state_dict()
    out = eps_fn_for(sd, 128)(g["x"], g["t"], g["cond"])
    assert (out - g["eps"]).abs().max().item() < 5e-6


def test_tables_golden():
    from oracle import sampler_oracle as so

    g = load("tables.npz")
    alpha, beta, alpha_bar = so.ldm_schedule()
    assert torch.equal(alpha, g["alpha"]) and torch.equal(beta, g["beta"]) and torch.equal(alpha_bar, g["alpha_bar"])
    tb = so.ddpm_tables(alpha_bar, beta)
    for key, gk in (("sqrt_ab", "sdf_sqrt_ab"), ("sqrt_1m_ab", "sdf_sqrt_1m_ab"), ("sqrt_recip_ab", "sdf_sqrt_recip_ab"),
                    ("sqrt_recip_m1_ab", "sdf_sqrt_recip_m1_ab"), ("log_var", "sdf_log_var"),
                    ("mean_x0", "sdf_mean_x0"), ("mean_xt", "sdf_mean_xt")):
        assert torch.equal(tb[key], g[gk]), key
    # integer timestep path: bit exact
    tau = so.ddim_time_steps(1000, 50, "uniform")
    assert np.array_equal(tau, g["d50_tau"].numpy())
    tauq = so.ddim_time_steps(1000, 20, "quad")
    assert np.array_equal(tauq, g["dq_tau"].numpy())
    d = so.ddim_tables(alpha_bar, tau, 0.0)
    assert torch.equal(d["alpha"], g["d50_alpha"]) and torch.equal(d["alpha_prev"], g["d50_alpha_prev"])
    assert torch.equal(d["sigma"], g["d50_sigma"]) and torch.equal(d["sqrt_1m_alpha"], g["d50_sqrt_1m_alpha"])
    dq = so.ddim_tables(alpha_bar, tauq, 0.5)
    assert torch.equal(dq["sigma"], g["dq_sigma"]) and torch.equal(dq["alpha_prev"], g["dq_alpha_prev"])


def test_paint_ddpm_cfg_golden(sd512):
    from oracle import sampler_oracle as so

    g = load("paint_ddpm_cfg5.npz")
    _, beta, alpha_bar = so.ldm_schedule()
    tape = NoiseTape(int(g["tape_seed"]))
    tb = so.ddpm_tables(alpha_bar, beta)
    xt = tb["sqrt_ab"][2] * g["orig"] + tb["sqrt_1m_ab"][2] * tape(g["orig"].shape)
    assert (xt - g["x_t"]).abs().max().item() < 1e-6
    out = so.ddpm_paint(alpha_bar, beta, eps_fn_for(sd512, 512), xt, g["cond"], 2, tape, orig=g["orig"],
                        mask=g["mask"], uncond_scale=5.0, uncond_cond=g["uncond"])
    assert (out - g["out"]).abs().max().item() < 2e-5


def test_paint_ddpm_repaint2_and_sample_golden(sd512):
    from oracle import sampler_oracle as so

    _, beta, alpha_bar = so.ldm_schedule()
    g = load("paint_ddpm_repaint2.npz")
    out = so.ddpm_paint(alpha_bar, beta, eps_fn_for(sd512, 512), g["x_start"], g["cond"], 1,
                        NoiseTape(int(g["tape_seed"])), orig=g["orig"], mask=g["mask"], repaint_n=2)
    assert (out - g["out"]).abs().max().item() < 2e-5
    g = load("sample_ddpm.npz")
    out = so.ddpm_sample(alpha_bar, beta, eps_fn_for(sd512, 512), g["x_start"], g["cond"],
                         NoiseTape(int(g["tape_seed"])), t_start=int(g["t_start"]))
    assert (out - g["out"]).abs().max().item() < 2e-5


def test_ddim_golden(sd512):
    from oracle import sampler_oracle as so

    g = load("ddim.npz")
    _, _, alpha_bar = so.ldm_schedule()
    out = so.ddim_run(alpha_bar, eps_fn_for(sd512, 512), g["x_start"], g["cond"], NoiseTape(int(g["sample_tape_seed"])),
                      n_steps=4, t_start=int(g["sample_t_start"]))
    assert (out - g["sample_out"]).abs().max().item() < 2e-5
    out = so.ddim_run(alpha_bar, eps_fn_for(sd512, 512), g["x_start"], g["cond"], NoiseTape(int(g["paint_tape_seed"])),
                      n_steps=4, eta=float(g["paint_eta"]), t_start=int(g["paint_t_start"]), paint=True,
                      orig=g["orig"], mask=g["mask"], orig_noise=g["orig_noise"])
    assert (out - g["paint_out"]).abs().max().item() < 2e-5


def test_legacy_ddpm_golden():
    from oracle import sampler_oracle as so

    g = load("legacy_ddpm.npz")
    alpha, beta, alpha_bar = so.legacy_schedule(1000)
    assert torch.equal(beta, g["beta"]) and torch.equal(alpha_bar, g["alpha_bar"])
    torch.manual_seed(7)
    conv = torch.nn.Conv2d(2, 2, 3, padding=1)
    eps_fn = lambda x, t: (conv(x) + (t.float() / 1000.0)[:, None, None, None]).detach()
    tape = NoiseTape(int(g["tape_seed"]))
    x = g["x_T"]
    for ti in range(999, 989, -1):
        x = so.legacy_p_sample(alpha, beta, alpha_bar, eps_fn, x, x.new_full((4,), ti, dtype=torch.long), tape)
    assert (x - g["out"]).abs().max().item() < 1e-5
