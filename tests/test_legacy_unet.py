"""Legacy unconditional eps-model ``ddpm.unet.UNet`` (BASELINE config 1:
``UNet(2, 64, [1, 2, 2, 4], [F, F, F, T])`` under ``DenoiseDiffusion.p_sample``, B = 4, 10 steps).

Golden: tests/golden/legacy_unet.npz, produced by the REAL reference on the CPU
(``python -m oracle.make_golden legacy_unet``): one evaluation at t = [999, 500, 3, 250] and the state
after 10 reverse steps (999..990) with taped noise; weights are the seeded default initialisation,
which the drop-in reproduces from the same seed.
"""
import os

import numpy as np
import pytest
import torch

from _util import ATOL, RTOL, NoiseTape, close_report
from oracle import reference_loader

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CFG = (2, 64, [1, 2, 2, 4], [False, False, False, True])


def _golden():
    return {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(os.path.join(GOLD, "legacy_unet.npz")).items()}


def _build():
    from polyffusion_b200.ddpm.unet import UNet

    torch.manual_seed(0)
    return UNet(*CFG).eval()


@pytest.mark.skipif(not reference_loader.available(), reason="reference tree not present")
def test_state_dict_and_seeded_init_match_reference():
    ref = reference_loader.load()
    torch.manual_seed(0)
    theirs = ref.UNet(*CFG).state_dict()
    mine = _build().state_dict()
    assert list(theirs.keys()) == list(mine.keys()) and len(mine) == 308
    assert all(torch.equal(theirs[k], mine[k]) for k in mine)
    assert sum(v.numel() for v in mine.values()) == 167_776_834


def test_autograd_graph_matches_golden_and_differentiates():
    """The PyTorch graph used for training (train/train_ddpm.py) reproduces the reference's output."""
    g = _golden()
    m = _build().train()
    x, t = g["x"][:1].clone().requires_grad_(True), g["t"][:1]
    y = m(x, t)
    assert y.requires_grad
    assert (y.detach() - g["eps"][:1]).abs().max().item() < 2e-5  # fp32 reorder noise (thread count)
    y.square().mean().backward()
    assert x.grad is not None and m.final.weight.grad is not None


def test_cpu_no_grad_raises():
    m = _build()
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 2, 128, 128), torch.zeros(1, dtype=torch.long))


@pytest.mark.gpu
def test_cuda_forward_vs_reference_golden():
    g = _golden()
    m = _build().cuda()
    with torch.no_grad():
        out = m(g["x"].cuda(), g["t"].cuda())
    err, frac = close_report(out, g["eps"])
    print(f"legacy UNet eps: max abs err {err:.3e}, within tol {frac:.6f}")
    assert frac == 1.0, f"max abs err {err}, fraction within rtol {RTOL}/atol {ATOL} = {frac}"


@pytest.mark.gpu
def test_config1_ten_reverse_steps_vs_reference_golden():
    """BASELINE configs[0]: DenoiseDiffusion.p_sample plumbing around the real legacy UNet."""
    from polyffusion_b200.ddpm import DenoiseDiffusion

    g = _golden()
    dd = DenoiseDiffusion(_build(), 1000).cuda()
    tape = NoiseTape(int(g["tape_seed"]))
    randn = torch.randn
    torch.randn = lambda *size, **kw: tape(size[0] if len(size) == 1 and not isinstance(size[0], int) else size).cuda()
    try:
        x = g["x"].cuda()
        for ti in range(999, 989, -1):
            x = dd.p_sample(x, x.new_full((4,), ti, dtype=torch.long))
    finally:
        torch.randn = randn
    err, frac = close_report(x, g["out"])
    print(f"legacy DDPM 10 steps: max abs err {err:.3e}, within tol {frac:.6f}")
    assert frac == 1.0, f"max abs err {err}, fraction within rtol {RTOL}/atol {ATOL} = {frac}"
