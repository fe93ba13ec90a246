"""CPU: the training path.  ``LatentDiffusion.loss`` over the drop-in
``UNetModel`` in training mode runs the differentiable PyTorch graph (stable_diffusion/model/unet_torch.py)
and must equal the reference's loss and gradients for the same seed (latent_diffusion.py:203-240)."""
import copy
import warnings

import pytest
import torch

from _util import SDF_KW, build_unet
from oracle import reference_loader

warnings.filterwarnings("ignore")


def _batch():
    g = torch.Generator().manual_seed(9)
    return torch.randn(2, 2, 128, 128, generator=g), torch.randn(2, 1, 512, generator=g)


def test_loss_backward_and_optimizer_step():
    from polyffusion_b200.stable_diffusion.latent_diffusion import LatentDiffusion

    unet = build_unet(512).train()
    ldm = LatentDiffusion(unet, None, 0.18215, 1000, 0.00085, 0.012)
    opt = torch.optim.SGD([p for p in ldm.parameters() if p.requires_grad], lr=1e-3)
    x0, cond = _batch()
    torch.manual_seed(4)
    loss = ldm.loss(x0, cond)
    loss.backward()
    assert torch.isfinite(loss) and unet.out[2].weight.grad is not None
    before = unet.out[2].weight.detach().clone()
    opt.step()
    assert not torch.equal(before, unet.out[2].weight.detach())
    # the module (no engine was ever created on the CPU) copies and pickles like any nn.Module
    assert copy.deepcopy(unet)._engine is None


@pytest.mark.skipif(not reference_loader.available(), reason="reference tree not present")
def test_loss_and_gradients_equal_reference():
    from polyffusion_b200.stable_diffusion.latent_diffusion import LatentDiffusion

    ref = reference_loader.load()
    torch.manual_seed(0)
    r_unet = ref.UNetModel(**SDF_KW, d_cond=512).train()
    r_ldm = ref.LatentDiffusion(r_unet, None, 0.18215, 1000, 0.00085, 0.012)
    unet = build_unet(512).train()
    ldm = LatentDiffusion(unet, None, 0.18215, 1000, 0.00085, 0.012)
    x0, cond = _batch()
    torch.manual_seed(4)
    want = r_ldm.loss(x0, cond)
    want.backward()
    torch.manual_seed(4)
    got = ldm.loss(x0, cond)
    got.backward()
    assert abs(got.item() - want.item()) < 1e-6
    for name in ("out.2.weight", "input_blocks.0.0.weight", "middle_block.1.transformer_blocks.0.attn2.to_v.weight"):
        a = dict(r_unet.named_parameters())[name].grad
        b = dict(unet.named_parameters())[name].grad
        assert (a - b).abs().max().item() <= 1e-6 * max(1.0, a.abs().max().item()), name
