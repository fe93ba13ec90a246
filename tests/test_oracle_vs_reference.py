"""CPU, build container only (skipped where /root/reference is absent): pin the oracle restatement
and the drop-in module surface directly against the imported, unmodified reference."""
import warnings

import pytest
import torch

from _util import NoiseTape, SDF_KW, build_unet, oracle_cfg
from oracle import reference_loader

pytestmark = pytest.mark.skipif(not reference_loader.available(), reason="reference tree not present")
warnings.filterwarnings("ignore")


@pytest.fixture(scope="module")
def ref():
    return reference_loader.load()


@pytest.fixture(scope="module")
def ref_unet(ref):
    torch.manual_seed(0)
    return ref.UNetModel(**SDF_KW, d_cond=512).eval()


def test_dropin_state_dict_and_seeded_init_match(ref_unet):
    mine = build_unet(512)
    sa, sb = ref_unet.state_dict(), mine.state_dict()
    assert list(sa.keys()) == list(sb.keys()) and len(sa) == 556
    assert all(sa[k].shape == sb[k].shape and torch.equal(sa[k], sb[k]) for k in sa)
    mine.load_state_dict(sa)  # reference checkpoints load unchanged


def test_oracle_unet_equals_reference(ref_unet):
    from oracle.unet_oracle import unet_forward

    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 2, 128, 128, generator=g)
    c = torch.randn(1, 1, 512, generator=g)
    t = torch.tensor([640])
    with torch.no_grad():
        want = ref_unet(x, t, c)
    got = unet_forward(ref_unet.state_dict(), oracle_cfg(512), x, t, c)
    assert (got - want).abs().max().item() < 5e-6


def test_dropin_tables_equal_reference(ref, ref_unet):
    from polyffusion_b200.sampler_ddim import DDIMSampler
    from polyffusion_b200.sampler_sdf import SDFSampler
    from polyffusion_b200.stable_diffusion.latent_diffusion import LatentDiffusion

    rl = ref.LatentDiffusion(ref_unet, None, 0.18215, 1000, 0.00085, 0.012)
    ml = LatentDiffusion(build_unet(512), None, 0.18215, 1000, 0.00085, 0.012)
    for k in ("alpha", "beta", "alpha_bar"):
        assert torch.equal(getattr(rl, k).data, getattr(ml, k).data)
    assert ml.sigma2 is ml.beta
    rs, ms = ref.SDFSampler(rl), SDFSampler(ml)
    assert (rs.time_steps == ms.time_steps).all() and rs.time_steps.dtype == ms.time_steps.dtype
    for k in ("sqrt_alpha_bar", "sqrt_1m_alpha_bar", "sqrt_recip_alpha_bar", "sqrt_recip_m1_alpha_bar",
              "log_var", "mean_x0_coef", "mean_xt_coef"):
        assert torch.equal(getattr(rs, k), getattr(ms, k)), k
    for disc, n, eta in (("uniform", 50, 0.0), ("quad", 20, 0.5), ("uniform", 10, 1.0)):
        rd, md = ref.DDIMSampler(rl, n, disc, eta), DDIMSampler(ml, n, disc, eta)
        assert (rd.time_steps == md.time_steps).all()
        for k in ("ddim_alpha", "ddim_alpha_sqrt", "ddim_alpha_prev", "ddim_sigma", "ddim_sqrt_one_minus_alpha"):
            assert torch.equal(getattr(rd, k), getattr(md, k)), (disc, k)
    with pytest.raises(NotImplementedError):
        DDIMSampler(ml, 10, "cosine")
    with pytest.raises(NotImplementedError):
        ref.DDIMSampler(rl, 10, "cosine")


def test_oracle_paint_equals_reference(ref, ref_unet):
    """Two RePaint steps through the real SDFSampler.paint vs the oracle, same injected noise."""
    from oracle import sampler_oracle as so
    from oracle.make_golden import Tape
    from oracle.unet_oracle import unet_forward

    ldm = ref.LatentDiffusion(ref_unet, None, 0.18215, 1000, 0.00085, 0.012)
    sdf = ref.SDFSampler(ldm)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 2, 128, 128, generator=g)
    cond = torch.randn(1, 1, 512, generator=g)
    orig = (torch.rand(1, 2, 128, 128, generator=g) < 0.05).float()
    mask = (torch.rand(1, 2, 128, 128, generator=g) < 0.5).float()
    with Tape(77), torch.no_grad():
        want = sdf.paint(x, cond, 1, orig=orig, mask=mask)
    sd = ref_unet.state_dict()
    eps_fn = lambda xx, tt, cc: unet_forward(sd, oracle_cfg(512), xx, tt, cc)
    got = so.ddpm_paint(ldm.alpha_bar.data, ldm.beta.data, eps_fn, x, cond, 1, NoiseTape(77), orig=orig, mask=mask)
    assert (got - want).abs().max().item() < 1e-5
