"""CPU: host-side logic of the drop-in classes that needs no GPU -- schedule tables, the integer
DDIM timestep path (bit-exact), CFG dispatch, error behaviour, module aliasing, sharding math."""
import numpy as np
import pytest
import torch

from _util import SDF_KW, build_unet


def small_ldm():
    from polyffusion_b200.stable_diffusion.latent_diffusion import LatentDiffusion

    return LatentDiffusion(build_unet(512), None, 0.18215, 1000, 0.00085, 0.012)


def test_schedule_tables_match_oracle():
    from oracle import sampler_oracle as so
    from polyffusion_b200.sampler_sdf import SDFSampler

    ldm = small_ldm()
    alpha, beta, alpha_bar = so.ldm_schedule()
    assert torch.equal(ldm.alpha.data, alpha) and torch.equal(ldm.beta.data, beta)
    assert torch.equal(ldm.alpha_bar.data, alpha_bar)
    s = SDFSampler(ldm)
    tb = so.ddpm_tables(alpha_bar, beta)
    assert torch.equal(s.log_var, tb["log_var"]) and torch.equal(s.mean_x0_coef, tb["mean_x0"])
    assert s.time_steps.dtype == np.int32 and s.time_steps[0] == 0 and s.time_steps[-1] == 999
    # host copies used on the launch path are the same fp32 values
    assert s._h["c4"][0] == float((0.5 * tb["log_var"][0]).exp())
    assert abs(float(tb["log_var"][0]) - (-46.0517)) < 1e-3  # clamp(var, 1e-20) at t = 0


@pytest.mark.parametrize("n,disc", [(50, "uniform"), (10, "uniform"), (20, "quad"), (7, "uniform"), (333, "quad")])
def test_ddim_integer_path_bit_exact(n, disc):
    from oracle import sampler_oracle as so
    from polyffusion_b200.sampler_ddim import DDIMSampler

    d = DDIMSampler(small_ldm(), n, disc, 0.0)
    tau = so.ddim_time_steps(1000, n, disc)
    assert np.array_equal(d.time_steps, tau)
    if disc == "uniform" and n == 50:
        assert list(tau[:3]) == [1, 21, 41] and tau[-1] == 981


def test_ddim_1000_steps_indexes_out_of_range_like_reference():
    """tau max = 1000 indexes a 1000-entry table -> IndexError (inherited)."""
    from polyffusion_b200.sampler_ddim import DDIMSampler

    with pytest.raises(IndexError):
        DDIMSampler(small_ldm(), 1000, "uniform", 0.0)


def test_cfg_dispatch_matches_reference_semantics():
    from polyffusion_b200.stable_diffusion.sampler import DiffusionSampler

    calls = []

    class Fake:
        n_steps = 10

        def __call__(self, x, t, c):
            calls.append((x.shape[0], float(c.flatten()[0])))
            return x * 0 + c.flatten()[0]

    s = DiffusionSampler(Fake())
    x, t = torch.zeros(2, 2, 4, 4), torch.zeros(2, dtype=torch.long)
    c, u = torch.full((2, 1, 3), 2.0), torch.full((2, 1, 3), -1.0)
    e = s.get_eps(x, t, c, uncond_scale=1.0, uncond_cond=u)
    assert calls[-1] == (2, 2.0) and float(e[0, 0, 0, 0]) == 2.0
    e = s.get_eps(x, t, c, uncond_scale=0.0, uncond_cond=u)
    assert calls[-1] == (2, -1.0)
    e = s.get_eps(x, t, c, uncond_scale=5.0, uncond_cond=None)
    assert calls[-1] == (2, 2.0)
    e_c, e_u = s._eps_pair(x, t, c, uncond_scale=5.0, uncond_cond=u)
    assert calls[-1][0] == 4  # doubled batch, uncond first (sampler/__init__.py:72-74)
    assert e_u is not None


def test_cpu_tensors_fail_loudly():
    from polyffusion_b200._lib import PfError

    m = build_unet(512)
    with pytest.raises(PfError):
        with torch.no_grad():
            m(torch.zeros(1, 2, 128, 128), torch.zeros(1, dtype=torch.long), torch.zeros(1, 1, 512))
    from polyffusion_b200.sampler_sdf import SDFSampler

    s = SDFSampler(small_ldm())
    with pytest.raises(RuntimeError):
        s.q_sample(torch.zeros(1, 2, 8, 8), 3, torch.zeros(1, 2, 8, 8))


def test_forward_dispatch_between_autograd_graph_and_cuda_plan():
    """Gradients required -> differentiable PyTorch graph (tests/test_training_path.py checks its values);
    no_grad / eval on CPU tensors -> the CUDA plan, which refuses CPU tensors (no CPU fallback)."""
    from polyffusion_b200._lib import PfError

    m = build_unet(512)
    x, t, c = torch.zeros(1, 2, 128, 128), torch.zeros(1, dtype=torch.long), torch.zeros(1, 1, 512)
    assert m.train()._wants_autograd(x, c) and not m.eval()._wants_autograd(x, c)
    assert m._wants_autograd(x.clone().requires_grad_(True), c)
    with torch.no_grad():
        assert not m.train()._wants_autograd(x, c)
        with pytest.raises(PfError):
            m.eval()(x, t, c)


def test_install_dropin_aliases_reference_import_paths():
    import importlib
    import sys

    import polyffusion_b200

    saved = {k: sys.modules.get(k) for k in polyffusion_b200._DROPIN_MODULES}
    try:
        polyffusion_b200.install_dropin()
        unet_mod = importlib.import_module("stable_diffusion.model.unet")
        assert unet_mod.UNetModel is polyffusion_b200.stable_diffusion.model.unet.UNetModel
        assert importlib.import_module("sampler_sdf").SDFSampler.__module__ == "polyffusion_b200.sampler_sdf"
        assert hasattr(importlib.import_module("ddpm"), "DenoiseDiffusion")
        assert importlib.import_module("dl_modules.chord_enc").RnnEncoder.__module__ == "polyffusion_b200.dl_modules.chord_enc"
        assert importlib.import_module("dl_modules.txt_enc").TextureEncoder.__module__ == "polyffusion_b200.dl_modules.txt_enc"
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_legacy_ddpm_tables():
    from oracle import sampler_oracle as so
    from polyffusion_b200.ddpm import DenoiseDiffusion

    d = DenoiseDiffusion(torch.nn.Identity(), 1000)
    alpha, beta, alpha_bar = so.legacy_schedule(1000)
    assert torch.equal(d.beta, beta) and torch.equal(d.alpha_bar, alpha_bar) and d.sigma2 is d.beta
    assert "beta" in dict(d.named_buffers()) and "alpha" not in dict(d.named_buffers())


def test_shard_bounds_cover_batch():
    from polyffusion_b200.parallel import sample_seed, shard_bounds

    for total, world in ((512, 8), (64, 1), (10, 4), (3, 8), (0, 2)):
        spans = [shard_bounds(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)
    assert sample_seed(1, 5) != sample_seed(1, 6)
