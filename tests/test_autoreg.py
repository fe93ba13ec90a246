"""Autoregressive driver pieces: get_mask / get_autoreg_data.
CPU: oracle vs golden outputs of the reference functions (+ direct pin where the reference exists);
GPU: pf_get_mask (bit-exact, it is integer/binary work) and the song-batched autoregressive paint."""
import os

import numpy as np
import pytest
import torch

from _util import NoiseTape, build_unet, close_report, oracle_cfg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mask_autoreg.npz")


def cases():
    from oracle.make_golden import synthetic_melody

    g = np.load(GOLD)
    for i in range(4):
        n_seg, seed, blank, p0 = (int(v) for v in g[f"case{i}_args"])
        o = synthetic_melody(n_seg, seed, blank, bool(p0))
        want = {k: torch.from_numpy(np.unpackbits(g[f"case{i}_{k}"])[: o.numel()].reshape(o.shape).astype(np.float32))
                for k in ("below", "above")}
        yield o, want


def test_oracle_get_mask_matches_reference_golden():
    from oracle import autoreg_oracle as ao

    for o, want in cases():
        for k in ("below", "above"):
            assert torch.equal(ao.get_mask(o, k), want[k])
    g = np.load(GOLD)
    assert np.array_equal(ao.get_autoreg_data(torch.from_numpy(g["autoreg_in"]), 2).numpy(), g["autoreg_out"])


def test_get_mask_without_onsets_raises_like_reference():
    from oracle import autoreg_oracle as ao

    with pytest.raises(IndexError):
        ao.get_mask(torch.zeros(1, 2, 128, 128), "below")


def test_get_autoreg_data_per_song_matches_reference_per_song():
    from oracle import autoreg_oracle as ao
    from polyffusion_b200.autoreg import get_autoreg_data

    g = torch.Generator().manual_seed(3)
    d = torch.randn(6, 2, 8, 4, generator=g)  # 2 songs x 3 segments
    got = get_autoreg_data(d, 2, seg_per_song=3)
    want = torch.cat([ao.get_autoreg_data(d[:3], 2), ao.get_autoreg_data(d[3:], 2)])
    assert torch.equal(got, want)
    assert torch.equal(get_autoreg_data(d, 2), ao.get_autoreg_data(d, 2))


@pytest.mark.gpu
def test_gpu_get_mask_bit_exact():
    from oracle import autoreg_oracle as ao
    from polyffusion_b200.autoreg import get_mask

    for o, want in cases():
        for k in ("below", "above"):
            got = get_mask(o.cuda(), k).cpu()
            assert torch.equal(got, want[k]), k
    # song-batched: every song is scanned on its own
    from oracle.make_golden import synthetic_melody

    o = torch.cat([synthetic_melody(3, 11), synthetic_melody(3, 12, blank_first=9)])
    got = get_mask(o.cuda(), "below", seg_per_song=3).cpu()
    want = torch.cat([ao.get_mask(o[:3], "below"), ao.get_mask(o[3:], "below")])
    assert torch.equal(got, want)
    assert torch.equal(get_mask(o.cuda(), "remaining").cpu(), o)
    bars = get_mask(o.cuda(), "bars", bar_list=[1, 6]).cpu()
    assert bars[:, :, 16:32].sum() == 0 and bars[:, :, 0:16].min() == 1
    from polyffusion_b200._lib import PfError

    with pytest.raises(PfError):
        get_mask(torch.zeros(1, 2, 128, 128).cuda(), "below")


@pytest.mark.gpu
def test_song_batched_autoreg_matches_sequential_oracle():
    """2 songs x 2 segments, inpaint 'below', t_idx = 0 (one deterministic reverse step per window):
    the song-batched GPU driver equals the reference's sequential batch-1 procedure on the oracle."""
    from oracle import autoreg_oracle as ao
    from oracle import sampler_oracle as so
    from oracle.make_golden import synthetic_melody
    from oracle.unet_oracle import unet_forward
    from polyffusion_b200.autoreg import autoreg_paint, get_mask
    from polyffusion_b200.sampler_sdf import SDFSampler
    from polyffusion_b200.stable_diffusion.latent_diffusion import LatentDiffusion

    unet = build_unet(512)
    sd = {k: v.clone() for k, v in unet.state_dict().items()}
    ldm = LatentDiffusion(unet, None, 0.18215, 1000, 0.00085, 0.012).cuda()
    sampler = SDFSampler(ldm)
    S, B = 2, 2
    g = torch.Generator().manual_seed(5)
    orig = torch.cat([synthetic_melody(B, 21), synthetic_melody(B, 22)])
    cond = torch.randn(S * B, 1, 512, generator=g)
    cond_mid = torch.randn(S * B, 1, 512, generator=g)
    noise = torch.randn(S * B, 2, 128, 128, generator=g)
    mask = get_mask(orig.cuda(), "below", seg_per_song=B)
    got = autoreg_paint(sampler, cond.cuda(), cond_mid.cuda(), 0, seg_per_song=B, orig=orig.cuda(),
                        mask=mask, noise=noise.cuda()).cpu()
    assert got.shape == (S, 2 * B, 2, 64, 128)

    _, beta, alpha_bar = so.ldm_schedule()
    tb = so.ddpm_tables(alpha_bar, beta)
    eps_fn = lambda x, t, c: unet_forward(sd, oracle_cfg(512), x, t, c)
    paint = lambda xt, c, t_idx, o, m: so.ddpm_paint(alpha_bar, beta, eps_fn, xt, c, t_idx, NoiseTape(0), orig=o, mask=m)
    qs = lambda o, t_idx, n: tb["sqrt_ab"][t_idx] * o + tb["sqrt_1m_ab"][t_idx] * n
    for s in range(S):
        sl = slice(s * B, (s + 1) * B)
        m_cpu = ao.get_mask(orig[sl], "below")
        want = ao.predict_autoreg(paint, qs, cond[sl], cond_mid[sl], orig[sl], m_cpu, noise[sl], 0)
        err, frac = close_report(got[s], want)
        print(f"song {s}: max abs err {err:.3e}")
        assert frac == 1.0, f"song {s}: max abs err {err}"


PREDICT_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "predict_autoreg.npz")


def _predict_golden():
    g = np.load(PREDICT_GOLD)
    t = {k: torch.from_numpy(g[k]) for k in ("orig", "mask", "cond", "cond_mid")}
    cases = [(tag, float(g[f"{tag}_scale"]), int(g[f"{tag}_tape_seed"]), torch.from_numpy(g[f"{tag}_out"]))
             for tag in ("plain", "cfg")]
    return t, int(g["t_idx"]), cases


def test_oracle_predict_matches_the_reference_predict_golden():
    """tests/golden/predict_autoreg.npz holds the output of the reference's OWN ``Experiments.predict(autoreg=True)``
    (inference_sdf.py:202-283, lifted by ast and run on the reference's SDFSampler / UNetModel on the CPU by
    oracle/make_golden.py).  The oracle restatement of the driver must reproduce it with the same noise tape."""
    from oracle import autoreg_oracle as ao
    from oracle import sampler_oracle as so
    from oracle.unet_oracle import unet_forward

    t, t_idx, cases = _predict_golden()
    sd = build_unet(512).state_dict()
    _, beta, alpha_bar = so.ldm_schedule()
    tb = so.ddpm_tables(alpha_bar, beta)
    eps_fn = lambda x, tt, c: unet_forward(sd, oracle_cfg(512), x, tt, c)
    qs = lambda o, ti, n: tb["sqrt_ab"][ti] * o + tb["sqrt_1m_ab"][ti] * n
    uncond = -torch.ones(1, 1, 512)
    for tag, scale, seed, want in cases:
        tape = NoiseTape(seed)
        noise = tape(tuple(t["orig"].shape))  # predict draws the q_sample noise first (inference_sdf.py:225)
        paint = lambda xt, c, ti, o, m: so.ddpm_paint(alpha_bar, beta, eps_fn, xt, c, ti, tape, orig=o, mask=m,
                                                      uncond_scale=scale, uncond_cond=uncond)
        got = ao.predict_autoreg(paint, qs, t["cond"], t["cond_mid"], t["orig"], t["mask"], noise, t_idx)
        err = (got - want).abs().max().item()
        print(f"{tag}: oracle vs reference predict max abs err {err:.3e}")
        assert got.shape == want.shape and err < 2e-5, (tag, err)


@pytest.mark.gpu
def test_gpu_autoreg_paint_matches_the_reference_predict_golden():
    """The CUDA drop-in driven the way the reference drives it -- one song, sequential half-overlapping windows,
    torch.randn in the reference's order -- against the reference's own ``Experiments.predict`` output."""
    from _util import CudaTape
    from polyffusion_b200.autoreg import autoreg_paint
    from polyffusion_b200.sampler_sdf import SDFSampler
    from polyffusion_b200.stable_diffusion.latent_diffusion import LatentDiffusion

    t, t_idx, cases = _predict_golden()
    ldm = LatentDiffusion(build_unet(512), None, 0.18215, 1000, 0.00085, 0.012).cuda()
    sampler = SDFSampler(ldm)
    uncond = -torch.ones(2, 1, 512).cuda()
    for tag, scale, seed, want in cases:
        with CudaTape(seed):
            got = autoreg_paint(sampler, t["cond"].cuda(), t["cond_mid"].cuda(), t_idx, seg_per_song=2,
                                orig=t["orig"].cuda(), mask=t["mask"].cuda(), uncond_scale=scale,
                                uncond_cond=uncond)
        err, frac = close_report(got[0], want)
        print(f"{tag}: CUDA autoreg_paint vs reference predict max abs err {err:.3e}, within tol {frac:.6f}")
        assert got.shape[1:] == want.shape and frac == 1.0, (tag, err, frac)
