"""In-kernel noise (Philox4x32-10) and the whole-step graph loop.

CPU: the numpy oracle against the Random123 known-answer vectors of Philox4x32-10 (Salmon et al., SC'11;
kat_vectors of the Random123 distribution).  GPU: ``pf_fill_normal`` against the oracle; a Philox-noise
``paint`` run sharded two ways against the unsharded run (rank-count invariance); the fused loop against
the per-step launch path with the same injected noise.
"""
import ctypes

import numpy as np
import pytest
import torch

from _util import NoiseTape, build_unet
from oracle import philox_oracle as po


def test_philox4x32_10_known_answers():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = tuple(int(v) for v in po.philox4x32_10(*ctr, *key))
        assert got == want, (ctr, key, [hex(v) for v in got])


def test_oracle_normals_are_standard_normal_and_sample_keyed():
    a = po.normal(99, 0, 4, 32768, 5, 0)
    assert abs(float(a.mean())) < 0.01 and abs(float(a.std()) - 1.0) < 0.01
    b = po.normal(99, 2, 2, 32768, 5, 0)
    assert np.array_equal(a[2:], b)  # samples 2, 3 are the same numbers wherever the shard starts
    assert not np.array_equal(po.normal(99, 0, 1, 64, 5, 1), a[:1, :64])  # known-region stream differs


@pytest.mark.gpu
def test_fill_normal_matches_oracle():
    from polyffusion_b200._lib import check, current_stream, lib, ptr

    for seed, sample0, index, which in ((1234, 0, 999, 0), (2**63 + 5, 37, 3, 1)):
        out = torch.empty(3, 32768, device="cuda")
        check(lib().pf_fill_normal(ptr(out), 3, 32768, ctypes.c_uint64(seed), sample0, index, which, current_stream()))
        want = po.normal(seed, sample0, 3, 32768, index, which)
        err = np.abs(out.cpu().numpy() - want).max()
        assert err < 2e-5, err  # same integers; fp32 log / cos differ by a few ulp between libm and CUDA


def _sampler(noise="torch"):
    from polyffusion_b200.sampler_sdf import SDFSampler
    from polyffusion_b200.stable_diffusion.latent_diffusion import LatentDiffusion

    ldm = LatentDiffusion(build_unet(512), None, 0.18215, 1000, 0.00085, 0.012).cuda()
    s = SDFSampler(ldm)
    s.noise = noise
    return s


@pytest.mark.gpu
def test_philox_paint_is_invariant_to_sharding():
    """4 samples in one batch == the same samples generated as two shards of 2 (sample0 = 0, 2)."""
    s = _sampler("philox")
    s.seed = 4242
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 2, 128, 128, generator=g).cuda()
    cond = torch.randn(4, 1, 512, generator=g).cuda()
    orig = (torch.rand(4, 2, 128, 128, generator=g) < 0.02).float().cuda()
    mask = torch.zeros(4, 2, 128, 128)
    mask[:, :, :, 64:] = 1.0
    mask = mask.cuda()
    s.sample0 = 0
    full = s.paint(x, cond, 3, orig=orig, mask=mask)
    nonce = s._loop.nonce()
    parts = []
    for lo in (0, 2):
        s.sample0 = lo
        s._loop._nonce = nonce - 1  # same run nonce as the unsharded run (ranks run in lockstep)
        parts.append(s.paint(x[lo:lo + 2], cond[lo:lo + 2], 3, orig=orig[lo:lo + 2], mask=mask[lo:lo + 2]))
    err = (torch.cat(parts) - full).abs().max().item()
    print("sharded vs unsharded Philox paint: max abs diff", err)
    assert err < 1e-4  # GroupNorm statistics use atomics: last-bit differences only
    # and the noise really entered: a different seed gives a different sample
    s.sample0, s.seed = 0, 4243
    assert (s.paint(x, cond, 3, orig=orig, mask=mask) - full).abs().max().item() > 1e-2


@pytest.mark.gpu
def test_fused_loop_equals_per_step_path():
    """Same injected noise through the whole-step graph and through the per-step launches."""
    import test_samplers_gpu as tsg

    s = _sampler("torch")
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 2, 128, 128, generator=g).cuda()
    cond = torch.randn(2, 1, 512, generator=g).cuda()
    orig = (torch.rand(2, 2, 128, 128, generator=g) < 0.02).float().cuda()
    mask = (torch.rand(2, 2, 128, 128, generator=g) < 0.5).float().cuda()
    outs = []
    for fused in (True, False, True):
        s.fused_loop = fused
        with tsg.CudaTape(123):
            outs.append(s.paint(x, cond, 4, orig=orig, mask=mask))
    assert (outs[0] - outs[1]).abs().max().item() < 1e-4
    assert (outs[0] - outs[2]).abs().max().item() < 1e-4  # second use of the captured graph
    # plain generation through sample(), incl. the last step (no noise at step 0)
    outs = []
    for fused in (True, False):
        s.fused_loop = fused
        with tsg.CudaTape(124):
            outs.append(s.sample([2, 2, 128, 128], cond, x_last=x, t_start=996, temperature=0.8))
    assert (outs[0] - outs[1]).abs().max().item() < 1e-4
