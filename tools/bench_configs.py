"""Timing of BASELINE.json configs[2..4] on ONE GPU through the public sampler API (the bench line
itself is configs[1], bench.py).  Every reverse step costs the same irrespective of t, so the DDPM
configs are timed over a bounded number of steps (t_start = STEPS - 1) and scaled to 1000.
Usage (GPU box): python tools/bench_configs.py > gpurun_out/configs_bench.txt"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import sdf_kwargs
from polyffusion_b200.autoreg import autoreg_paint, get_mask
from polyffusion_b200.sampler_ddim import DDIMSampler
from polyffusion_b200.sampler_sdf import SDFSampler
from polyffusion_b200.stable_diffusion.latent_diffusion import LatentDiffusion
from polyffusion_b200.stable_diffusion.model.unet import UNetModel

dev = torch.device("cuda:0")
STEPS = 20  # DDPM reverse steps actually timed


def ldm(d_cond):
    torch.manual_seed(0)
    kw = sdf_kwargs()
    kw["d_cond"] = d_cond
    return LatentDiffusion(UNetModel(**kw).eval(), None, 0.18215, 1000, 0.00085, 0.012).to(dev)


def timed(fn, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


with torch.no_grad():
    # configs[2]: sdf_txt conditional, batch 64, DDIM 50 steps eta 0
    m = ldm(1024)
    ddim = DDIMSampler(m, 50, "uniform", 0.0)
    cond = torch.randn(64, 1, 1024, device=dev)
    sec, x = timed(lambda: ddim.sample([64, 2, 128, 128], cond))
    print(f"configs[2] sdf_txt DDIM-50 eta=0, batch 64: {sec * 1e3:.0f} ms per batch = {64 / sec:.1f} samples/s "
          f"({sec / 50 * 1e3:.2f} ms per step), finite={bool(torch.isfinite(x).all())}")

    # configs[3]: sdf_chd8bar, classifier-free guidance scale 5 (2 UNet evaluations per step), 64 per GPU
    m = ldm(512)
    sdf = SDFSampler(m)
    cond = torch.randn(64, 1, 512, device=dev)
    uncond = -torch.ones(64, 1, 512, device=dev)
    orig = torch.zeros(64, 2, 128, 128, device=dev)
    mask = torch.zeros_like(orig)
    xt = sdf.q_sample(orig, STEPS - 1, torch.randn_like(orig))
    sec, x = timed(lambda: sdf.paint(xt, cond, STEPS - 1, orig=orig, mask=mask, uncond_scale=5.0, uncond_cond=uncond))
    per = sec / STEPS
    print(f"configs[3] sdf_chd8bar CFG scale 5, batch 64 per GPU (UNet batch 128): {per * 1e3:.2f} ms per step "
          f"-> {64 / (1000 * per):.2f} samples/s per GPU for 1000-step DDPM, finite={bool(torch.isfinite(x).all())}")

    # configs[4]: autoregressive inpainting "below", 10 segments per song, 32 songs per GPU (256 over 8 GPUs)
    songs, segs = 32, 10
    g = torch.Generator(device="cpu").manual_seed(3)
    o = torch.zeros(songs * segs, 2, 128, 128)
    on = (torch.rand(songs * segs, 128, generator=g) < 0.25)
    pitch = torch.randint(60, 85, (songs * segs, 128), generator=g)
    idx = on.nonzero()
    o[idx[:, 0], 0, idx[:, 1], pitch[idx[:, 0], idx[:, 1]]] = 1.0
    o[::segs, 0, 0, 72] = 1.0  # every song has at least one onset
    o = o.to(dev)
    sec_mask, msk = timed(lambda: get_mask(o, "below", seg_per_song=segs))
    cond = torch.randn(songs * segs, 1, 512, device=dev)
    cond_mid = torch.randn(songs * segs, 1, 512, device=dev)
    sec, gen = timed(lambda: autoreg_paint(sdf, cond, cond_mid, STEPS - 1, seg_per_song=segs, orig=o, mask=msk), warm=0)
    per_paint_step = sec / ((2 * segs - 1) * STEPS)
    full = per_paint_step * (2 * segs - 1) * 1000
    print(f"configs[4] autoregressive inpaint 'below', {songs} songs x {segs} segments per GPU: get_mask {sec_mask * 1e3:.2f} ms; "
          f"{per_paint_step * 1e3:.2f} ms per (window, step) at batch {songs} -> {full:.0f} s per {songs} songs for 19 windows x 1000 steps "
          f"= {songs / full:.3f} songs/s per GPU; output {tuple(gen.shape)}, finite={bool(torch.isfinite(gen).all())}")
