"""Max / rms error of the CUDA UNet against the committed reference goldens (tests/golden/unet_*.npz).
Usage (GPU box): python tools/unet_err.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _util import build_unet

for name, d_cond in (("unet_chd8bar_b2", 512), ("unet_txtvnl_b1", 128)):
    g = np.load(os.path.join("tests", "golden", name + ".npz"))
    unet = build_unet(d_cond).cuda()
    x, t, c = (torch.from_numpy(g[k]).cuda() for k in ("x", "t", "cond"))
    with torch.no_grad():
        out = unet(x, t, c).cpu().double()
    ref = torch.from_numpy(g["eps"]).double()
    d = (out - ref).abs()
    ok = (d <= 1e-4 + 1e-3 * ref.abs()).double().mean().item()
    print(f"{name}: max abs err {d.max():.3e} rms {d.pow(2).mean().sqrt():.3e} worst tol ratio "
          f"{(d / (1e-4 + 1e-3 * ref.abs())).max():.3f} within tol {ok:.6f}")
