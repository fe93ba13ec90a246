"""Small target for `ncu --set full`: two UNet evaluations at batch 64 (sdf_chd8bar geometry)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import sdf_kwargs
from polyffusion_b200.stable_diffusion.model.unet import UNetModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
m = UNetModel(**sdf_kwargs()).eval().cuda()
x = torch.randn(B, 2, 128, 128, device="cuda")
c = torch.randn(B, 1, 512, device="cuda")
t = torch.randint(0, 1000, (B,), device="cuda")
with torch.no_grad():
    for _ in range(2):
        m(x, t, c)
torch.cuda.synchronize()
print("done")
