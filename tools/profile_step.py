"""Per-launch breakdown of one UNet evaluation (CUDA events around every kernel of the plan).
Usage (on the GPU box): python tools/profile_step.py [B] > gpurun_out/step_profile.txt"""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import sdf_kwargs
from polyffusion_b200.stable_diffusion.model.unet import UNetModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_cond = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.manual_seed(0)
kw = sdf_kwargs()
if n_cond > 1:
    kw["d_cond"] = 128
m = UNetModel(**kw).eval().cuda()
x = torch.randn(B, 2, 128, 128, device="cuda")
c = torch.randn(B, n_cond, kw["d_cond"], device="cuda")
t = torch.randint(0, 1000, (B,), device="cuda")
with torch.no_grad():
    for _ in range(3):
        m(x, t, c)
    acc = None
    R = 5
    for _ in range(R):
        prof = {}
        m.engine.forward(x, t, c, profile=prof)
        acc = prof["ms"] if acc is None else [a + b for a, b in zip(acc, prof["ms"])]
ms = [a / R for a in acc]
desc = m.engine.op_descriptions()
tot = sum(ms)
print(f"B={B} total {tot:.3f} ms over {len(ms)} launches")
bykind = defaultdict(lambda: [0.0, 0])
for d, t_ in zip(desc, ms):
    k = d.split()[0]
    bykind[k][0] += t_
    bykind[k][1] += 1
for k, (t_, n) in sorted(bykind.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:14s} {t_:8.3f} ms  {n:4d} launches  {100*t_/tot:5.1f}%")
print()
gemm_shapes = defaultdict(lambda: [0.0, 0, 0.0])
for d, t_, fl in zip(desc, ms, prof["flops"]):
    if d.startswith("gemm"):
        gemm_shapes[d][0] += t_
        gemm_shapes[d][1] += 1
        gemm_shapes[d][2] += fl
print("GEMM shapes (aggregated):  ms  count  TFLOP/s(algorithmic)")
for d, (t_, n, fl) in sorted(gemm_shapes.items(), key=lambda kv: -kv[1][0]):
    print(f"  {t_:8.3f} {n:3d} {fl/t_/1e9:8.1f}  {d}")
print()
print("all launches in order:")
for i, (d, t_) in enumerate(zip(desc, ms)):
    print(f"{i:4d} {t_*1000:9.1f} us  {d}")
