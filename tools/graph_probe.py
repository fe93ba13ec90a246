"""Experiment: does replaying one UNet evaluation as a CUDA graph shorten it (launch gaps)?"""
import os, sys
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
out = torch.empty_like(x)
with torch.no_grad():
    for _ in range(3):
        m.engine.forward(x, t, c, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        m.engine.forward(x, t, c, out=out)
    e1.record(); torch.cuda.synchronize()
    print("eager  ms/eval", e0.elapsed_time(e1) / 20)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        m.engine.forward(x, t, c, out=out)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        m.engine.forward(x, t, c, out=out)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print("graph  ms/eval", e0.elapsed_time(e1) / 20)
