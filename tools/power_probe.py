"""SM clock / board power while the UNet evaluation runs back to back (what bounds the big GEMMs).
Usage (GPU box): python tools/power_probe.py > gpurun_out/power_probe.txt"""
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import sdf_kwargs
from polyffusion_b200.stable_diffusion.model.unet import UNetModel

samples = []
stop = False


def pump():
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,"
                          "clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,temperature.gpu",
                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        samples.append((time.perf_counter(), line.strip()))
        if stop:
            break
    p.terminate()


torch.manual_seed(0)
m = UNetModel(**sdf_kwargs()).eval().cuda()
x = torch.randn(64, 2, 128, 128, device="cuda")
c = torch.randn(64, 1, 512, device="cuda")
t = torch.randint(0, 1000, (64,), device="cuda")
with torch.no_grad():
    for _ in range(3):
        m(x, t, c)
    torch.cuda.synchronize()
    th = threading.Thread(target=pump, daemon=True)
    th.start()
    time.sleep(1.0)
    t_idle = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 300
    for _ in range(n):
        m(x, t, c)
    e1.record()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    time.sleep(0.5)
    stop = True
print(f"{n} UNet evaluations at batch 64 back to back: {e0.elapsed_time(e1) / n:.2f} ms each")
print("t_rel_s, sm_mhz, sm_max_mhz, power_w, power_limit_w, sw_power_cap, hw_slowdown, temp_c")
for ts, line in samples:
    tag = "idle" if ts < t_idle else ("load" if ts < t_end else "after")
    print(f"{ts - t_idle:7.2f} {tag:5s} {line}")
