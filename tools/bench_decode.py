"""Piano-roll decode timing: pf_prmat2c_to_prmat + pf_prmat_notes on the GPU
vs the reference's Python loops (oracle plain-loop restatement, bounded sample) on the host.
Usage (GPU box): python tools/bench_decode.py [n_segments] > gpurun_out/decode_bench.txt"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import decode_oracle as do
from oracle.make_golden import synthetic_prmat2c
from polyffusion_b200.utils import prmat2c_durations, prmat2c_to_notes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2560  # BASELINE config 5: 256 songs x 10 segments
base = synthetic_prmat2c(64, 128, 7)
x = torch.from_numpy(np.concatenate([base] * (n // 64))).cuda()
for _ in range(3):
    prmat2c_durations(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
R = 20
for _ in range(R):
    d = prmat2c_durations(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / R
bytes_alg = x.shape[0] * (2 * 128 * 128 * 4 + 128 * 128 * 8)
t0 = time.perf_counter()
nt = prmat2c_to_notes(x)
t_notes = time.perf_counter() - t0
t0 = time.perf_counter()
do.prmat2c_to_prmat(base[:2])
cpu_per_seg = (time.perf_counter() - t0) / 2
print(f"segments {x.shape[0]}: duration matrix {ms:.3f} ms on the GPU = {bytes_alg / ms / 1e6:.0f} GB/s algorithmic "
      f"(onset + sustain fp32 read once, int64 durations written once)")
print(f"notes (count + scan + ordered write + D2H of {len(nt)} notes): {t_notes * 1e3:.1f} ms wall")
print(f"reference Python loops (oracle restatement, 2 segments, 1 host thread): {cpu_per_seg * 1e3:.1f} ms per segment "
      f"-> {cpu_per_seg * x.shape[0]:.1f} s for {x.shape[0]} segments")
