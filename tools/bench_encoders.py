"""Condition-encoder timing: libpf_b200 kernels vs the oracle restatement of the
reference's PyTorch CPU path (all host threads), batch 64 at the sdf_chd8bar / sdf_txt sizes.
Usage (GPU box): python tools/bench_encoders.py > gpurun_out/encoders_bench.txt"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import encoder_oracle as eo
from polyffusion_b200.cond import encode_chord, encode_txt
from polyffusion_b200.dl_modules import RnnEncoder, TextureEncoder

B = 64
torch.manual_seed(0)
ce, te = RnnEncoder(36, 512, 512).eval(), TextureEncoder(256, 1024, 256, 10).eval()
sdc = {k: v.clone() for k, v in ce.state_dict().items()}
sdt = {k: v.clone() for k, v in te.state_dict().items()}
chord = (torch.rand(B, 32, 36) < 0.2).float()
prmat = (torch.rand(B, 128, 128) < 0.03).float()
ce, te = ce.cuda(), te.cuda()
cg, pg = chord.cuda(), prmat.cuda()


def gpu_ms(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cpu_ms(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


torch.set_num_threads(os.cpu_count() or 1)
with torch.no_grad():
    print(f"batch {B}, host threads {torch.get_num_threads()}")
    print(f"chord encoder (bi-GRU 36->512 over 32 steps + heads): GPU {gpu_ms(lambda: encode_chord(ce, cg)):.2f} ms, "
          f"CPU restatement {cpu_ms(lambda: eo.encode_chord(sdc, chord)):.1f} ms")
    print(f"texture encoder (4 x [CNN + fc + bi-GRU 256->1024 over 8 beats + heads]): GPU "
          f"{gpu_ms(lambda: encode_txt(te, pg)):.2f} ms, CPU restatement {cpu_ms(lambda: eo.encode_txt(sdt, prmat)):.1f} ms")
