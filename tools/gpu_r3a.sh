#!/bin/bash
# round-3 call A: parity with half-batch lanes, A/B of the lane / co-residency switches, sanitizer
tag=${1:-r3a}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q > $out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 $out/gpu_tests.log
ab() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"[{sys.argv[2]}] ms_per_step={d['ms_per_step']:.3f} e2e_ms={d['e2e']['ms_per_step']:.3f} breakdown={d['config'].get('step_breakdown_ms')} clocks={d['clocks']}")
except Exception as e:
    print(f"[{sys.argv[2]}] failed: {e}")
PY
}
ab base_lb320_nolanes PF_B200_LIB=$PWD/polyffusion_b200/libpf_b200_lb320.so PF_LANE_MIN_HW=0 PF_ACT_THREADS=256
ab cap128_nolanes PF_LANE_MIN_HW=0 PF_ACT_THREADS=256
ab lanes_default
ab lanes_128only PF_LANE_MIN_HW=16384
ab lanes_ge32 PF_LANE_MIN_HW=1024
ab lanes_act256 PF_ACT_THREADS=256
ab lanes_lb320 PF_B200_LIB=$PWD/polyffusion_b200/libpf_b200_lb320.so
ab lanes_serial PF_LANES_SERIAL=1
# sanitizer on a B=2 evaluation (lanes active) + one sampler step
cat > $out/san.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from bench import sdf_kwargs
from polyffusion_b200.stable_diffusion.model.unet import UNetModel
import os
os.environ["PF_CUDA_GRAPH"] = "0"
torch.manual_seed(0)
m = UNetModel(**sdf_kwargs()).eval().cuda()
x = torch.randn(2, 2, 128, 128, device="cuda"); c = torch.randn(2, 1, 512, device="cuda")
t = torch.randint(0, 1000, (2,), device="cuda")
with torch.no_grad():
    y = m(x, t, c)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
PY
PF_CUDA_GRAPH=0 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python $out/san.py > $out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $out/sanitizer_memcheck.log
PF_CUDA_GRAPH=0 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python $out/san.py > $out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 $out/sanitizer_racecheck.log
