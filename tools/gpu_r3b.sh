#!/bin/bash
# round-3 call B: f16f8 convolution operands -- parity first, then A/B against split-bf16
tag=${1:-r3b}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q > $out/ops.log 2>&1; echo "ops rc=$?"; tail -5 $out/ops.log
for hw in 4096 0 16384; do
  PF_CONV_F8_MAX_HW=$hw timeout 200 python tools/unet_err.py > $out/unet_err_$hw.log 2>&1; echo "unet_err hw=$hw rc=$?"; tail -3 $out/unet_err_$hw.log
done
timeout 400 python -m pytest tests -m gpu -x -q > $out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -6 $out/gpu_tests.log
PF_CONV_F8_MAX_HW=16384 timeout 400 python -m pytest tests -m gpu -q > $out/gpu_tests_f8all.log 2>&1; echo "gpu tests (f8 everywhere) rc=$?"; tail -6 $out/gpu_tests_f8all.log
ab() {
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
ab f8_le64
ab bf16x3 PF_CONV_F8_MAX_HW=0
ab f8_all PF_CONV_F8_MAX_HW=16384
ab f8_le64_again
timeout 300 python tools/profile_step.py 64 > $out/step_profile_b64.txt 2>&1; head -14 $out/step_profile_b64.txt
PF_CONV_F8_MAX_HW=16384 timeout 300 python tools/profile_step.py 64 > $out/step_profile_b64_f8all.txt 2>&1; head -6 $out/step_profile_b64_f8all.txt
