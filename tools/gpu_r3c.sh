#!/bin/bash
# round-3 call C: default configuration (f16f8 convolutions below 128x128) + legacy UNet + A/B
tag=${1:-r3c}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -12 $out/gpu_tests.log
timeout 200 python tools/unet_err.py > $out/unet_err_default.log 2>&1; tail -2 $out/unet_err_default.log
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
