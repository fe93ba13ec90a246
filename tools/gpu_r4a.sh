#!/bin/bash
# round-4 call A: first light of the RAW GEMM segments (in-kernel operand conversion): UNet parity per
# feature toggle, then the per-launch step profile with and without them
tag=${1:-r4a}
out=gpurun_out/$tag
mkdir -p $out
run() {
  name=$1; shift
  env "$@" timeout 240 python -m pytest tests/test_unet_gpu.py -x -q > $out/unet_$name.log 2>&1
  rc=$?
  echo "unet_$name rc=$rc"; tail -4 $out/unet_$name.log
  return $rc
}
run off PF_RAW=0
run skip PF_RAW_GN=0 PF_RAW_LN=0
run gn PF_RAW_SKIP=0 PF_RAW_LN=0
run ln PF_RAW_SKIP=0 PF_RAW_GN=0
run all
rc=$?
env PF_RAW=0 timeout 300 python tools/unet_err.py > $out/unet_err_off.txt 2>&1; tail -3 $out/unet_err_off.txt
timeout 300 python tools/unet_err.py > $out/unet_err_raw.txt 2>&1; tail -3 $out/unet_err_raw.txt
env PF_RAW=0 timeout 300 python tools/profile_step.py 64 > $out/step_profile_b64_off.txt 2>&1; head -12 $out/step_profile_b64_off.txt
if [ $rc -eq 0 ]; then
  timeout 300 python tools/profile_step.py 64 > $out/step_profile_b64_raw.txt 2>&1; head -12 $out/step_profile_b64_raw.txt
fi
