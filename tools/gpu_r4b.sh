#!/bin/bash
# round-4 call B: RAW segments after the proxy-fence fix: UNet parity, per-launch step profile
tag=${1:-r4b}
out=gpurun_out/$tag
mkdir -p $out
timeout 240 python -m pytest tests/test_unet_gpu.py -x -q > $out/unet_all.log 2>&1; rc=$?
echo "unet_all rc=$rc"; tail -4 $out/unet_all.log
timeout 300 python tools/unet_err.py > $out/unet_err_raw.txt 2>&1; tail -3 $out/unet_err_raw.txt
timeout 300 python tools/profile_step.py 64 > $out/step_profile_b64_raw.txt 2>&1; head -12 $out/step_profile_b64_raw.txt
for t in "PF_RAW_SKIP=0" "PF_RAW_GN=0" "PF_RAW_LN=0"; do
  env $t timeout 300 python tools/profile_step.py 64 > $out/step_profile_b64_$t.txt 2>&1; echo $t; head -5 $out/step_profile_b64_$t.txt
done
