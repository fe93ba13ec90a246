#!/bin/bash
# first-light check of a risky kernel change: UNet parity with a short timeout, then the fallback
tag=${1:-q}
out=gpurun_out/$tag
mkdir -p $out
timeout 150 python -m pytest tests/test_unet_gpu.py -x -q > $out/unet_default.log 2>&1
rc=$?
echo "unet_default rc=$rc"
tail -15 $out/unet_default.log
if [ $rc -ne 0 ]; then
  PF_GEMM_XF=0 timeout 150 python -m pytest tests/test_unet_gpu.py -x -q > $out/unet_noxf.log 2>&1
  echo "unet_noxf rc=$?"
  tail -3 $out/unet_noxf.log
else
  timeout 200 python tools/profile_step.py 64 > $out/step_profile_b64.txt 2>&1
  head -14 $out/step_profile_b64.txt
  grep "N=64 " $out/step_profile_b64.txt | head -8
fi
