#!/bin/bash
# round-4 call P: attention without the row-maximum pass (norm-bound stabiliser): parity, A/B
tag=${1:-r4r}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_ops_gpu.py tests/test_samplers_gpu.py -x -q > $out/tests.log 2>&1; echo "tests rc=$?"; tail -3 $out/tests.log
timeout 300 python tools/unet_err.py > $out/unet_err_1pass.txt 2>&1; tail -2 $out/unet_err_1pass.txt
env PF_ATTN_1PASS=0 timeout 300 python tools/unet_err.py > $out/unet_err_2pass.txt 2>&1; tail -2 $out/unet_err_2pass.txt
prof() { name=$1; shift; env "$@" timeout 300 python tools/profile_step.py 64 > $out/step_$name.txt 2>&1; echo "== $name"; head -5 $out/step_$name.txt | tail -4; grep -m2 "us  attn" $out/step_$name.txt; }
prof 1pass PF_X=0
prof 2pass PF_ATTN_1PASS=0
prof 1pass_b PF_X=0
prof 2pass_b PF_ATTN_1PASS=0
