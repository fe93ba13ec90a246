#!/bin/bash
# One gpurun call: parity tests, A/B step profiles, bench.  Usage: tools/gpu_round.sh <tag>
tag=${1:-x}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_unet_gpu.py -x -q > $out/unet_default.log 2>&1
rc=$?
echo "unet_default rc=$rc" | tee -a $out/summary.txt
tail -5 $out/unet_default.log
if [ $rc -ne 0 ]; then
  for v in "PF_GEMM_XF=0" "PF_GEMM_HALO=0" "PF_GEMM_STACK=0" "PF_CONV_SLOW=1" "PF_GEMM_STACK=0 PF_CONV_SLOW=1" "PF_GEMM_STACK=64" "PF_GEMM_STACK=128" "PF_GEMM_STACK=0 PF_CONV_SLOW=1 PF_ATTN_UNFUSED=1"; do
    env $v timeout 600 python -m pytest tests/test_unet_gpu.py -x -q > "$out/unet_$(echo $v | tr ' =' '__').log" 2>&1
    echo "unet [$v] rc=$?" | tee -a $out/summary.txt
  done
fi
timeout 1500 python -m pytest tests -m gpu -x -q > $out/gpu_tests.log 2>&1
echo "gpu_tests rc=$?" | tee -a $out/summary.txt
tail -3 $out/gpu_tests.log
timeout 300 python tools/profile_step.py 64 > $out/step_profile_b64.txt 2>&1
head -14 $out/step_profile_b64.txt
PF_GEMM_XF=0 timeout 300 python tools/profile_step.py 64 > $out/step_profile_b64_noxf.txt 2>&1
head -5 $out/step_profile_b64_noxf.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$?" | tee -a $out/summary.txt
cat $out/bench_n1.json
