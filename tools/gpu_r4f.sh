#!/bin/bash
tag=${1:-r4f}
out=gpurun_out/$tag
mkdir -p $out
env PF_RAW=0 timeout 300 python tools/profile_step.py 64 > $out/step_off.txt 2>&1
env PF_RAW=0 PF_RAW_FORCE=1 timeout 300 python tools/profile_step.py 64 > $out/step_force.txt 2>&1
for f in off force; do echo $f; head -3 $out/step_$f.txt | tail -2; grep -m6 "us  gemm M=65536 N=512 K=256\|us  gemm M=65536 N=2048\|us  gemm M=1048576 N=64 K=704\|us  gemm M=65536 N=256 K=2816\|us  gemm M=65536 N=256 K=256 bn=128 taps=1 nseg=1 z=1 mode=0" $out/step_$f.txt; done
