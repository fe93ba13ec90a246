#!/bin/bash
tag=${1:-r4e}
out=gpurun_out/$tag
mkdir -p $out
for d in 0 1 3 7 4; do
  env PF_RAW_DBG=$d timeout 300 python tools/profile_step.py 64 > $out/step_dbg$d.txt 2>&1
  echo "DBG=$d"; head -3 $out/step_dbg$d.txt | tail -2; grep -m3 "us  gemm M=65536 N=512 K=256\|us  gemm M=65536 N=2048\|us  gemm M=1048576 N=64 K=704" $out/step_dbg$d.txt
done
