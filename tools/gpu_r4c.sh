#!/bin/bash
# round-4 call C: ncu source-level capture of one RAW GEMM launch (conversion-warp stalls)
tag=${1:-r4c}
out=gpurun_out/$tag
mkdir -p $out
PF_CUDA_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2s_raw -s 0 -c 2 -o $out/raw_qk python tools/ncu_target.py 64 > $out/ncu.log 2>&1
echo "ncu rc=$?"; tail -5 $out/ncu.log; ls -la $out
