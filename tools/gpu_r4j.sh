#!/bin/bash
# round-4 call J: software-pipelined operand transform: parity + step profiles (77 registers / 3 blocks per SM vs
# 64 registers / 4 blocks per SM)
tag=${1:-r4j}
out=gpurun_out/$tag
mkdir -p $out
timeout 400 python -m pytest tests/test_unet_gpu.py tests/test_ops_gpu.py -x -q > $out/tests.log 2>&1; echo "tests rc=$?"; tail -3 $out/tests.log
prof() { name=$1; shift; env "$@" timeout 300 python tools/profile_step.py 64 > $out/step_$name.txt 2>&1; echo "== $name"; head -5 $out/step_$name.txt; }
prof default PF_X=0
prof mb4 PF_B200_LIB=$PWD/polyffusion_b200/libpf_b200_mb4.so
prof mb4_ipt8 PF_B200_LIB=$PWD/polyffusion_b200/libpf_b200_mb4.so PF_ACT_IPT=8
prof default_ipt8 PF_ACT_IPT=8
