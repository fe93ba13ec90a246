#!/bin/bash
# round-4 call O: persistent (grid-stride) operand transform: blocks = w resident waves
tag=${1:-r4o}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -m pytest tests/test_unet_gpu.py -x -q > $out/tests.log 2>&1; echo "tests rc=$?"; tail -2 $out/tests.log
prof() { name=$1; shift; env "$@" timeout 300 python tools/profile_step.py 64 > $out/step_$name.txt 2>&1; echo "== $name"; head -4 $out/step_$name.txt | tail -3; }
prof w0 PF_ACT_WAVES=0
prof w1 PF_ACT_WAVES=1
prof w2 PF_ACT_WAVES=2
prof w4 PF_ACT_WAVES=4
prof w0b PF_ACT_WAVES=0
