#!/bin/bash
# round-4 call N: fused q|k|v launch (OUT_QKV): parity, A/B step profile
tag=${1:-r4n}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_variants_gpu.py -x -q > $out/tests.log 2>&1; echo "tests rc=$?"; tail -3 $out/tests.log
prof() { name=$1; shift; env "$@" timeout 300 python tools/profile_step.py 64 > $out/step_$name.txt 2>&1; echo "== $name"; head -5 $out/step_$name.txt; grep "N=768\|N=512 K=256\|K=256 bn=128 taps=1 nseg=1 z=1 mode=2" $out/step_$name.txt | head -4; }
prof fused PF_X=0
prof split PF_QKV_FUSED=0
prof fused2 PF_X=0
prof split2 PF_QKV_FUSED=0
