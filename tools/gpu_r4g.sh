#!/bin/bash
# round-4 call G: f16f8 feed-forward GEMMs + per-group act_split prologue: parity, error report, step profiles
tag=${1:-r4g}
out=gpurun_out/$tag
mkdir -p $out
env PF_RAW=0 timeout 240 python -m pytest tests/test_unet_gpu.py tests/test_ops_gpu.py -x -q > $out/tests_noraw.log 2>&1; echo "tests_noraw rc=$?"; tail -3 $out/tests_noraw.log
timeout 240 python -m pytest tests/test_unet_gpu.py -x -q > $out/tests_raw.log 2>&1; echo "tests_raw rc=$?"; tail -3 $out/tests_raw.log
env PF_RAW=0 timeout 300 python tools/unet_err.py > $out/unet_err_noraw.txt 2>&1; tail -2 $out/unet_err_noraw.txt
env PF_RAW=0 PF_FF_F8=0 timeout 300 python tools/unet_err.py > $out/unet_err_noraw_noff8.txt 2>&1; tail -2 $out/unet_err_noraw_noff8.txt
prof() { name=$1; shift; env "$@" timeout 300 python tools/profile_step.py 64 > $out/step_$name.txt 2>&1; echo "== $name"; head -6 $out/step_$name.txt; }
prof noraw_noff8 PF_RAW=0 PF_FF_F8=0
prof noraw PF_RAW=0
prof force PF_RAW=0 PF_FF_F8=0 PF_RAW_FORCE=1
prof gnonly PF_RAW_LN=0 PF_RAW_SKIP=0
prof all PF_X=0
