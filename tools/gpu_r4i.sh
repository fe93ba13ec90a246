#!/bin/bash
# round-4 call I: all GPU tests on the new defaults (GroupNorm -> proj_in RAW, one-round-trip transform prologue),
# the opt-in variants, step profiles for the transform's items-per-thread
tag=${1:-r4i}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -5 $out/gpu_tests.log
prof() { name=$1; shift; env "$@" timeout 300 python tools/profile_step.py 64 > $out/step_$name.txt 2>&1; echo "== $name"; head -5 $out/step_$name.txt; }
prof ipt4 PF_X=0
prof ipt2 PF_ACT_IPT=2
prof ipt1 PF_ACT_IPT=1
