#!/bin/bash
# round-4 call S: two softmax groups (one per S/P buffer) vs sixteen warps in lockstep, same box
tag=${1:-r4s}
out=gpurun_out/$tag
mkdir -p $out
prof() { name=$1; shift; env "$@" timeout 300 python tools/profile_step.py 64 > $out/step_$name.txt 2>&1; echo "== $name"; head -4 $out/step_$name.txt | tail -3; grep -m1 "us  attn B=64 heads=4 N=1024" $out/step_$name.txt; grep -m1 "us  attn B=64 heads=4 N=256" $out/step_$name.txt; }
prof grp PF_X=0
prof nogrp PF_B200_LIB=$PWD/polyffusion_b200/libpf_b200_nogrp.so
prof grp2 PF_X=0
prof nogrp2 PF_B200_LIB=$PWD/polyffusion_b200/libpf_b200_nogrp.so
