#!/bin/bash
tag=${1:-r4u}
out=gpurun_out/$tag
mkdir -p $out
prof() { name=$1; shift; env "$@" timeout 300 python tools/profile_step.py 64 > $out/step_$name.txt 2>&1; echo "== $name"; head -3 $out/step_$name.txt | tail -2; grep "us  act_split" $out/step_$name.txt | awk '{k=$5" "$6" "$7; s[k]+=$2; n[k]++} END {for (k in s) printf "%8.1f %3d %6.1f  %s\n", s[k], n[k], s[k]/n[k], k}' | sort -k4 | grep "16x16\|32x32"; }
prof ipt4 PF_X=0
prof ipt2 PF_ACT_SMALL_IPT=2
prof ipt1 PF_ACT_SMALL_IPT=1
