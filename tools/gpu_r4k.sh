#!/bin/bash
# round-4 call K: 256-wide tiles for the small-M linears (A/B), single-pass fast mode (error + bench)
tag=${1:-r4k}
out=gpurun_out/$tag
mkdir -p $out
timeout 400 python -m pytest tests/test_unet_gpu.py -x -q > $out/tests.log 2>&1; echo "tests rc=$?"; tail -3 $out/tests.log
prof() { name=$1; shift; env "$@" timeout 300 python tools/profile_step.py 64 > $out/step_$name.txt 2>&1; echo "== $name"; head -4 $out/step_$name.txt; grep -m2 "gemm M=16384 N=256 K=256" $out/step_$name.txt; }
prof bn256small PF_X=0
prof bn128small PF_BN256_SMALL=0
prof fast PF_FAST=1
env PF_FAST=1 timeout 300 python tools/unet_err.py > $out/unet_err_fast.txt 2>&1; tail -2 $out/unet_err_fast.txt
timeout 400 python bench.py --fast --steps 20 --warmup 3 --no-cpu > $out/bench_fast.json 2> $out/bench_fast.err; python -c "
import json; d=json.load(open('$out/bench_fast.json')); print('fast', d['ms_per_step'], d['value'], d['sustained']['ms_per_step'], d['config']['step_breakdown_ms'])"
