#!/bin/bash
# round-2 final evidence: every GPU test, then tools/gpu_final.sh (smoke, bench both arms, ncu launch list + DRAM
# bytes of one step, step profile), the other BASELINE configs, the non-parity fast mode
tag=${1:-r4z}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $out/gpu_tests.log
bash tools/gpu_final.sh $tag
for c in 2 3 4; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --sustain-seconds 0 > $out/bench_config$c.json 2> $out/bench_config$c.err
  python -c "
import json; d=json.load(open('$out/bench_config$c.json')); print('config $c', d['metric'], d['value'], d['unit'], d['ms_per_step'])"
done
timeout 400 python bench.py --fast --steps 20 --warmup 3 --no-cpu > $out/bench_fast.json 2> $out/bench_fast.err
python -c "
import json; d=json.load(open('$out/bench_fast.json')); print('fast', d['ms_per_step'], d['value'], d['sustained']['ms_per_step'])"
