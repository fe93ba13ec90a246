#!/bin/bash
# A/B of environment switches on the bench: tools/gpu_ab.sh <tag> "<ENV=..>" "<ENV=..>" ...
tag=$1; shift
out=gpurun_out/$tag
mkdir -p $out
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 400 python bench.py --steps 10 --warmup 3 > $out/bench_$i.json 2> $out/bench_$i.err
  python - "$out/bench_$i.json" "$v" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"[{sys.argv[2]}] ms_per_step={d['ms_per_step']:.3f} value={d['value']:.3f} e2e={d['e2e']['value']:.3f} breakdown={d['config'].get('step_breakdown_ms')} clocks={d['clocks']}")
except Exception as e:
    print(f"[{sys.argv[2]}] failed: {e}")
PY
done
