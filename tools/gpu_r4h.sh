#!/bin/bash
# round-4 call H: bench A/B of the RAW-segment toggles (whole-step CUDA graph, sustained block)
tag=${1:-r4h}
out=gpurun_out/$tag
mkdir -p $out
ab() {
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    s = d.get("sustained") or {}
    print(f"[{sys.argv[2]}] ms_per_step={d['ms_per_step']:.3f} e2e_ms={d['e2e']['ms_per_step']:.3f} sustained_ms={s.get('ms_per_step'):.3f} breakdown={d['config']['step_breakdown_ms']}")
except Exception as e:
    print(f"[{sys.argv[2]}] failed: {e}")
PY
}
ab off PF_RAW=0
ab gnonly PF_RAW_LN=0 PF_RAW_SKIP=0
ab gnskip PF_RAW_LN=0
ab all PF_X=0
ab off2 PF_RAW=0
