#!/bin/bash
# round-3 call E: all GPU tests (incl. Philox / fused loop), BN=256 f16f8 tiles A/B, step profile
tag=${1:-r3e}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -15 $out/gpu_tests.log
ab() {
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu $BARGS > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    s = d.get("sustained") or {}
    print(f"[{sys.argv[2]}] value={d['value']:.4f} {d['unit']} ms_per_step={d['ms_per_step']:.3f} e2e_ms={d['e2e']['ms_per_step']:.3f} sustained_ms={s.get('ms_per_step')} breakdown={d['config']['step_breakdown_ms']}")
except Exception as e:
    print(f"[{sys.argv[2]}] failed: {e}")
PY
}
BARGS=""
ab bn256
ab bn128 PF_F8_BN256=0
ab bn256_again
timeout 300 python tools/profile_step.py 64 > $out/step_profile_b64.txt 2>&1; head -60 $out/step_profile_b64.txt | tail -48
