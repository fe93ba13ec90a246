#!/bin/bash
# round-3 call D: whole-step graph / fused epilogue / Philox + new parity tests + bench configs
tag=${1:-r3d}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -15 $out/gpu_tests.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $out/smoke.log
ab() {
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu $BARGS > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    s = d.get("sustained") or {}
    print(f"[{sys.argv[2]}] value={d['value']:.4f} {d['unit']} ms_per_step={d['ms_per_step']:.3f} e2e_ms={d['e2e']['ms_per_step']:.3f} sustained_ms={s.get('ms_per_step')} launches={d['gpu_launches']} clocks={d['clocks']}")
except Exception as e:
    print(f"[{sys.argv[2]}] failed: {e}")
PY
}
BARGS=""
ab fused_philox
ab fused_torchnoise PF_NOISE=torch
ab unfused PF_FUSED_LOOP=0 PF_NOISE=torch
BARGS="--config 2 --sustain-seconds 0"; ab config2
BARGS="--config 3 --sustain-seconds 0"; ab config3
BARGS="--config 4 --sustain-seconds 0"; ab config4
