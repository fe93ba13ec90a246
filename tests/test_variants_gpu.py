"""Opt-in kernel variants of the UNet plan (environment switches read once per process, so each case runs
`tools/unet_err.py` in a fresh interpreter): every variant must reproduce the reference goldens
(tests/golden/unet_*.npz, generated from the real reference by oracle/make_golden.py) within
rtol 1e-3 / atol 1e-4.

  PF_RAW=0        no RAW GEMM segments at all (GroupNorm -> proj_in through the operand transform)
  PF_RAW_SKIP=1   ResBlock 1x1 skip conv reads the fp32 inputs through the conversion warps (f16f8 RAW kernels)
  PF_RAW_LN=1     LayerNorm -> q/k/v / GeGLU through the conversion warps, row statistics from GEMM epilogues
  PF_FF_F8=1      feed-forward GEMMs on f16f8 operands (GeGLU epilogue writing an f16f8 operand)
  PF_CONV_F8_MAX_HW=0  split-bf16 convolutions everywhere
  PF_QKV_FUSED=0  separate q|k and V projection launches instead of the fused OUT_QKV launch
  PF_ATTN_1PASS=0 attention always computes its row maxima in a first pass (no norm-bound stabiliser)
  PF_RAW_GN_BN=128 GroupNorm -> proj_in RAW segment on 128-wide stacked tiles (default: 256-wide)
"""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    {"PF_RAW": "0"},
    {"PF_RAW_SKIP": "1"},
    {"PF_RAW_LN": "1"},
    {"PF_RAW_SKIP": "1", "PF_RAW_LN": "1"},
    {"PF_FF_F8": "1"},
    {"PF_CONV_F8_MAX_HW": "0", "PF_RAW_LN": "1"},
    {"PF_QKV_FUSED": "0"},
    {"PF_ATTN_1PASS": "0"},
    {"PF_RAW_GN_BN": "128"},
]


@pytest.mark.parametrize("env", CASES, ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_variant_matches_reference_goldens(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "unet_err.py")], cwd=ROOT, env=e,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = re.findall(r"(unet_\w+): max abs err (\S+) rms (\S+) worst tol ratio (\S+) within tol (\S+)", r.stdout)
    assert len(rows) == 2, r.stdout
    for name, mx, rms, ratio, frac in rows:
        print(f"{env} {name}: max abs err {mx}, worst tolerance ratio {ratio}")
        assert float(frac) == 1.0 and float(ratio) < 1.0, (env, name, mx, ratio, frac)
