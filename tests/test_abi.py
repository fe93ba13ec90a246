"""CPU: the C-ABI library loads and exports every symbol include/pf_b200.h declares; error paths
that need no GPU behave as documented (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from polyffusion_b200 import _lib

    lib = _lib.lib()
    names = header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"libpf_b200.so does not export {n}"
    # the ctypes table binds exactly the header's functions
    assert sorted(_lib.SYMBOLS) == names


def test_version_string():
    from polyffusion_b200 import _lib

    assert b"sm_100a" in _lib.lib().pf_version()


def test_struct_layouts_match_header():
    from polyffusion_b200 import _lib

    # pf_unet_cfg: 5 + 8 + 8 + 3 int32; pf_step_args: 10 pointers + 2 int64 + 9 floats (+ pad)
    assert ctypes.sizeof(_lib.UNetCfg) == 4 * 24
    assert ctypes.sizeof(_lib.StepArgs) == 10 * 8 + 2 * 8 + 9 * 4 + 4


def test_create_without_gpu_fails_loudly():
    import torch

    from polyffusion_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = _lib.UNetCfg()
    h = ctypes.c_void_p()
    rc = _lib.lib().pf_unet_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0
    assert b"no CUDA device" in _lib.lib().pf_last_error() or b"CUDA" in _lib.lib().pf_last_error()


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (or torch CPU fallbacks of the hot path)."""
    pkg = os.path.join(ROOT, "polyffusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
