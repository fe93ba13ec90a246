"""ctypes binding of libpf_b200.so (C ABI declared in include/pf_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (``make -C polyffusion_b200/csrc``).
There is deliberately no fallback: if the library is missing, or no sm_100 GPU is present when a
compute entry point is called, the call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# PF_B200_LIB selects a variant build of the same sources (A/B measurements, csrc/Makefile)
LIB_PATH = os.environ.get("PF_B200_LIB") or os.path.join(_HERE, "libpf_b200.so")


class PfError(RuntimeError):
    """Raised when a libpf_b200 entry point returns a non-zero status."""


class UNetCfg(Structure):
    _fields_ = [
        ("in_channels", c_int32),
        ("out_channels", c_int32),
        ("channels", c_int32),
        ("n_res_blocks", c_int32),
        ("n_levels", c_int32),
        ("channel_multipliers", c_int32 * 8),
        ("attention_levels", c_int32 * 8),
        ("n_heads", c_int32),
        ("tf_layers", c_int32),
        ("d_cond", c_int32),
    ]


class FusedStepArgs(Structure):
    """pf_fused_step (include/pf_b200.h)."""
    _fields_ = [
        ("kind", c_int32),
        ("flags", c_int32),
        ("index", c_void_p),
        ("coef", c_void_p),
        ("t_table", c_void_p),
        ("x", c_void_p),
        ("eps_out", c_void_p),
        ("noise", c_void_p),
        ("noise_kn", c_void_p),
        ("orig", c_void_p),
        ("mask", c_void_p),
        ("temperature", c_float),
        ("seed", ctypes.c_uint64),
        ("sample0", c_int64),
    ]


class StepArgs(Structure):
    _fields_ = [
        ("x", c_void_p),
        ("e_cond", c_void_p),
        ("e_uncond", c_void_p),
        ("noise", c_void_p),
        ("orig", c_void_p),
        ("mask", c_void_p),
        ("noise_kn", c_void_p),
        ("x_prev", c_void_p),
        ("x0", c_void_p),
        ("e_t", c_void_p),
        ("n", c_int64),
        ("noise_bcast", c_int64),
        ("uncond_scale", c_float),
        ("c0", c_float),
        ("c1", c_float),
        ("c2", c_float),
        ("c3", c_float),
        ("c4", c_float),
        ("temperature", c_float),
        ("kn_a", c_float),
        ("kn_b", c_float),
    ]


# every symbol include/pf_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pf_last_error": (c_char_p, []),
    "pf_version": (c_char_p, []),
    "pf_unet_create": (c_int32, [POINTER(UNetCfg), POINTER(c_void_p)]),
    "pf_unet_destroy": (None, [c_void_p]),
    "pf_unet_set_weight": (c_int32, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int32]),
    "pf_unet_finalize": (c_int32, [c_void_p, c_void_p]),
    "pf_unet_workspace_bytes": (c_size_t, [c_void_p, c_int32, c_int32, c_int32, c_int32]),
    "pf_unet_forward": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p,
         c_void_p, c_size_t, c_void_p],
    ),
    "pf_unet_forward_step": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, POINTER(FusedStepArgs),
         c_void_p, c_size_t, c_void_p],
    ),
    "pf_unet_prepare_cond": (
        c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_size_t, c_void_p]),
    "pf_unet_enable_time_lut": (c_int32, [c_void_p, c_int32, c_void_p]),
    "pf_fill_normal": (c_int32, [c_void_p, c_int64, c_int64, ctypes.c_uint64, c_int64, c_int32, c_int32, c_void_p]),
    "pf_unet_forward_profiled": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p,
         c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p],
    ),
    "pf_unet_op_desc": (c_int32, [c_void_p, c_int32, c_char_p, c_int32]),
    "pf_unet_launch_count": (c_int32, [c_void_p]),
    "pf_sample_step_ddpm": (c_int32, [POINTER(StepArgs), c_void_p]),
    "pf_sample_step_ddim": (c_int32, [POINTER(StepArgs), c_void_p]),
    "pf_sample_step_ddpm_legacy": (c_int32, [POINTER(StepArgs), c_void_p]),
    "pf_q_sample": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p]),
    "pf_get_mask": (
        c_int32,
        [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p],
    ),
    "pf_linear": (
        c_int32,
        [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p],
    ),
    "pf_gru_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "pf_gru_bidir_last": (
        c_int32,
        [c_void_p, c_int32, c_int32, c_int32, c_int32, POINTER(c_void_p), POINTER(c_void_p),
         POINTER(c_void_p), POINTER(c_void_p), c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "pf_txt_cnn": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "pf_prmat2c_to_prmat": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "pf_prmat_notes": (
        c_int32,
        [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int64, POINTER(c_int64), c_void_p],
    ),
    "pf_op_conv2d_nhwc": (
        c_int32,
        [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32,
         c_void_p, c_void_p, c_void_p, c_int32, c_void_p],
    ),
    "pf_op_conv2d_nhwc_ex": (
        c_int32,
        [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32,
         c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p],
    ),
    "pf_op_groupnorm_generic": (
        c_int32,
        [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_float, c_int32, c_void_p, c_void_p],
    ),
    "pf_op_softmax_rows": (c_int32, [c_void_p, c_float, c_void_p, c_int64, c_int32, c_void_p]),
    "pf_op_time_sincos": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "pf_op_conv3x3_direct": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
         c_void_p],
    ),
    "pf_op_attention": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p],
    ),
    "pf_op_groupnorm_nhwc": (
        c_int32,
        [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_float, c_int32, c_void_p, c_void_p],
    ),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load the library once; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PfError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C polyffusion_b200/csrc`). polyffusion_b200 has no CPU/PyTorch fallback."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise PfError(lib().pf_last_error().decode("utf-8", "replace"))


def ptr(t) -> c_void_p:
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def current_stream():
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)
