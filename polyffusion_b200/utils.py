"""GPU piano-roll decode.

Mirrors ``utils.prmat2c_to_prmat`` (reference ``utils.py:240-269``) and the note loop of
``utils.prmat2c_to_midi_file`` (``utils.py:446-470``): the reference walks every (segment, step, pitch)
cell in Python -- O(N * 128 * 128) interpreter iterations per batch, which dominates wall-clock once
sampling is fast.  Here both are CUDA kernels behind ``pf_prmat2c_to_prmat`` / ``pf_prmat_notes``
(include/pf_b200.h); results are bit-identical integers.  CUDA only, no CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ._lib import check, current_stream, lib, ptr


def _as_cuda(prmat2c) -> torch.Tensor:
    t = torch.as_tensor(prmat2c)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("polyffusion_b200.utils runs on CUDA only (no CPU fallback)")
        t = t.cuda()
    if t.dim() != 4 or t.shape[1] < 2:
        raise AssertionError(f"prmat2c must be (N, 2, T, P), got {tuple(t.shape)}")
    return t.detach().contiguous().float()


def prmat2c_durations(prmat2c) -> torch.Tensor:
    """Duration matrix [N, T, P] int64 on the GPU (0 = no onset)."""
    x = _as_cuda(prmat2c)
    n, c, t, p = x.shape
    out = torch.empty((n, t, p), dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().pf_prmat2c_to_prmat(ptr(x), n, c, t, p, ptr(out), current_stream()))
    return out


def prmat2c_to_prmat(prmat2c, n_step: int = 32) -> np.ndarray:
    """Drop-in for ``utils.prmat2c_to_prmat``: (N, 2, 32*ratio, 128) -> (N*ratio, n_step, 128) int64
    numpy array of note durations."""
    dur = prmat2c_durations(prmat2c)
    n, t, p = dur.shape
    ratio = t // n_step
    return dur.reshape(n * ratio, n_step, p).cpu().numpy()


def prmat2c_to_notes(prmat2c) -> np.ndarray:
    """Notes in the order ``prmat2c_to_midi_file`` appends them: int32 array [n_notes, 4] of
    (segment, step, pitch, duration in steps).  A note starts at ``t_seg + step / 8`` seconds and
    ends at ``min(t_seg + (step + dur) / 8, t_seg + T / 8)`` (utils.py:462-467)."""
    dur = prmat2c_durations(prmat2c)
    n, t, p = dur.shape
    rows = n * t
    with torch.cuda.device(dur.device):
        offsets = torch.empty(rows + 1, dtype=torch.int32, device=dur.device)
        total = ctypes.c_int64(0)
        check(lib().pf_prmat_notes(ptr(dur), rows, p, ptr(offsets), None, 0, ctypes.byref(total),
                                   current_stream()))
        notes = torch.empty((max(total.value, 1), 3), dtype=torch.int32, device=dur.device)
        check(lib().pf_prmat_notes(ptr(dur), rows, p, ptr(offsets), ptr(notes), total.value,
                                   ctypes.byref(total), current_stream()))
    nt = notes[: total.value].cpu().numpy()
    out = np.empty((total.value, 4), dtype=np.int32)
    out[:, 0] = nt[:, 0] // t
    out[:, 1] = nt[:, 0] % t
    out[:, 2] = nt[:, 1]
    out[:, 3] = nt[:, 2]
    return out
