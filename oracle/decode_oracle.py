"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's piano-roll decode.

Reference followed (relative to /root/reference/polyffusion/): utils.py:240-269 (prmat2c_to_prmat:
onset / sustain channels -> per-onset duration, Python ``round`` = round half to even) and
utils.py:446-470 (the note loop of prmat2c_to_midi_file: one note per onset in (segment, step, pitch)
order, start = t + step/8, end = min(t + (step + dur)/8, t + T/8)).

Pin status: pinned -- tests/test_decode.py compares this restatement with golden outputs of the real
``prmat2c_to_prmat`` (lifted from the reference's utils.py by oracle/make_golden.py; the module
itself needs pretty_midi, which is not installed) and, when /root/reference is present, with the
lifted function directly.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import
this module.
"""
from __future__ import annotations

import numpy as np


def _on(v) -> bool:
    return int(round(float(v))) > 0  # Python round on a float: half to even, like np.float32.__round__


def prmat2c_to_prmat(prmat2c: np.ndarray, n_step: int = 32) -> np.ndarray:
    """Plain loops, for small cases."""
    prmat2c = np.asarray(prmat2c)
    assert prmat2c.ndim == 4
    N, _, T, P = prmat2c.shape
    ratio = T // n_step
    out = np.zeros((N * ratio, n_step, P), dtype=np.int64)
    for n in range(N):
        onset, sustain = prmat2c[n, 0], prmat2c[n, 1]
        for s in range(T):
            for k in range(P):
                if _on(onset[s, k]):
                    dur = 1
                    while s + dur < T and _on(sustain[s + dur, k]):
                        dur += 1
                    out[n * ratio + s // n_step, s % n_step, k] = dur
    return out


def prmat2c_to_prmat_fast(prmat2c: np.ndarray, n_step: int = 32) -> np.ndarray:
    """Vectorised numpy form of the same definition (np.rint = round half to even)."""
    x = np.asarray(prmat2c, dtype=np.float32)
    N, _, T, P = x.shape
    on = np.rint(x[:, 0]) > 0
    su = np.rint(x[:, 1]) > 0
    run = np.zeros((N, P), dtype=np.int64)  # consecutive sustained steps starting at s + 1
    dur = np.zeros((N, T, P), dtype=np.int64)
    for s in range(T - 1, -1, -1):
        dur[:, s] = np.where(on[:, s], 1 + run, 0)
        run = np.where(su[:, s], run + 1, 0)
    return dur.reshape(N * (T // n_step), n_step, P)


def notes(prmat2c: np.ndarray) -> np.ndarray:
    """(segment, step, pitch, dur) rows in the order prmat2c_to_midi_file appends notes."""
    x = np.asarray(prmat2c)
    N, _, T, P = x.shape
    dur = prmat2c_to_prmat_fast(x, T).reshape(N, T, P)
    seg, step, key = np.nonzero(dur)  # C order = (segment, step, pitch) lexicographic
    return np.stack([seg, step, key, dur[seg, step, key]], axis=1).astype(np.int32)


def note_times(note_rows: np.ndarray, T: int):
    """Start / end seconds as prmat2c_to_midi_file computes them (utils.py:446-467)."""
    t_bar = int(T / 8)
    t0 = note_rows[:, 0].astype(np.float64) * t_bar
    start = t0 + note_rows[:, 1] * 1 / 8
    end = np.minimum(t0 + (note_rows[:, 1] + note_rows[:, 3]) * 1 / 8, t0 + t_bar)
    return start, end
