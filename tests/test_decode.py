"""Piano-roll decode: prmat2c -> duration matrix / note list.
CPU: the oracle restatement vs golden outputs of the reference's own ``prmat2c_to_prmat``
(+ the lifted reference function itself where /root/reference exists).
GPU: pf_prmat2c_to_prmat / pf_prmat_notes through the C ABI -- integer work, bit-exact."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decode.npz")


def cases():
    from oracle.make_golden import synthetic_prmat2c

    g = np.load(GOLD)
    i = 0
    while f"case{i}_args" in g:
        n, T, seed = (int(v) for v in g[f"case{i}_args"])
        yield synthetic_prmat2c(n, T, seed), g[f"case{i}_prmat"].astype(np.int64), g[f"case{i}_prmat_t"].astype(np.int64)
        i += 1


def test_oracle_matches_reference_golden():
    from oracle import decode_oracle as do

    n_cases = 0
    for x, want, want_t in cases():
        assert np.array_equal(want, want_t)  # numpy and Tensor input branches of the reference agree
        assert np.array_equal(do.prmat2c_to_prmat_fast(x), want)
        n_cases += 1
    assert n_cases == 3
    # the plain-loop form on the smallest case
    x, want, _ = list(cases())[2]
    assert np.array_equal(do.prmat2c_to_prmat(x), want)


def test_oracle_matches_lifted_reference_function():
    from oracle import decode_oracle as do
    from oracle import reference_loader
    from oracle.make_golden import reference_utils_function, synthetic_prmat2c

    if not reference_loader.available():
        pytest.skip("reference tree not present (GPU box)")
    ref = reference_utils_function("prmat2c_to_prmat")
    x = synthetic_prmat2c(1, 64, 99)
    assert np.array_equal(ref(x), do.prmat2c_to_prmat_fast(x))
    assert np.array_equal(ref(x, n_step=16), do.prmat2c_to_prmat_fast(x, 16))


def test_oracle_notes_properties():
    from oracle import decode_oracle as do
    from oracle.make_golden import synthetic_prmat2c

    x = synthetic_prmat2c(2, 128, 5)
    nt = do.notes(x)
    dur = do.prmat2c_to_prmat_fast(x, 128).reshape(2, 128, 128)
    assert len(nt) == np.count_nonzero(dur)
    keys = nt[:, 0].astype(np.int64) * 128 * 128 + nt[:, 1] * 128 + nt[:, 2]
    assert np.all(np.diff(keys) > 0)  # the reference's loop order, strictly increasing
    assert np.all(nt[:, 1] + nt[:, 3] <= 128)  # a note never runs past its segment
    start, end = do.note_times(nt, 128)
    assert np.all(end > start) and np.all(end <= (nt[:, 0] + 1) * 16)
    # the hand-placed cases: a note sustained to the end, an onset on the last step
    assert dur[0, 3, 60] == 125 and dur[0, 127, 61] == 1


@pytest.mark.gpu
def test_gpu_prmat2c_to_prmat_bit_exact():
    from polyffusion_b200.utils import prmat2c_to_prmat

    for x, want, _ in cases():
        got = prmat2c_to_prmat(torch.from_numpy(x).cuda())
        assert got.dtype == np.int64 and got.shape == want.shape
        assert np.array_equal(got, want)
        assert np.array_equal(prmat2c_to_prmat(x), want)  # numpy input, like the reference accepts


@pytest.mark.gpu
def test_gpu_notes_match_oracle_and_full_size_properties():
    from oracle import decode_oracle as do
    from oracle.make_golden import synthetic_prmat2c
    from polyffusion_b200.utils import prmat2c_durations, prmat2c_to_notes

    x = synthetic_prmat2c(3, 128, 21)
    assert np.array_equal(prmat2c_to_notes(torch.from_numpy(x).cuda()), do.notes(x))
    # BASELINE config 5 size: 256 songs x 10 segments; checked through properties and the fast oracle
    big = synthetic_prmat2c(64, 128, 22)
    big = np.concatenate([big] * 4)  # 256 segments
    dur = prmat2c_durations(torch.from_numpy(big).cuda())
    want = do.prmat2c_to_prmat_fast(big, 128).reshape(256, 128, 128)
    assert torch.equal(dur.cpu(), torch.from_numpy(want))
    nt = prmat2c_to_notes(torch.from_numpy(big).cuda())
    assert len(nt) == int((dur > 0).sum())
    keys = nt[:, 0].astype(np.int64) * 128 * 128 + nt[:, 1] * 128 + nt[:, 2]
    assert np.all(np.diff(keys) > 0)
    # empty input: no notes
    z = torch.zeros(2, 2, 128, 128, device="cuda")
    assert prmat2c_to_notes(z).shape == (0, 4)


def test_cpu_input_without_gpu_raises():
    from polyffusion_b200.utils import prmat2c_to_prmat

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        prmat2c_to_prmat(np.zeros((1, 2, 32, 128), dtype=np.float32))
