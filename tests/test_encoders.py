"""Condition encoders: ChordEncoder (RnnEncoder) and TextureEncoder.
CPU: oracle restatement vs golden outputs of the real reference modules (seeded random init at the
sdf_chd8bar / sdf_txt sizes), drop-in parameter tree / init stream; GPU: libpf_b200 kernels through
the drop-in modules vs the oracle and the goldens, rtol 1e-3 / atol 1e-4 (fp32 path)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encoders.npz")
RTOL, ATOL = 1e-3, 1e-4


def build():
    from polyffusion_b200.dl_modules import RnnEncoder, TextureEncoder

    torch.manual_seed(5)
    ce = RnnEncoder(36, 512, 512).eval()
    torch.manual_seed(6)
    te = TextureEncoder(256, 1024, 256, 10).eval()
    return ce, te


def close(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return bool(((a - b).abs() <= ATOL + RTOL * b.abs()).all()), float((a - b).abs().max())


def test_oracle_matches_reference_golden_and_dropin_init_stream():
    """The drop-in modules draw the same init stream as the reference's (same submodules in the same
    order), so the oracle evaluated on THEIR state_dict must reproduce the reference's goldens."""
    from oracle import encoder_oracle as eo
    from oracle.make_golden import encoder_inputs

    g = np.load(GOLD)
    chord, prmat = encoder_inputs()
    ce, te = build()
    mu, scale = eo.chord_encoder(ce.state_dict(), chord)
    assert close(mu, g["chord_mu"])[0] and close(scale, g["chord_scale"])[0]
    for i, seg in enumerate(prmat.split(32, 1)):
        mu, scale = eo.texture_encoder(te.state_dict(), seg)
        assert close(mu, g["txt_mu"][:, i])[0] and close(scale, g["txt_scale"][:, i])[0]
    z = eo.encode_txt(te.state_dict(), prmat)
    assert z.shape == (3, 1, 1024) and close(z[:, 0], g["txt_mu"].reshape(3, -1))[0]


def test_oracle_matches_reference_modules_directly():
    from oracle import encoder_oracle as eo
    from oracle import reference_loader
    from oracle.make_golden import encoder_inputs, reference_encoder_classes

    if not reference_loader.available():
        pytest.skip("reference tree not present (GPU box)")
    Rnn, Txt = reference_encoder_classes()
    chord, prmat = encoder_inputs()
    torch.manual_seed(11)
    ce = Rnn(36, 64, 32).eval()
    te = Txt(48, 40, 24, 10).eval()
    with torch.no_grad():
        d = ce(chord)
        mu, scale = eo.chord_encoder(ce.state_dict(), chord)
        assert torch.allclose(mu, d.mean, atol=1e-6) and torch.allclose(scale, d.scale, rtol=1e-5)
        d = te(prmat[:, :32])
        mu, scale = eo.texture_encoder(te.state_dict(), prmat[:, :32])
        assert torch.allclose(mu, d.mean, atol=1e-5) and torch.allclose(scale, d.scale, rtol=1e-4)
    ours = build()
    assert list(ours[0].state_dict()) == list(Rnn(36, 512, 512).state_dict())
    assert list(ours[1].state_dict()) == list(Txt(256, 1024, 256, 10).state_dict())


def test_cpu_tensors_raise():
    ce, _ = build()
    with pytest.raises(RuntimeError):
        ce(torch.zeros(1, 32, 36))


@pytest.mark.gpu
def test_gpu_chord_encoder_matches_reference():
    from oracle import encoder_oracle as eo
    from oracle.make_golden import encoder_inputs
    from polyffusion_b200.cond import encode_chord

    g = np.load(GOLD)
    chord, _ = encoder_inputs()
    ce, _ = build()
    sd = {k: v.clone() for k, v in ce.state_dict().items()}
    ce = ce.cuda()
    d = ce(chord.cuda())
    ok, err = close(d.mean, g["chord_mu"])
    assert ok, f"chord mu max abs err {err}"
    ok, err = close(d.scale, g["chord_scale"])
    assert ok, f"chord scale max abs err {err}"
    z = encode_chord(ce, chord.cuda())
    assert z.shape == (3, 1, 512) and close(z, eo.encode_chord(sd, chord))[0]
    # batch 64 (the bench batch), against the oracle
    g2 = torch.Generator().manual_seed(3)
    big = (torch.rand(64, 32, 36, generator=g2) < 0.2).float()
    ok, err = close(ce(big.cuda()).mean, eo.chord_encoder(sd, big)[0])
    assert ok, f"batch-64 chord mu max abs err {err}"


@pytest.mark.gpu
def test_gpu_texture_encoder_matches_reference():
    from oracle import encoder_oracle as eo
    from oracle.make_golden import encoder_inputs
    from polyffusion_b200.cond import encode_txt

    g = np.load(GOLD)
    _, prmat = encoder_inputs()
    _, te = build()
    sd = {k: v.clone() for k, v in te.state_dict().items()}
    te = te.cuda()
    for i, seg in enumerate(prmat.split(32, 1)):
        d = te(seg.cuda())
        ok, err = close(d.mean, g["txt_mu"][:, i])
        assert ok, f"txt mu seg {i} max abs err {err}"
        ok, err = close(d.scale, g["txt_scale"][:, i])
        assert ok, f"txt scale seg {i} max abs err {err}"
    z = encode_txt(te, prmat.cuda())
    assert z.shape == (3, 1, 1024)
    ok, err = close(z, eo.encode_txt(sd, prmat))
    assert ok, f"encode_txt max abs err {err}"


def test_cond_glue_batching_matches_reference_segment_loop():
    """cond.encode_txt runs the four 32-step segments as ONE batch; the reference loops over segments
    and concatenates (model_sdf.py:153-164).  Host logic only: a CPU stand-in encoder built on the
    oracle shows the reshapes put every segment's latent in the reference's position."""
    from types import SimpleNamespace

    from oracle import encoder_oracle as eo
    from oracle.make_golden import encoder_inputs
    from polyffusion_b200.cond import encode_chord, encode_txt

    chord, prmat = encoder_inputs()
    ce, te = build()
    sdc, sdt = ce.state_dict(), te.state_dict()
    fake_txt = lambda pr: SimpleNamespace(mean=eo.texture_encoder(sdt, pr)[0])
    fake_chd = lambda ch: SimpleNamespace(mean=eo.chord_encoder(sdc, ch)[0])
    assert torch.allclose(encode_txt(fake_txt, prmat), eo.encode_txt(sdt, prmat), atol=1e-6)
    assert torch.equal(encode_chord(fake_chd, chord), eo.encode_chord(sdc, chord))
    # without encoders: the reference's pass-through / flatten branches
    assert encode_txt(None, prmat) is prmat
    assert encode_chord(None, chord).shape == (3, 1, 32 * 36)
