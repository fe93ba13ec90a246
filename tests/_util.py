"""Shared helpers for the parity tests."""
import torch

_RANDN = torch.randn  # captured before any test monkeypatches torch.randn

SDF_KW = dict(in_channels=2, out_channels=2, channels=64, n_res_blocks=2, attention_levels=[2, 3],
              channel_multipliers=[1, 2, 4, 4], n_heads=4, tf_layers=1)

# BASELINE.json north_star tolerance for the UNet output: rtol 1e-3 / atol 1e-4 (fp32)
RTOL, ATOL = 1e-3, 1e-4


def build_unet(d_cond: int, seed: int = 0):
    """Seeded random-init drop-in UNetModel; the same seed gives the reference's own initialisation
    (tests/test_oracle_vs_reference.py checks that where the reference is importable)."""
    from polyffusion_b200.stable_diffusion.model.unet import UNetModel

    torch.manual_seed(seed)
    return UNetModel(**SDF_KW, d_cond=d_cond).eval()


def oracle_cfg(d_cond: int):
    from oracle.unet_oracle import UNetCfg

    return UNetCfg(d_cond=d_cond)


def close_report(out: torch.Tensor, ref: torch.Tensor, rtol=RTOL, atol=ATOL):
    out, ref = out.detach().cpu().double(), ref.detach().cpu().double()
    diff = (out - ref).abs()
    bound = atol + rtol * ref.abs()
    frac_ok = (diff <= bound).double().mean().item()
    return diff.max().item(), frac_ok


class NoiseTape:
    """Deterministic noise source shared by the CUDA samplers (via monkeypatched torch.randn /
    randn_like) and the CPU oracle (via noise_fn)."""

    def __init__(self, seed: int):
        self.gen = torch.Generator().manual_seed(seed)

    def __call__(self, shape):
        return _RANDN(tuple(shape), generator=self.gen)


class CudaTape:
    """Monkeypatch torch.randn / randn_like so the CUDA samplers draw from a CPU-seeded tape (the same tape the
    golden generator used on the reference: oracle/make_golden.py ``Tape``)."""

    def __init__(self, seed):
        self.tape = NoiseTape(seed)
        self._randn, self._randn_like = torch.randn, torch.randn_like

    def __enter__(self):
        def randn(*size, **kw):
            if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
                size = tuple(size[0])
            return self.tape(size).cuda()

        torch.randn = randn
        torch.randn_like = lambda t, **kw: self.tape(t.shape).cuda()
        return self

    def __exit__(self, *a):
        torch.randn, torch.randn_like = self._randn, self._randn_like
