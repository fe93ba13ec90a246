"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the REAL reference.

Run in the build container (where /root/reference exists):  python -m oracle.make_golden
The reference ships no golden vectors of its own, so these files pin the
oracle restatement and the CUDA path to outputs of the unmodified reference modules, imported through
oracle/reference_loader.py, with seeded random-init weights (torch.manual_seed(0); the drop-in
UNetModel reproduces the same initialisation from the same seed on any box).
Noise is injected by monkeypatching torch.randn / torch.randn_like with a seeded tape so the reference
samplers (which draw on-device noise internally) are reproducible.
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import reference_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
KW = dict(in_channels=2, out_channels=2, channels=64, n_res_blocks=2, attention_levels=[2, 3],
          channel_multipliers=[1, 2, 4, 4], n_heads=4, tf_layers=1)


class Tape:
    """Seeded replacement for torch.randn / torch.randn_like (CPU)."""

    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)
        self._randn = torch.randn

    def randn(self, *size, **kw):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            size = tuple(size[0])
        return self._randn(tuple(size), generator=self.gen)

    def randn_like(self, t, **kw):
        return self._randn(tuple(t.shape), generator=self.gen)

    def __enter__(self):
        torch.randn, torch.randn_like = self.randn, self.randn_like
        return self

    def __exit__(self, *a):
        torch.randn = self._randn
        torch.randn_like = torch._C._VariableFunctions.randn_like


def save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = reference_loader.load()
    torch.set_num_threads(os.cpu_count() or 1)

    # ---- 1. UNet forward, sdf_chd8bar (d_cond 512, n_cond 1)
    torch.manual_seed(0)
    unet = ref.UNetModel(**KW, d_cond=512).eval()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 2, 128, 128, generator=g)
    cond = torch.randn(2, 1, 512, generator=g)
    t = torch.tensor([999, 3])
    with torch.no_grad():
        eps = unet(x, t, cond)
    save("unet_chd8bar_b2.npz", x=x, t=t, cond=cond, eps=eps, seed=0)

    # ---- 2. UNet forward, sdf_txtvnl geometry (d_cond 128, n_cond 128)
    torch.manual_seed(0)
    unet_v = ref.UNetModel(**KW, d_cond=128).eval()
    xv = torch.randn(1, 2, 128, 128, generator=g)
    cv = torch.randn(1, 128, 128, generator=g)
    tv = torch.tensor([417])
    with torch.no_grad():
        ev = unet_v(xv, tv, cv)
    save("unet_txtvnl_b1.npz", x=xv, t=tv, cond=cv, eps=ev, seed=0)

    # ---- 3. schedules and sampler tables
    ldm = ref.LatentDiffusion(unet, None, 0.18215, 1000, 0.00085, 0.012)
    sdf = ref.SDFSampler(ldm)
    d50 = ref.DDIMSampler(ldm, 50, "uniform", 0.0)
    dq = ref.DDIMSampler(ldm, 20, "quad", 0.5)
    save("tables.npz",
         alpha=ldm.alpha.data, beta=ldm.beta.data, alpha_bar=ldm.alpha_bar.data,
         sdf_sqrt_ab=sdf.sqrt_alpha_bar, sdf_sqrt_1m_ab=sdf.sqrt_1m_alpha_bar,
         sdf_sqrt_recip_ab=sdf.sqrt_recip_alpha_bar, sdf_sqrt_recip_m1_ab=sdf.sqrt_recip_m1_alpha_bar,
         sdf_log_var=sdf.log_var, sdf_mean_x0=sdf.mean_x0_coef, sdf_mean_xt=sdf.mean_xt_coef,
         d50_tau=d50.time_steps, d50_alpha=d50.ddim_alpha, d50_alpha_sqrt=d50.ddim_alpha_sqrt,
         d50_alpha_prev=d50.ddim_alpha_prev, d50_sigma=d50.ddim_sigma,
         d50_sqrt_1m_alpha=d50.ddim_sqrt_one_minus_alpha,
         dq_tau=dq.time_steps, dq_alpha=dq.ddim_alpha, dq_alpha_prev=dq.ddim_alpha_prev,
         dq_sigma=dq.ddim_sigma, dq_sqrt_1m_alpha=dq.ddim_sqrt_one_minus_alpha)

    # ---- 4. SDFSampler.paint (RePaint), 3 steps, CFG scale 5, non-trivial orig/mask, taped noise
    g2 = torch.Generator().manual_seed(21)
    orig = (torch.rand(1, 2, 128, 128, generator=g2) < 0.02).float()
    mask = torch.zeros(1, 2, 128, 128)
    mask[:, :, :, 60:] = 1.0
    cond1 = torch.randn(1, 1, 512, generator=g2)
    uncond = -torch.ones(1, 1, 512)
    x_start = torch.randn(1, 2, 128, 128, generator=g2)
    with Tape(31), torch.no_grad():
        xt = sdf.q_sample(orig, 2, torch.randn(1, 2, 128, 128))
        out = sdf.paint(xt, cond1, 2, orig=orig, mask=mask, uncond_scale=5.0, uncond_cond=uncond)
    save("paint_ddpm_cfg5.npz", orig=orig, mask=mask, cond=cond1, uncond=uncond, tape_seed=31,
         t_start=2, uncond_scale=5.0, x_t=xt, out=out)
    with Tape(32), torch.no_grad():
        out2 = sdf.paint(x_start, cond1, 1, orig=orig, mask=mask, repaint_n=2)
    save("paint_ddpm_repaint2.npz", orig=orig, mask=mask, cond=cond1, tape_seed=32, t_start=1,
         x_start=x_start, out=out2)
    with Tape(33), torch.no_grad():
        out3 = sdf.sample([1, 2, 128, 128], cond1, x_last=x_start, t_start=997)
    save("sample_ddpm.npz", cond=cond1, tape_seed=33, t_start=997, x_start=x_start, out=out3)

    # ---- 5. DDIM sample / paint
    d4 = ref.DDIMSampler(ldm, 4, "uniform", 0.0)
    with Tape(41), torch.no_grad():
        o_s = d4.sample([1, 2, 128, 128], cond1, x_last=x_start, t_start=1)
    d4e = ref.DDIMSampler(ldm, 4, "uniform", 1.0)
    orig_noise = torch.randn(1, 2, 128, 128, generator=g2)
    with Tape(42), torch.no_grad():
        o_p = d4e.paint(x_start, cond1, 2, orig=orig, mask=mask, orig_noise=orig_noise)
    save("ddim.npz", cond=cond1, x_start=x_start, orig=orig, mask=mask, orig_noise=orig_noise,
         sample_tape_seed=41, sample_out=o_s, sample_t_start=1, paint_tape_seed=42, paint_out=o_p,
         paint_t_start=2, paint_eta=1.0)

    # ---- 6. legacy DenoiseDiffusion.p_sample plumbing (BASELINE config 1 shape: B=4, 10 steps, CPU)
    # eps model: a fixed, seeded 3x3 conv + time shift (the legacy 168 M-parameter UNet is out of the
    # CUDA scope; this pins the step arithmetic and table indexing of ddpm/__init__.py:66-88)
    class TinyEps(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(7)
            self.conv = torch.nn.Conv2d(2, 2, 3, padding=1)

        def forward(self, x, t):
            return self.conv(x) + (t.float() / 1000.0)[:, None, None, None]

    dd = ref.DenoiseDiffusion(TinyEps(), 1000)
    xT = torch.randn(4, 2, 128, 128, generator=g2)
    xx = xT
    with Tape(51), torch.no_grad():
        for ti in range(999, 989, -1):
            xx = dd.p_sample(xx, xx.new_full((4,), ti, dtype=torch.long))
    save("legacy_ddpm.npz", x_T=xT, tape_seed=51, out=xx, beta=dd.beta, alpha_bar=dd.alpha_bar)


def reference_glue_functions():
    """get_mask / get_autoreg_data lifted from inference_sdf.py:121-193 by ast (the module itself
    needs pretty_midi / omegaconf / matplotlib)."""
    import ast

    src = open(os.path.join(reference_loader.REFERENCE_ROOT, "inference_sdf.py")).read()
    ns = {"torch": torch}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("get_mask", "get_autoreg_data"):
            exec(compile(ast.Module([node], []), "inference_sdf.py", "exec"), ns)
    return ns["get_mask"], ns["get_autoreg_data"]


def synthetic_melody(n_seg, seed, blank_first=0, pitch0=False):
    """Synthetic melody prmat2c (BASELINE config 5): per time step an onset at pitch U{60..84} with
    probability 0.25 (sometimes a second, lower one) in channel 0; binary fp32."""
    g = torch.Generator().manual_seed(seed)
    o = torch.zeros(n_seg, 2, 128, 128)
    for b in range(n_seg):
        for t in range(128):
            if b == 0 and t < blank_first:
                continue
            if torch.rand(1, generator=g) < 0.25:
                o[b, 0, t, int(torch.randint(60, 85, (1,), generator=g))] = 1
                if torch.rand(1, generator=g) < 0.3:
                    o[b, 0, t, int(torch.randint(30, 60, (1,), generator=g))] = 1
    if pitch0:
        o[0, 0, 5, 0] = 1
    return o


def make_mask_golden():
    get_mask, get_autoreg = reference_glue_functions()
    out = {}
    for i, (n_seg, seed, blank, p0) in enumerate([(2, 1, 0, False), (3, 2, 7, False), (1, 3, 0, True), (4, 4, 20, True)]):
        o = synthetic_melody(n_seg, seed, blank, p0)
        out[f"case{i}_args"] = np.asarray([n_seg, seed, blank, int(p0)])
        # masks are binary: store packed bits
        out[f"case{i}_below"] = np.packbits(get_mask(o, "below").contiguous().numpy().astype(np.uint8))
        out[f"case{i}_above"] = np.packbits(get_mask(o, "above").contiguous().numpy().astype(np.uint8))
    g = torch.Generator().manual_seed(9)
    d = torch.randn(4, 2, 8, 4, generator=g)
    out["autoreg_in"] = d.numpy()
    out["autoreg_out"] = get_autoreg(d, 2).numpy()
    save("mask_autoreg.npz", **out)



def reference_experiments_class():
    """``Experiments`` (and the helpers its ``predict`` calls) lifted from inference_sdf.py:121-303 by ast: the module
    itself cannot be imported here (omegaconf, pretty_midi, lightning, the data pipeline), and its ``predict`` reads
    the module globals ``args`` and ``device``.  Returns (Experiments, namespace) with those globals set for a CPU
    DDPM run (``args.ddim = False``, ``args.repaint_n = 1``)."""
    import ast
    from types import SimpleNamespace
    from typing import Optional

    src = open(os.path.join(reference_loader.REFERENCE_ROOT, "inference_sdf.py")).read()
    ns = {"torch": torch, "np": np, "Optional": Optional, "device": "cpu", "DiffusionSampler": object,
          "args": SimpleNamespace(ddim=False, ddim_steps=None, repaint_n=1)}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("get_mask", "get_autoreg_data"):
            exec(compile(ast.Module([node], []), "inference_sdf.py", "exec"), ns)
        if isinstance(node, ast.ClassDef) and node.name == "Experiments":
            # only __init__ and predict: the other methods need the data pipeline / MIDI writers
            node.body = [n for n in node.body if isinstance(n, ast.FunctionDef) and n.name in ("__init__", "predict")]
            exec(compile(ast.Module([node], []), "inference_sdf.py", "exec"), ns)
    return ns["Experiments"], ns


def make_predict_golden():
    """The reference's own ``Experiments.predict(autoreg=True)`` (inference_sdf.py:202-283) on the CPU with the
    reference's SDFSampler / UNetModel: one song of 2 segments -> 3 half-overlapping windows, ``params.n_steps = 2``
    (t_idx = 1: two reverse steps per window), inpaint type "below", with and without classifier-free guidance."""
    import contextlib
    import io
    from types import SimpleNamespace

    ref = reference_loader.load()
    torch.set_num_threads(os.cpu_count() or 1)
    Experiments, ns = reference_experiments_class()
    torch.manual_seed(0)
    unet = ref.UNetModel(**KW, d_cond=512).eval()
    ldm = ref.LatentDiffusion(unet, None, 0.18215, 1000, 0.00085, 0.012)
    sdf = ref.SDFSampler(ldm)
    params = SimpleNamespace(out_channels=2, img_h=128, img_w=128, d_cond=512, n_steps=2)
    exp = Experiments("sdf", params, sdf)
    g = torch.Generator().manual_seed(71)
    orig = synthetic_melody(2, 31)
    mask = ns["get_mask"](orig, "below")
    cond = torch.randn(2, 1, 512, generator=g)
    cond_mid = torch.randn(2, 1, 512, generator=g)
    out = {}
    for tag, scale, seed in (("plain", 1.0, 61), ("cfg", 2.0, 62)):
        with Tape(seed), contextlib.redirect_stdout(io.StringIO()):
            gen = exp.predict(cond, cond_mid, uncond_scale=scale, autoreg=True, orig=orig.clone(), mask=mask.clone())
        out[f"{tag}_out"] = gen
        out[f"{tag}_tape_seed"] = seed
        out[f"{tag}_scale"] = scale
    save("predict_autoreg.npz", orig=orig, mask=mask, cond=cond, cond_mid=cond_mid, t_idx=params.n_steps - 1, **out)

def reference_utils_function(name):
    """A pure-numpy function lifted from the reference's utils.py by ast (the module imports
    pretty_midi / matplotlib at the top and cannot be imported here)."""
    import ast

    src = open(os.path.join(reference_loader.REFERENCE_ROOT, "utils.py")).read()
    ns = {"np": np, "torch": torch}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            exec(compile(ast.Module([node], []), "utils.py", "exec"), ns)
    return ns[name]


def synthetic_prmat2c(n, T, seed):
    """prmat2c-like tensor with the value classes the decode has to distinguish: exact 0 / 1, sampler
    noise around them, the rounding boundaries 0.5 / 1.5 / 2.5, negatives, long sustains that run to
    the end of the segment, sustain without onset."""
    rng = np.random.default_rng(seed)
    vals = np.asarray([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.3, 0.5, 0.5000001, 0.4999999, 0.6, 1.5, 2.5,
                       -0.7, -0.5, 0.97, 1.02, 0.04], dtype=np.float32)
    x = np.zeros((n, 2, T, 128), dtype=np.float32)
    x[:, 0] = vals[rng.integers(0, len(vals), size=(n, T, 128))] * (rng.random((n, T, 128)) < 0.15)
    x[:, 1] = vals[rng.integers(0, len(vals), size=(n, T, 128))] * (rng.random((n, T, 128)) < 0.6)
    x[0, 0, 3, 60] = 1.0
    x[0, 1, 4:, 60] = 1.0   # sustained to the end of the segment
    x[0, 0, T - 1, 61] = 1.0  # onset on the last step
    return x


def reference_encoder_classes():
    """RnnEncoder / TextureEncoder loaded by FILE PATH: the dl_modules package import fails on
    pretty_midi (pianotree_dec.py), these two files are pure torch."""
    import importlib.util

    out = []
    for fname, cls in (("chord_enc.py", "RnnEncoder"), ("txt_enc.py", "TextureEncoder")):
        spec = importlib.util.spec_from_file_location(
            "ref_" + fname[:-3], os.path.join(reference_loader.REFERENCE_ROOT, "dl_modules", fname))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out.append(getattr(mod, cls))
    return out


def encoder_inputs():
    """Synthetic chord matrices [B, 32, 36] (root one-hot | chroma multi-hot | bass one-hot, as
    data/ chord features are laid out) and binary piano rolls [B, 128, 128]."""
    g = torch.Generator().manual_seed(77)
    B = 3
    chord = torch.zeros(B, 32, 36)
    for b in range(B):
        for t in range(32):
            chord[b, t, int(torch.randint(0, 12, (1,), generator=g))] = 1
            chord[b, t, 12:24] = (torch.rand(12, generator=g) < 0.3).float()
            chord[b, t, 24 + int(torch.randint(0, 12, (1,), generator=g))] = 1
    prmat = (torch.rand(B, 128, 128, generator=g) < 0.03).float() * torch.randint(1, 9, (B, 128, 128), generator=g)
    return chord, prmat


def make_encoder_golden():
    """Seeded random-init reference encoders at the sdf_chd8bar / sdf_txt sizes (no checkpoint is in
    the tree): outputs of the real modules on the synthetic inputs."""
    Rnn, Txt = reference_encoder_classes()
    chord, prmat = encoder_inputs()
    torch.manual_seed(5)
    ce = Rnn(36, 512, 512).eval()
    torch.manual_seed(6)
    te = Txt(256, 1024, 256, 10).eval()
    with torch.no_grad():
        dc = ce(chord)
        zs = [te(seg) for seg in prmat.split(32, 1)]
    save("encoders.npz", chord_mu=dc.mean, chord_scale=dc.scale,
         txt_mu=torch.stack([z.mean for z in zs], 1), txt_scale=torch.stack([z.scale for z in zs], 1))


def make_decode_golden():
    ref = reference_utils_function("prmat2c_to_prmat")
    out = {}
    for i, (n, T, seed) in enumerate([(3, 128, 11), (2, 64, 12), (1, 32, 13)]):
        x = synthetic_prmat2c(n, T, seed)
        out[f"case{i}_args"] = np.asarray([n, T, seed])
        out[f"case{i}_prmat"] = ref(x).astype(np.int16)  # durations <= 128
        out[f"case{i}_prmat_t"] = ref(torch.from_numpy(x)).astype(np.int16)  # Tensor input branch
    save("decode.npz", **out)


def make_legacy_unet_golden():
    """BASELINE config 1 with the REAL legacy eps-model: ddpm.unet.UNet(2, 64, [1,2,2,4], [F,F,F,T])
    (params/ddpm.yaml:13-27; 167.8 M parameters, seeded init), one evaluation at four timesteps and
    DenoiseDiffusion.p_sample over 10 reverse steps (999..990) at B = 4 with taped noise, on the CPU."""
    ref = reference_loader.load()
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    unet = ref.UNet(2, 64, [1, 2, 2, 4], [False, False, False, True]).eval()
    g = torch.Generator().manual_seed(61)
    x = torch.randn(4, 2, 128, 128, generator=g)
    t = torch.tensor([999, 500, 3, 250])
    with torch.no_grad():
        eps = unet(x, t)
    dd = ref.DenoiseDiffusion(unet, 1000)
    xx = x
    with Tape(62), torch.no_grad():
        for ti in range(999, 989, -1):
            xx = dd.p_sample(xx, xx.new_full((4,), ti, dtype=torch.long))
    save("legacy_unet.npz", x=x, t=t, eps=eps, seed=0, tape_seed=62, out=xx)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "legacy_unet":
        os.makedirs(OUT, exist_ok=True)
        make_legacy_unet_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "decode":
        os.makedirs(OUT, exist_ok=True)
        make_decode_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "predict":
        os.makedirs(OUT, exist_ok=True)
        make_predict_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "encoders":
        os.makedirs(OUT, exist_ok=True)
        make_encoder_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "masks":
        os.makedirs(OUT, exist_ok=True)
        reference_loader.load()
        make_mask_golden()
    else:
        main()
        make_mask_golden()
        make_decode_golden()
        make_predict_golden()
        make_encoder_golden()
        make_legacy_unet_golden()
