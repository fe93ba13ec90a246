"""polyffusion_b200: B200-native (sm_100a) implementation of Polyffusion's DDPM/DDIM sampling hot path.

Drop-in module surface (same import paths below this package as below the reference's
``polyffusion/`` directory): ``stable_diffusion.model.unet.UNetModel``,
``stable_diffusion.latent_diffusion.LatentDiffusion``, ``stable_diffusion.sampler.DiffusionSampler``,
``sampler_sdf.SDFSampler``, ``sampler_ddim.DDIMSampler``, ``ddpm.DenoiseDiffusion``, ``ddpm.unet.UNet``,
``dl_modules.chord_enc.RnnEncoder``, ``dl_modules.txt_enc.TextureEncoder``; plus ``autoreg``, ``cond``
and ``utils`` (song-batched autoregressive driver, condition glue, piano-roll decode).
All arithmetic runs in hand-written CUDA kernels behind the C ABI in ``include/pf_b200.h``.
"""
__version__ = "0.1.0"


_DROPIN_MODULES = {
    "stable_diffusion": "polyffusion_b200.stable_diffusion",
    "stable_diffusion.model": "polyffusion_b200.stable_diffusion.model",
    "stable_diffusion.model.unet": "polyffusion_b200.stable_diffusion.model.unet",
    "stable_diffusion.model.unet_attention": "polyffusion_b200.stable_diffusion.model.unet_attention",
    "stable_diffusion.latent_diffusion": "polyffusion_b200.stable_diffusion.latent_diffusion",
    "stable_diffusion.sampler": "polyffusion_b200.stable_diffusion.sampler",
    "sampler_sdf": "polyffusion_b200.sampler_sdf",
    "sampler_ddim": "polyffusion_b200.sampler_ddim",
    "ddpm": "polyffusion_b200.ddpm",
    "ddpm.unet": "polyffusion_b200.ddpm.unet",
    "ddpm.utils": "polyffusion_b200.ddpm.utils",
    # only the two encoder SUBMODULES: the reference's dl_modules/__init__.py then picks them up through
    # its own `from .chord_enc import RnnEncoder as ChordEncoder` / `from .txt_enc import TextureEncoder`
    # (the decoders and PianoTree modules stay the reference's)
    "dl_modules.chord_enc": "polyffusion_b200.dl_modules.chord_enc",
    "dl_modules.txt_enc": "polyffusion_b200.dl_modules.txt_enc",
}


# alias packages that replace only PART of a reference package: their __path__ is extended with the
# reference's own directory so that every submodule this package does not provide still resolves
_DROPIN_PACKAGES = ("stable_diffusion", "stable_diffusion.model", "stable_diffusion.sampler", "ddpm")


def _reference_dirs(pkg: str, reference_dir=None):
    """Directories named like the package `pkg` (dotted) below the reference checkout(s): the explicit
    `reference_dir` (the reference's ``polyffusion/`` directory) and every ``sys.path`` entry."""
    import os
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    rel = os.path.join(*pkg.split("."))
    roots = ([reference_dir] if reference_dir else []) + [p or os.getcwd() for p in sys.path]
    out = []
    for root in roots:
        d = os.path.join(root, rel)
        if os.path.isdir(d) and not os.path.abspath(d).startswith(here) and d not in out:
            out.append(d)
    return out


def install_dropin(reference_dir=None) -> None:
    """Alias the reference's module names (``stable_diffusion.model.unet``, ``sampler_sdf``,
    ``sampler_ddim``, ``ddpm`` ...) to this package in ``sys.modules``, so that the reference's own
    ``inference_sdf.py`` / ``inference.py`` / ``models/model_sdf.py`` / ``train/*.py`` imports resolve
    to the CUDA implementation.  Call it before importing the reference scripts, with the reference's
    ``polyffusion/`` directory on ``sys.path`` (it is ``sys.path[0]`` when its scripts run) or passed
    as ``reference_dir``.

    Only what this package implements is replaced.  The alias packages keep the reference's own
    directories on their ``__path__``, so modules this package does not provide --
    ``stable_diffusion.util``, ``stable_diffusion.model.autoencoder``, ``stable_diffusion.losses``,
    ``stable_diffusion.sampler.ddim`` / ``.ddpm``, ``ddpm.sampling`` / ``.training`` -- import from
    the reference unchanged (and see the drop-in classes through their relative imports)."""
    import importlib
    import sys

    for alias, target in _DROPIN_MODULES.items():
        mod = importlib.import_module(target)
        sys.modules[alias] = mod
        if alias in _DROPIN_PACKAGES:
            for d in _reference_dirs(alias, reference_dir):
                if d not in mod.__path__:
                    mod.__path__.append(d)
