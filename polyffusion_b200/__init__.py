"""polyffusion_b200: B200-native (sm_100a) implementation of Polyffusion's DDPM/DDIM sampling hot path.

Drop-in module surface (same import paths below this package as below the reference's
``polyffusion/`` directory): ``stable_diffusion.model.unet.UNetModel``,
``stable_diffusion.latent_diffusion.LatentDiffusion``, ``stable_diffusion.sampler.DiffusionSampler``,
``sampler_sdf.SDFSampler``, ``sampler_ddim.DDIMSampler``, ``ddpm.DenoiseDiffusion``,
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
    # only the two encoder SUBMODULES: the reference's dl_modules/__init__.py then picks them up through
    # its own `from .chord_enc import RnnEncoder as ChordEncoder` / `from .txt_enc import TextureEncoder`
    # (the decoders and PianoTree modules stay the reference's)
    "dl_modules.chord_enc": "polyffusion_b200.dl_modules.chord_enc",
    "dl_modules.txt_enc": "polyffusion_b200.dl_modules.txt_enc",
}


def install_dropin() -> None:
    """Alias the reference's top-level module names (``stable_diffusion.model.unet``, ``sampler_sdf``,
    ``sampler_ddim``, ``ddpm`` ...) to this package in ``sys.modules``, so that the reference's own
    ``inference_sdf.py`` / ``models/model_sdf.py`` imports resolve to the CUDA implementation.
    Call it before importing the reference scripts."""
    import importlib
    import sys

    for alias, target in _DROPIN_MODULES.items():
        sys.modules[alias] = importlib.import_module(target)
