"""TEST INFRASTRUCTURE ONLY -- import the real reference (aik2mlj/polyffusion) when it is present.

Recipe verified in SURVEY.md Appendix C: three stub modules (labml.monit, labml_helpers.module,
top-level utils.show_image) + sys.path to /root/reference/polyffusion.  /root/reference exists only
in the build container, never on the GPU box, so callers must handle ``available() == False``.
"""
import os
import sys
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("PF_REFERENCE_ROOT", "/root/reference/polyffusion")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "stable_diffusion"))


def load() -> SimpleNamespace:
    if not available():
        raise ImportError(f"reference not found at {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _SHIM):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, _SHIM)  # shim `utils` must shadow the reference's utils.py (needs pretty_midi)
    from ddpm import DenoiseDiffusion
    from ddpm.unet import UNet
    from sampler_ddim import DDIMSampler
    from sampler_sdf import SDFSampler
    from stable_diffusion.latent_diffusion import LatentDiffusion
    from stable_diffusion.model.unet import UNetModel
    from stable_diffusion.sampler import DiffusionSampler

    return SimpleNamespace(UNetModel=UNetModel, LatentDiffusion=LatentDiffusion, SDFSampler=SDFSampler,
                           DDIMSampler=DDIMSampler, DiffusionSampler=DiffusionSampler,
                           DenoiseDiffusion=DenoiseDiffusion, UNet=UNet)
