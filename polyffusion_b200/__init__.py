"""polyffusion_b200: B200-native (sm_100a) implementation of Polyffusion's DDPM/DDIM sampling hot path.

Drop-in module surface (same import paths below this package as below the reference's
``polyffusion/`` directory): ``stable_diffusion.model.unet.UNetModel``,
``stable_diffusion.latent_diffusion.LatentDiffusion``, ``stable_diffusion.sampler.DiffusionSampler``,
``sampler_sdf.SDFSampler``, ``sampler_ddim.DDIMSampler``, ``ddpm.DenoiseDiffusion``.
All arithmetic runs in hand-written CUDA kernels behind the C ABI in ``include/pf_b200.h``.
"""
__version__ = "0.1.0"
