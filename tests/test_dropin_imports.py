"""CPU, build container only: after ``polyffusion_b200.install_dropin()`` every import line of the
reference's drivers that touches a replaced package still resolves -- replaced classes come from
polyffusion_b200, everything else falls through to the reference's own files.

Import sites exercised (paths under /root/reference/polyffusion/): inference.py:12-14,
train/train_ddpm.py:2-3, models/model_autoencoder.py:5, train/train_autoencoder.py:8,
inference_sdf.py:37-41, train/train_ldm.py:9-10.  Runs in a subprocess so the aliases do not leak into
the other tests' ``sys.modules``."""
import os
import subprocess
import sys

import pytest

from oracle import reference_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not reference_loader.available(), reason="reference tree not present")

SCRIPT = r"""
import sys, warnings
warnings.filterwarnings("ignore")
sys.path[:0] = [{shim!r}, {ref!r}, {root!r}]   # the reference's scripts run with polyffusion/ on sys.path
import polyffusion_b200
polyffusion_b200.install_dropin()

# inference.py:12-14, train/train_ddpm.py:2-3
from ddpm import DenoiseDiffusion
from ddpm.unet import UNet
from ddpm.utils import gather
# inference_sdf.py:37-41, train/train_ldm.py:9-10
from sampler_ddim import DDIMSampler
from sampler_sdf import SDFSampler
from stable_diffusion.latent_diffusion import LatentDiffusion
from stable_diffusion.model.unet import UNetModel
from stable_diffusion.sampler import DiffusionSampler
# models/model_autoencoder.py:5, train/train_autoencoder.py:8 -- NOT replaced: the reference's own file
from stable_diffusion.model.autoencoder import Autoencoder, Decoder, Encoder
# other reference submodules of the aliased packages
import stable_diffusion.sampler.ddim as ref_ddim
import stable_diffusion.sampler.ddpm as ref_ddpm
import stable_diffusion.losses
try:  # found through the alias package's __path__; its own third-party import (labml.experiment) is absent here
    import ddpm.sampling  # noqa: F401
except ImportError as e:
    assert "labml" in str(e), e

for cls in (DenoiseDiffusion, UNet, DDIMSampler, SDFSampler, LatentDiffusion, UNetModel, DiffusionSampler):
    assert cls.__module__.startswith("polyffusion_b200."), (cls, cls.__module__)
assert gather.__module__ == "polyffusion_b200.ddpm.utils"
for mod in (sys.modules["stable_diffusion.model.autoencoder"], ref_ddim, ref_ddpm, sys.modules["stable_diffusion.losses"]):
    assert mod.__file__.startswith({ref!r}), mod.__file__
# the reference's un-replaced samplers see the drop-in base classes through their relative imports
assert ref_ddim.DiffusionSampler is DiffusionSampler and ref_ddim.LatentDiffusion is LatentDiffusion
print("OK")
"""


def test_reference_import_sites_resolve_after_install_dropin():
    shim = os.path.join(ROOT, "oracle", "shim")
    code = SCRIPT.format(shim=shim, ref=reference_loader.REFERENCE_ROOT, root=ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
