"""Condition encoders on the GPU path: drop-ins for the reference's
``dl_modules.chord_enc.RnnEncoder`` and ``dl_modules.txt_enc.TextureEncoder`` (same constructors,
parameter tree / ``state_dict`` keys and default-init RNG stream), evaluated by libpf_b200 kernels."""
from .chord_enc import RnnEncoder
from .txt_enc import TextureEncoder

ChordEncoder = RnnEncoder  # the reference imports it under this name (utils.py:8)

__all__ = ["RnnEncoder", "ChordEncoder", "TextureEncoder"]
