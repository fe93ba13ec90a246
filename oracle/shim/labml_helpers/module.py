"""Stub of labml_helpers.module: the reference's legacy ddpm UNet subclasses this alias of nn.Module."""
import torch.nn as nn

Module = nn.Module
