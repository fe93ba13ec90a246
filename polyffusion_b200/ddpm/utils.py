"""``ddpm.utils`` of the reference: ``gather`` (ddpm/utils.py:13-16)."""
import torch


def gather(consts: torch.Tensor, t: torch.Tensor):
    """consts[t] reshaped to broadcast over [B, C, H, W]."""
    return consts.gather(-1, t).reshape(-1, 1, 1, 1)
