"""Whole-step CUDA graph for the sampling loops (``pf_unet_forward_step``, include/pf_b200.h).

A reverse-diffusion step of the reference is one UNet evaluation followed by 31-37 ATen ops and 5
device->host syncs.  Here a step is ONE graph replay: the UNet plan with the
step arithmetic applied to eps inside its last kernel, the timestep embedding gathered from a table, the
cross-attention vectors hoisted out of the loop, and the step counter / timestep tensor advanced on the
device.  The host only replays the graph (and, in the reference-order noise mode, draws the two
``torch.randn`` tensors per step exactly where the reference draws them).

Noise modes
* ``"torch"`` (default, parity): ``torch.randn_like(orig)`` then ``torch.randn(x.shape)`` per step, in the
  reference's order (sampler_sdf.py:318, 157-160), copied into the graph's static buffers.
* ``"philox"``: drawn inside the kernel from Philox4x32-10 keyed by (seed, global sample index, element,
  step) -- no noise tensors, and the result does not depend on how the batch is sharded over ranks
  (SURVEY.md section 8e).
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict
from typing import Optional, Sequence

import torch

from ._lib import FusedStepArgs, check, current_stream, lib, ptr

_GOLDEN = 0x9E3779B97F4A7C15
_MASK64 = (1 << 64) - 1


def effective_seed(seed: int, nonce: int) -> int:
    """Philox key of run `nonce` (kernels.cu fused_step_apply): seed + nonce * 2^64/phi (mod 2^64)."""
    return (int(seed) + int(nonce) * _GOLDEN) & _MASK64


class FusedLoop:
    """Static buffers + one captured graph per (batch, n_cond, H, W, known-region, noise-mode) signature."""

    def __init__(self, unet, kind: int, coef_rows: Sequence[Sequence[float]], t_table: Sequence[int]):
        self.unet = unet
        self.kind = int(kind)
        self._coef_host = torch.tensor([list(r)[:7] + [0.0] * (8 - len(list(r)[:7])) for r in coef_rows],
                                       dtype=torch.float32)
        self._t_host = torch.tensor([int(v) for v in t_table], dtype=torch.int64)
        self._dev = {}
        self._graphs: "OrderedDict[tuple, dict]" = OrderedDict()
        self._nonce = 0

    def _tables(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (self._coef_host.to(device), self._t_host.to(device))
        return self._dev[key]

    def _state(self, key, x, cond, has_orig, injected, temperature, seed, sample0):
        st = self._graphs.get(key)
        if st is not None and st["generation"] == self.unet.engine.generation:
            self._graphs.move_to_end(key)
            return st
        dev = x.device
        coef, t_table = self._tables(dev)
        st = dict(
            x=torch.empty_like(x, dtype=torch.float32).contiguous(),
            cond=torch.empty_like(cond, dtype=torch.float32).contiguous(),
            t=torch.zeros(x.shape[0], dtype=torch.int64, device=dev),
            index=torch.zeros(2, dtype=torch.int32, device=dev),  # [index, run nonce]
            noise=torch.zeros_like(x, dtype=torch.float32) if injected else None,
            noise_kn=torch.zeros_like(x, dtype=torch.float32) if (injected and has_orig) else None,
            orig=torch.zeros_like(x, dtype=torch.float32) if has_orig else None,
            mask=torch.zeros_like(x, dtype=torch.float32) if has_orig else None,
            graph=None,
            generation=self.unet.engine.generation,
        )
        a = FusedStepArgs()
        a.kind, a.flags = self.kind, 1
        a.index, a.coef, a.t_table = st["index"].data_ptr(), coef.data_ptr(), t_table.data_ptr()
        a.x = st["x"].data_ptr()
        a.eps_out = None
        a.noise = st["noise"].data_ptr() if st["noise"] is not None else None
        a.noise_kn = st["noise_kn"].data_ptr() if st["noise_kn"] is not None else None
        a.orig = st["orig"].data_ptr() if has_orig else None
        a.mask = st["mask"].data_ptr() if has_orig else None
        a.temperature, a.seed, a.sample0 = float(temperature), int(seed) & _MASK64, int(sample0)
        st["args"] = a
        self._graphs[key] = st
        while len(self._graphs) > 4:
            self._graphs.popitem(last=False)
        return st

    def _launch(self, st, shape):
        eng = self.unet.engine
        B, n_cond, H, W = shape
        ws, base, nbytes = eng.workspace((B, n_cond, H, W), st["x"].device)
        check(lib().pf_unet_forward_step(eng.handle, ptr(st["x"]), ptr(st["t"]), ptr(st["cond"]), B, n_cond, H, W,
                                         ctypes.byref(st["args"]), ctypes.c_void_p(base), nbytes, current_stream()))

    @torch.no_grad()
    def run(self, x: torch.Tensor, cond: torch.Tensor, start_index: int, n_steps: int, *,
            orig: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
            noise_mode: str = "torch", fixed_noise_kn: Optional[torch.Tensor] = None, temperature: float = 1.0,
            seed: int = 0, sample0: int = 0, draw=None) -> torch.Tensor:
        """Advance x by n_steps indices start_index, start_index - 1, ...  `draw(index)` (torch mode) returns
        the (noise_kn, noise) tensors of that index in the reference's order (either may be None)."""
        if not x.is_cuda:
            raise RuntimeError("polyffusion_b200 samplers run on CUDA tensors only (no CPU fallback)")
        eng = self.unet.engine
        dev = x.device
        B, _, H, W = x.shape
        n_cond = cond.shape[1]
        has_orig = orig is not None
        injected = noise_mode == "torch"
        with torch.cuda.device(dev):
            eng.sync_weights(dev)
            eng.enable_time_lut(int(self._t_host.max().item()) + 1)
            key = (B, n_cond, H, W, has_orig, injected, float(temperature), int(seed), int(sample0))
            st = self._state(key, x, cond, has_orig, injected, temperature, seed, sample0)
            coef, t_table = self._tables(dev)
            st["x"].copy_(x)
            st["cond"].copy_(cond.to(dev))
            if has_orig:
                st["orig"].copy_(orig.expand_as(x))
                st["mask"].copy_(mask.expand_as(x))
                if fixed_noise_kn is not None and injected:
                    st["noise_kn"].copy_(fixed_noise_kn.expand_as(x))
            self._nonce += 1
            st["index"].copy_(torch.tensor([start_index, self._nonce & 0x7FFFFFFF], dtype=torch.int32), non_blocking=True)
            st["t"].fill_(int(self._t_host[start_index]))
            shape = (B, n_cond, H, W)
            ws, base, nbytes = eng.workspace(shape, dev)
            check(lib().pf_unet_prepare_cond(eng.handle, ptr(st["cond"]), B, n_cond, H, W, ctypes.c_void_p(base),
                                             nbytes, current_stream()))
            index = start_index
            for _ in range(n_steps):
                if injected and draw is not None:
                    nk, nz = draw(index)
                    if nk is not None and has_orig and fixed_noise_kn is None:
                        st["noise_kn"].copy_(nk)
                    if nz is not None:
                        st["noise"].copy_(nz.expand_as(x))
                if st["graph"] is None and eng.use_graph and not torch.cuda.is_current_stream_capturing():
                    # first step of this signature: one eager launch (builds the plan), then capture the
                    # launch sequence for every later step; the device-side state is restored in between
                    saved = (st["x"].clone(), st["index"].clone(), st["t"].clone())
                    self._launch(st, shape)
                    torch.cuda.current_stream().synchronize()
                    st["x"].copy_(saved[0]); st["index"].copy_(saved[1]); st["t"].copy_(saved[2])
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._launch(st, shape)
                    st["x"].copy_(saved[0]); st["index"].copy_(saved[1]); st["t"].copy_(saved[2])
                    st["graph"] = g
                if st["graph"] is not None:
                    st["graph"].replay()
                else:
                    self._launch(st, shape)
                index -= 1
            return st["x"].clone()

    def nonce(self) -> int:
        return self._nonce & 0x7FFFFFFF
