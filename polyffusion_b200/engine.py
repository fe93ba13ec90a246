"""Host-side owner of one ``pf_unet`` handle (include/pf_b200.h) for a ``UNetModel`` on one GPU.

Lifetime: created lazily on the first CUDA forward; weights are pushed through
``pf_unet_set_weight`` / ``pf_unet_finalize`` (the library keeps its own packed split-bf16 copy) and
re-pushed whenever a parameter's storage or version counter changes (``load_state_dict``, ``.to()``,
optimizer steps).  Workspaces / CUDA graphs are cached per (batch, n_cond, H, W) in a small LRU
(PF_ENGINE_CACHE entries, default 4: a ragged last batch, the CFG 2B evaluation, the autoregressive
song batch), so a workload with many batch sizes cannot pile up activation workspaces.
"""
from __future__ import annotations

import ctypes
import math
import os
from collections import OrderedDict
from typing import Tuple

import torch

from ._lib import PfError, UNetCfg, check, current_stream, lib, ptr


class UNetEngine:
    def __init__(self, module: "torch.nn.Module", cfg: dict):
        self.module = module
        self.cfg = cfg
        self.handle = ctypes.c_void_p()
        self.device = None
        self._stamp = None
        self._workspaces: "OrderedDict[Tuple[int, int, int, int], torch.Tensor]" = OrderedDict()
        self._graphs: "OrderedDict[Tuple[int, int, int, int], tuple]" = OrderedDict()
        self._cache_entries = max(1, int(os.environ.get("PF_ENGINE_CACHE", "4")))
        self._lut_rows = 0
        self.generation = 0
        # replay each (batch, n_cond, H, W) evaluation as a CUDA graph (measured ~3 % faster than the
        # 261 individual launches); PF_CUDA_GRAPH=0 disables it
        self.use_graph = os.environ.get("PF_CUDA_GRAPH", "1") != "0"

    # ------------------------------------------------------------------ handle management
    def _create(self, device: torch.device) -> None:
        c = UNetCfg()
        c.in_channels = self.cfg["in_channels"]
        c.out_channels = self.cfg["out_channels"]
        c.channels = self.cfg["channels"]
        c.n_res_blocks = self.cfg["n_res_blocks"]
        mult = list(self.cfg["channel_multipliers"])
        if len(mult) > 8:
            raise PfError("at most 8 resolution levels are supported")
        c.n_levels = len(mult)
        for i, m in enumerate(mult):
            c.channel_multipliers[i] = int(m)
            c.attention_levels[i] = int(i in set(self.cfg["attention_levels"]))
        c.n_heads = self.cfg["n_heads"]
        c.tf_layers = self.cfg["tf_layers"]
        c.d_cond = self.cfg["d_cond"]
        with torch.cuda.device(device):
            check(lib().pf_unet_create(ctypes.byref(c), ctypes.byref(self.handle)))
        self.device = device

    def _params_stamp(self):
        return tuple((p.data_ptr(), p._version) for p in self.module.parameters())

    def _touch(self, key) -> None:
        """Mark `key` most recently used and evict the oldest (batch, n_cond, H, W) entries: graph first
        (it replays launches that point into the workspace), then the workspace."""
        for cache in (self._graphs, self._workspaces):
            if key in cache:
                cache.move_to_end(key)
        while len(self._workspaces) > self._cache_entries:
            old, _ = self._workspaces.popitem(last=False)
            self._graphs.pop(old, None)
            self.generation += 1  # a workspace went away: graphs captured over it are stale

    def sync_weights(self, device: torch.device, force: bool = False) -> None:
        """(Re)pack the module's current parameters into the library."""
        if self.handle.value is None or self.device != device:
            self.close()
            self._create(device)
            force = True
        stamp = self._params_stamp()
        if not force and stamp == self._stamp:
            return
        keep = []
        with torch.cuda.device(device):
            for name, p in self.module.state_dict().items():
                t = p.detach().to(device=device, dtype=torch.float32).contiguous()
                keep.append(t)
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                check(lib().pf_unet_set_weight(self.handle, name.encode(), ptr(t), shape, t.dim()))
            # frequency table of the sinusoidal timestep embedding, evaluated exactly as
            # stable_diffusion/model/unet.py:158-164 does (fp32 torch ops on the host)
            half = self.cfg["channels"] // 2
            freqs = torch.exp(
                -math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half
            ).to(device)
            keep.append(freqs)
            shape = (ctypes.c_int64 * 1)(half)
            check(lib().pf_unet_set_weight(self.handle, b"__time_freqs", ptr(freqs), shape, 1))
            check(lib().pf_unet_finalize(self.handle, current_stream()))
        del keep
        self._stamp = stamp
        self._lut_rows = 0  # pf_unet_finalize dropped the time-embedding table
        self.generation += 1  # plans were rebuilt: graphs captured by callers (FusedLoop) are stale
        self._workspaces.clear()
        self._graphs.clear()

    def close(self) -> None:
        if self.handle.value is not None:
            lib().pf_unet_destroy(self.handle)
            self.handle = ctypes.c_void_p()
        self._workspaces.clear()
        self._graphs.clear()
        self._stamp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, t: torch.Tensor, cond: torch.Tensor, out=None, profile=None) -> torch.Tensor:
        """``profile``: optional dict; when given, the plan runs once with CUDA events around every
        launch and the dict receives ``ms`` / ``flops`` / ``kind`` lists (pf_unet_forward_profiled)."""
        if not x.is_cuda:
            raise PfError("polyffusion_b200.UNetModel runs on CUDA tensors only (no CPU fallback)")
        dev = x.device
        self.sync_weights(dev)
        B, Cin, H, W = x.shape
        if Cin != self.cfg["in_channels"]:
            raise PfError(f"expected {self.cfg['in_channels']} input channels, got {Cin}")
        if cond.dim() != 3 or cond.shape[0] != B or cond.shape[2] != self.cfg["d_cond"]:
            raise PfError(f"cond must be [B, n_cond, {self.cfg['d_cond']}], got {tuple(cond.shape)}")
        if t.shape != (B,):
            raise PfError(f"time_steps must be [B], got {tuple(t.shape)}")
        n_cond = cond.shape[1]
        x = x.contiguous().float()
        cond = cond.to(dev).contiguous().float()
        t = t.to(device=dev, dtype=torch.int64).contiguous()
        key = (B, n_cond, H, W)
        if self.use_graph and profile is None and not torch.cuda.is_current_stream_capturing():
            return self._forward_graph(key, x, t, cond, out)
        return self._forward_eager(key, x, t, cond, out, profile)

    def _forward_graph(self, key, x, t, cond, out):
        self._touch(key)
        entry = self._graphs.get(key)
        if entry is None:
            # static buffers + one eager call (builds the plan) + capture
            xs, ts, cs = x.clone(), t.clone(), cond.clone()
            os_ = torch.empty((key[0], self.cfg["out_channels"], key[2], key[3]), dtype=torch.float32,
                              device=x.device)
            self._forward_eager(key, xs, ts, cs, os_, None)
            torch.cuda.current_stream().synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._forward_eager(key, xs, ts, cs, os_, None)
            entry = (graph, xs, ts, cs, os_)
            self._graphs[key] = entry
        graph, xs, ts, cs, os_ = entry
        xs.copy_(x)
        ts.copy_(t)
        cs.copy_(cond)
        graph.replay()
        if out is None:
            return os_.clone()
        out.copy_(os_)
        return out

    def workspace(self, key, dev):
        """(tensor, 1024-aligned base address, usable bytes) of the plan workspace for (B, n_cond, H, W)."""
        B, n_cond, H, W = key
        with torch.cuda.device(dev):
            ws = self._workspaces.get(key)
            if ws is not None:
                self._workspaces.move_to_end(key)
            if ws is None:
                nbytes = lib().pf_unet_workspace_bytes(self.handle, B, n_cond, H, W)
                if nbytes == 0:
                    raise PfError(lib().pf_last_error().decode("utf-8", "replace"))
                ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
                self._workspaces[key] = ws
                self._touch(key)
        base = (ws.data_ptr() + 1023) // 1024 * 1024
        return ws, base, ws.numel() - (base - ws.data_ptr())

    def enable_time_lut(self, n_steps: int) -> None:
        """Tabulate the time-embedding path for t = 0 .. n_steps-1 (pf_unet_enable_time_lut); samplers call
        this, a bare UNetModel.forward keeps evaluating it.  Plans and graphs are rebuilt against the table."""
        if getattr(self, "_lut_rows", 0) == n_steps:
            return
        check(lib().pf_unet_enable_time_lut(self.handle, int(n_steps), current_stream()))
        self._lut_rows = n_steps
        self.generation += 1
        self._graphs.clear()

    def _forward_eager(self, key, x, t, cond, out, profile):
        B, n_cond, H, W = key
        dev = x.device
        with torch.cuda.device(dev):
            ws, base, _ = self.workspace(key, dev)
            if out is None:
                out = torch.empty((B, self.cfg["out_channels"], H, W), dtype=torch.float32, device=dev)
            if profile is None:
                check(lib().pf_unet_forward(self.handle, ptr(x), ptr(t), ptr(cond), B, n_cond, H, W,
                                            ptr(out), ctypes.c_void_p(base),
                                            ws.numel() - (base - ws.data_ptr()), current_stream()))
            else:
                cap = 4096
                ms = (ctypes.c_float * cap)()
                fl = (ctypes.c_double * cap)()
                kd = (ctypes.c_int32 * cap)()
                n = ctypes.c_int32(0)
                check(lib().pf_unet_forward_profiled(
                    self.handle, ptr(x), ptr(t), ptr(cond), B, n_cond, H, W, ptr(out),
                    ctypes.c_void_p(base), ws.numel() - (base - ws.data_ptr()), current_stream(),
                    ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(fl, ctypes.c_void_p),
                    ctypes.cast(kd, ctypes.c_void_p), cap, ctypes.cast(ctypes.byref(n), ctypes.c_void_p)))
                profile["ms"] = list(ms[: n.value])
                profile["flops"] = list(fl[: n.value])
                profile["kind"] = list(kd[: n.value])
        return out

    def op_descriptions(self):
        out = []
        buf = ctypes.create_string_buffer(256)
        for i in range(self.launch_count()):
            check(lib().pf_unet_op_desc(self.handle, i, buf, 256))
            out.append(buf.value.decode())
        return out

    def launch_count(self) -> int:
        return int(lib().pf_unet_launch_count(self.handle)) if self.handle.value is not None else 0
