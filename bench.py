#!/usr/bin/env python
"""bench.py -- throughput of the Polyffusion DDPM sampling hot path on B200 (driver contract).

Metric (BASELINE.json): 8-bar piano-roll samples/sec for 1000-step DDPM.  Workload at every N:
configs[1] = sdf_chd8bar conditional, batch 64 per GPU (weak scaling), uncond_scale 1.
One bench "step" = one reverse-diffusion step over the batch: one UNet evaluation
(pf_unet_forward, B=64) + the fused step epilogue (pf_sample_step_ddpm, RePaint branch active with
orig = mask = 0 exactly as inference_sdf.py:218-220,289-301 runs plain generation) + the two
torch.randn draws the reference makes per step.  Every step costs the same irrespective of t, so

    samples/sec = N_gpus * 64 / (1000 * seconds_per_step + seconds_for_the_final_all_gather)

`value` is timed with inputs resident in HBM; `e2e` runs the SAME step function with HOST (pinned)
buffers: H2D of x_t/cond, the step, D2H of x_{t-1}, every step.  `sustained` repeats the step back to
back for >= 3 s.  --config 2/3/4 time BASELINE.json configs[2..4] (DDIM-50, CFG scale 5, one window-step
of the song-batched autoregressive driver) on the same contract.
Inputs (8.4 MB x_t + 8.4 MB noise x2 + ~2.5 GB of activations) exceed nothing cache-wise: the
activation working set of one step is ~40x the 126 MB L2, so L2 is effectively flushed between
steps ("l2": "inputs larger than L2").

--impl reference times the CPU restatement of the reference (oracle/, the reference itself is
Python/PyTorch and is not present on the GPU box) with all host threads on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "8bar_pianoroll_samples_per_sec_ddpm1000"
UNIT = "samples/s"
B_PER_GPU = 64
DDPM_STEPS = 1000
D_COND = 512
WORKLOAD = "configs[1]: sdf_chd8bar conditional, batch=64 per GPU, 1000-step DDPM (SDFSampler.paint path), uncond_scale=1"
GFLOP_PER_SAMPLE_EVAL = 90.21  # SURVEY.md section 8(d), torch FlopCounterMode on the reference UNet


def sdf_kwargs():
    return dict(in_channels=2, out_channels=2, channels=64, n_res_blocks=2, attention_levels=[2, 3],
                channel_multipliers=[1, 2, 4, 4], n_heads=4, tf_layers=1, d_cond=D_COND)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback (B200_PROFILING.md)")


def load_gemm_traffic():
    """DRAM bytes moved by the tcgen05 GEMM launches of ONE step, from a committed ncu launch list
    (profiles/*_gemm_traffic.json, written by tools/ncu_launches_step.py): the file captured at the current
    kernel sources when there is one (file times mean nothing after a checkout), else the last by name
    (reported as stale by the caller); None if absent."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_gemm_traffic.json")))
    if not files:
        return None, None, None
    docs = []
    for path in files:
        with open(path) as f:
            docs.append((path, json.load(f)))
    cur = kernel_source_hash()
    match = [pd for pd in docs if pd[1].get("kernel_source_hash") == cur]
    path, d = (match or docs)[-1]
    return (d["gemm_dram_bytes_read"] + d["gemm_dram_bytes_write"], os.path.relpath(path, ROOT),
            d.get("kernel_source_hash"))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.lines = []
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU arms
def cpu_port_samples_per_sec(batch: int, steps: int, warmup: int):
    """Oracle restatement (== the reference's PyTorch CPU path, bit-identical on CPU) timed on all
    host cores: `steps` DDPM reverse steps at a bounded batch."""
    import torch

    from oracle.sampler_oracle import ddpm_p_sample, ddpm_tables, ldm_schedule
    from oracle.unet_oracle import UNetCfg, unet_forward
    from polyffusion_b200.stable_diffusion.model.unet import UNetModel

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = {k: v.detach() for k, v in UNetModel(**sdf_kwargs()).state_dict().items()}
    cfg = UNetCfg(d_cond=D_COND)
    _, beta, alpha_bar = ldm_schedule()
    tb = ddpm_tables(alpha_bar, beta)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(batch, 2, 128, 128, generator=g)
    cond = torch.randn(batch, 1, D_COND, generator=g)
    eps_fn = lambda xx, tt, cc: unet_forward(sd, cfg, xx, tt, cc)
    noise_fn = lambda shape: torch.randn(shape, generator=g)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        x, _, _ = ddpm_p_sample(tb, eps_fn, x, cond, 999 - i, noise_fn)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch / (DDPM_STEPS * sec), sec, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 4
    steps = max(1, min(args.steps, 8))
    warmup = max(1, min(args.warmup, 1))
    value, sec, cores = cpu_port_samples_per_sec(batch, steps, warmup)
    sample = f"{steps} DDPM reverse steps at batch {batch} (UNet eval + step), {warmup} warm-up, oracle port of the reference PyTorch CPU path"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "cpu_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
# --config N selects BASELINE.json configs[N]; configs[1] is the one the metric is quoted on (default).
CONFIGS = {
    1: dict(workload=WORKLOAD, metric=METRIC, unit=UNIT, d_cond=D_COND, kind="ddpm", scale=1.0, units_per_gpu=B_PER_GPU,
            steps_per_unit=DDPM_STEPS, evals_per_step=1),
    2: dict(workload="configs[2]: sdf_txt conditional, batch=64 per GPU, DDIM 50 steps eta=0.0 (DDIMSampler.sample path)",
            metric="8bar_pianoroll_samples_per_sec_ddim50", unit=UNIT, d_cond=1024, kind="ddim", scale=1.0,
            units_per_gpu=B_PER_GPU, steps_per_unit=50, evals_per_step=1),
    3: dict(workload="configs[3]: sdf_chd8bar uncond_scale=5 classifier-free guidance (2 UNet evals per step), "
                     "batch=64 per GPU (512 over 8 GPUs; UNet batch 128), 1000-step DDPM",
            metric="8bar_pianoroll_samples_per_sec_ddpm1000_cfg5", unit=UNIT, d_cond=D_COND, kind="ddpm", scale=5.0,
            units_per_gpu=B_PER_GPU, steps_per_unit=DDPM_STEPS, evals_per_step=2),
    4: dict(workload="configs[4]: autoregressive inpaint (inpaint_type=below), 10 segments per song = 19 windows, "
                     "32 songs per GPU (256 over 8 GPUs), 1000-step DDPM per window; one step = one (window, step) "
                     "over the 32 songs",
            metric="songs_per_sec_autoreg10_ddpm1000", unit="songs/s", d_cond=D_COND, kind="autoreg", scale=1.0,
            units_per_gpu=32, steps_per_unit=19 * DDPM_STEPS, evals_per_step=1),
}


def kernel_source_hash():
    """sha1 over the CUDA sources: the committed ncu traffic file records the hash it was measured at."""
    import glob
    import hashlib
    h = hashlib.sha1()
    for f in sorted(glob.glob(os.path.join(ROOT, "polyffusion_b200", "csrc", "*.cu*"))):
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: polyffusion_b200 has no CPU fallback "
                           "(use --impl reference for the CPU baseline)")
    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    # stdout must carry exactly ONE JSON line, but NCCL prints its version banner there (C-level
    # printf, NCCL_DEBUG_FILE does not move it): point fd 1 at stderr for the whole run and keep a
    # private handle on the real stdout for the result line
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from polyffusion_b200.sampler_ddim import DDIMSampler
    from polyffusion_b200.sampler_sdf import SDFSampler
    from polyffusion_b200.stable_diffusion.latent_diffusion import LatentDiffusion
    from polyffusion_b200.stable_diffusion.model.unet import UNetModel

    torch.manual_seed(0)  # identical weights on every rank
    kw = sdf_kwargs()
    kw["d_cond"] = cfg["d_cond"]
    unet = UNetModel(**kw).eval()
    ldm = LatentDiffusion(unet, None, 0.18215, DDPM_STEPS, 0.00085, 0.012).to(dev)
    B = cfg["units_per_gpu"]  # chains advanced by one step (samples; songs for configs[4])
    torch.manual_seed(1000 + rank)  # per-rank noise stream
    cond = torch.randn(B, 1, cfg["d_cond"], device=dev)
    uncond = -torch.ones(B, 1, cfg["d_cond"], device=dev) if cfg["scale"] != 1.0 else None
    orig = torch.zeros(B, 2, 128, 128, device=dev)
    mask = torch.zeros_like(orig)
    if cfg["kind"] == "autoreg":
        # one window of the song-batched autoregressive driver (autoreg.py): the first half of every
        # window is the known region generated by the previous window, "below" mask on the rest
        mask[:, :, :64, :] = 1.0
        mask[:, :, 64:, :60] = 1.0
        orig[:, 0, ::4, 72] = 1.0
    noise_mode = os.environ.get("PF_NOISE", "philox")
    if cfg["kind"] == "ddim":
        sampler = DDIMSampler(ldm, 50, "uniform", 0.0)
        x0 = torch.randn(B, 2, 128, 128, device=dev)
        n_idx = len(sampler.time_steps)

        def run_steps(xx, start, n):  # n DDIM steps (sampler_ddim.py:145-163) from schedule index `start`
            while n > 0:
                k = min(n, start % n_idx + 1)
                xx = sampler.advance(xx, cond, start % n_idx, k)
                start, n = start - k, n - k
            return xx
    else:
        sampler = SDFSampler(ldm)
        x0 = sampler.q_sample(orig, DDPM_STEPS - 1, torch.randn(B, 2, 128, 128, device=dev))

        def run_steps(xx, start, n):
            # n iterations of SDFSampler.paint's loop body (sampler_sdf.py:313-336): known-region noise, UNet
            # evaluation(s), CFG / x0 / mean / noise / RePaint-blend epilogue.  start stays >= n so that no
            # slice reaches step 0 (which draws no noise)
            # (--steps >= 1000 runs whole 1000-step chains, step 0 included)
            while n > 0:
                k = min(n, DDPM_STEPS)
                s = DDPM_STEPS - 1 if k == DDPM_STEPS else k + start % (DDPM_STEPS - k)
                xx = sampler.advance(xx, cond, s, k, orig=orig, mask=mask, uncond_scale=cfg["scale"],
                                     uncond_cond=uncond)
                n -= k
            return xx
    # in-kernel Philox noise keyed by the GLOBAL sample index: the sharded run reproduces the 1-GPU samples
    sampler.noise = noise_mode
    sampler.seed = 20261017
    sampler.sample0 = rank * B

    def step_fn(xx, i):
        return run_steps(xx, i, 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    x = x0
    step_id = DDPM_STEPS - 1
    x = run_steps(x, step_id, max(args.warmup, 3))
    step_id -= max(args.warmup, 3)
    launches_per_eval = unet.engine.launch_count()

    # ---- timed region: device-resident inputs
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    x = run_steps(x, step_id, args.steps)  # ONE call of the public sampler API: K graph replays, no host work
    step_id -= args.steps
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clock_info = clocks.stop() if rank == 0 else None

    # ---- sustained: the same step back to back for >= 3 s (what a 1000-step run actually sees once the
    # board has reached its power limit), with its own clock record
    sus_steps, sus_ms, sus_clock = 0, 0.0, None
    if args.sustain_seconds > 0:
        n_sus = max(args.steps, int(args.sustain_seconds * 1e3 / max(ms_total / args.steps, 1e-3)) + 1)
        sclk = ClockSampler(local_rank)
        if rank == 0:
            sclk.start()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        x = run_steps(x, step_id, n_sus)
        step_id -= n_sus
        s1.record()
        barrier()
        sus_steps, sus_ms = n_sus, s0.elapsed_time(s1)
        sus_clock = sclk.stop() if rank == 0 else None

    # ---- final all-gather of the finished samples (the path's only collective)
    gather_ms = 0.0
    if world > 1:
        outs = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(outs, x)  # warm-up (communicator setup)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather(outs, x)
        g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1)

    # ---- e2e: the SAME step through the public sampler API with HOST (pinned) buffers: x_t and cond go
    # host -> device and x_{t-1} comes back device -> host every step, inside the timed region
    x_host = x.detach().cpu().pin_memory()
    cond_host = cond.detach().cpu().pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    cond_dev = cond

    def e2e_step(i):
        nonlocal cond
        xd = x_host.to(dev, non_blocking=True)
        cond = cond_host.to(dev, non_blocking=True)
        out_host.copy_(step_fn(xd, i), non_blocking=True)

    for i in range(3):
        e2e_step(500 - i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_e2e = max(3, min(args.steps, 10))
    for i in range(n_e2e):
        e2e_step(490 - i)
    e1.record()
    barrier()
    cond = cond_dev
    e2e_ms = e0.elapsed_time(e1) / n_e2e

    # ---- per-kernel breakdown of one UNet evaluation (CUDA events around every launch of the plan)
    prof = {}
    Bp = B * cfg["evals_per_step"]
    xp = torch.randn(Bp, 2, 128, 128, device=dev)
    cp = torch.randn(Bp, 1, cfg["d_cond"], device=dev)
    ts = xp.new_full((Bp,), 500, dtype=torch.long)
    gemm_ms, gemm_flops, other_ms, gemm_launches, attn_ms, attn_flops = 0.0, 0.0, 0.0, 0, 0.0, 0.0
    n_prof = 3
    for _ in range(n_prof):
        unet.engine.forward(xp, ts, cp, profile=prof)
        for ms, fl, kd in zip(prof["ms"], prof["flops"], prof["kind"]):
            if kd == 0:  # gemm_tc_kernel
                gemm_ms += ms
                gemm_flops += fl
                gemm_launches += 1
            elif kd == 12:  # attn_tc_kernel
                attn_ms += ms
                attn_flops += fl
            else:
                other_ms += ms
    gemm_ms /= n_prof
    gemm_flops /= n_prof
    attn_ms /= n_prof
    attn_flops /= n_prof
    other_ms /= n_prof
    gemm_launches //= n_prof

    # ---- reduce over ranks (max time)
    vals = torch.tensor([ms_total, gather_ms, e2e_ms, sus_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    ms_total, gather_ms, e2e_ms, sus_ms = (float(v) for v in vals.tolist())
    ms_per_step = ms_total / args.steps
    spu = cfg["steps_per_unit"]
    rate = lambda ms: world * B / (spu * ms / 1e3 + gather_ms / 1e3)
    value, e2e_value = rate(ms_per_step), rate(e2e_ms)

    if rank == 0:
        peaks = load_peaks()
        traffic, traffic_src, traffic_hash = load_gemm_traffic()
        traffic_note = None
        if traffic is not None and traffic_hash != kernel_source_hash():
            traffic_note = (f"{traffic_src} was captured at kernel sources {traffic_hash}, current {kernel_source_hash()}: "
                            "stale, not reported (re-run tools/gpu_final.sh)")
            traffic = None
        elif traffic is not None:
            traffic_note = (f"dram__bytes_read.sum + dram__bytes_write.sum summed over the GEMM launches of one step "
                            f"({traffic_src}, same kernel sources)")
        achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        step_tflop = GFLOP_PER_SAMPLE_EVAL * Bp / 1e3
        cpu_value, cpu_sec, cores = (cpu_port_samples_per_sec(4, 3, 1)
                                     if world == 1 and not args.no_cpu and args.config == 1 else (None, None, None))
        line = {
            "metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": ("NON-PARITY fast mode: single fp16 / bf16 product per GEMM (attention still bf16x3); fp32 accumulate"
                      if os.environ.get("PF_FAST") == "1" else
                      "f16+f8 convolutions (fp16 product + e4m3 cross terms), bf16x3 linears/attention; fp32 accumulate"),
            "data": "synthetic",
            "config": {
                "workload": cfg["workload"], "batch_per_gpu": B, "global_batch": world * B,
                "parallelism": f"dp{world} (independent chains, one all-gather at the end)",
                "l2": "inputs larger than L2 (per-step activation working set >> 126 MB)",
                "allgather_ms": gather_ms,
                "unet_eval_algorithmic_tflop": step_tflop,
                "unet_algorithmic_tflops": step_tflop / (ms_per_step / 1e3),
                "precision_note": "convolutions: 1 fp16 MMA + 1 fp8 MMA at twice the rate per product (2 tensor-time "
                                  "units); linears / attention: 3 bf16 MMAs per product (hi*hi + lo*hi + hi*lo)",
                "conv_operands": os.environ.get("PF_CONV_F8_MAX_HW", "f16f8 everywhere (default)"),
                "noise": ("in-kernel Philox4x32-10 keyed by (seed, global sample index, element, step)"
                          if noise_mode == "philox" else "torch.randn per step in the reference's order"),
                "loop": "whole-step CUDA graph (UNet + fused step epilogue + device-side step counter)"
                        if getattr(sampler, "fused_loop", False) and cfg["scale"] == 1.0 else "per-step launches",
                "step_breakdown_ms": {"tcgen05_gemm": gemm_ms, "tcgen05_attention": attn_ms,
                                      "other_kernels": other_ms},
                "attention_algorithmic_tflops": (attn_flops / (attn_ms * 1e-3) / 1e12) if attn_ms > 0 else None,
            },
            "clocks": clock_info,
            "sustained": None if sus_steps == 0 else {
                "value": rate(sus_ms / sus_steps), "unit": cfg["unit"], "ms_per_step": sus_ms / sus_steps,
                "steps": sus_steps, "seconds": sus_ms / 1e3, "clocks": sus_clock,
                "note": "same step, back to back; the number a 1000-step run sees at the board power limit"},
            "e2e": {"value": e2e_value, "unit": cfg["unit"],
                    "h2d_bytes_per_step": x_host.numel() * 4 + cond_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": e2e_ms,
                    "note": "same step function as `value`, x_t / cond from pinned host memory, x_{t-1} back to the host"},
            # kernels of one step: the plan's launches (the step arithmetic lives in its last kernel) + the
            # device-side counter advance; guided (CFG) steps launch the separate step kernel instead
            "gpu_launches": args.steps * (launches_per_eval + 1),
            "roofline": {
                "bound": "tensor", "kernel": "gemm_tc*_kernel (all tcgen05 GEMM launches of one step)",
                "achieved": achieved_tf, "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                "frac": achieved_tf / peaks["tf_sust"], "traffic": traffic, "traffic_note": traffic_note,
                "launches_per_step": gemm_launches,
                "algorithmic_flops_per_step": gemm_flops, "kernel_ms_per_step": gemm_ms,
                "peak_source": "bf16 dense sustained, " + peaks["source"],
                "note": "achieved = algorithmic FLOP (2*M*N*K per product) / measured GEMM time; the split-precision "
                        "schemes issue 2 (f16f8) or 3 (bf16x3) tensor-time units per algorithmic product",
                # the whole step (GEMMs + attention + HBM-bound transforms + step epilogue) on the same scale
                "whole_step_frac": step_tflop / (ms_per_step / 1e3) / peaks["tf_sust"],
            },
        }
        if cpu_value is not None:
            line["cpu_baseline"] = {
                "value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "3 DDPM reverse steps at batch 4 (1 warm-up), oracle port of the reference PyTorch CPU path, all host threads",
            }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS),
                    help="BASELINE.json configs[N]; 1 (default) is the configuration the metric is quoted on")
    ap.add_argument("--sustain-seconds", type=float, default=3.0,
                    help="length of the extra back-to-back run reported under `sustained` (0 = skip)")
    ap.add_argument("--fast", action="store_true",
                    help="NON-PARITY single-pass tensor math (PF_FAST=1): every GEMM issues only its hi x hi product; "
                         "reported separately (BASELINE.md section 2), never the headline line")
    args = ap.parse_args()
    if args.fast:
        os.environ["PF_FAST"] = "1"  # read once by libpf_b200.so when the first plan is built
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
