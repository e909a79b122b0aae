#!/usr/bin/env python
"""Benchmark of the hot path: DPM++2S-ancestral sampler steps / second on the reference's headline workload
(BASELINE.json configs[1]: tv2v depth-ControlNet, 17 keyframes at 512x768 -> latent 64x96, CFG 7.5, 30-step schedule).

One "step" = DPMPP2SAncestralSampler.sampler_step = TWO network calls (OpenAIWrapperControlLDM3DTV2V.forward at CFG
batch 2: ControlNet2D + ControlledUNetModel3DTV2V) + DiscreteDenoiser / CFG / sampler elementwise math, on one clip.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port) on the host cores

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM; `e2e`: the same step through the public wrapper with
pinned HOST buffers copied in and the result copied out every step; `roofline`: the dominant kernel, timed per launch
with CUDA events in an un-graphed profiling pass; `cpu_baseline`: the oracle port on a bounded sample (N=1, rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "UNet denoise-steps/sec (17x512x768 fp16, CFG on)"
UNIT = "sampler steps/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], burst=p["bf16_tflops"], sustained=p["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, source="fallback")


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region (B200_PROFILING.md "clocks" line)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(self.NAMES, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------------------
def synthetic_clip(kind, T, h, w, seed):
    """Host-side synthetic inputs of one clip: latent, cond / uncond dicts (hint shared)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, 4, T, h, w, generator=g)
    c = {"crossattn": torch.randn(1, 77, 768, generator=g)}
    uc = {"crossattn": torch.randn(1, 77, 768, generator=g)}
    c["control_hint"] = torch.rand(1, 3, T, 8 * h, 8 * w, generator=g) * 2 - 1
    uc["control_hint"] = c["control_hint"].clone()
    if kind == "tvi2v":
        c["cond_feat"] = torch.randn(1, 4, h, w, generator=g)
        uc["cond_feat"] = c["cond_feat"].clone()
    return x, c, uc


def workload_config(args):
    """`config` of the JSON line: the workload only (identical for `--impl ours` and `--impl reference`);
    implementation details of an arm go under its `impl_notes` key."""
    kind, T, h, w = args.kind, args.frames, args.height // 8, args.width // 8
    return {
        "workload": f"{kind} depth-ControlNet, {T} keyframes {args.height}x{args.width} (latent {h}x{w}), "
                    f"DPM++2S-ancestral {args.sampler_steps}-step schedule, cfg {args.cfg_scale}, "
                    f"{args.clips_per_gpu} clip(s) per GPU; 1 step = 2 network calls at CFG batch {2 * args.clips_per_gpu}",
        "kind": kind, "frames": T, "height": args.height, "width": args.width, "cfg_scale": args.cfg_scale,
        "sampler": "DPMPP2SAncestral", "sampler_steps": args.sampler_steps, "clips_per_gpu": args.clips_per_gpu,
    }


class Workload:
    """One rank's share of a workload: `clips` clips of (kind, T, h, w) run as ONE batch through the sampler step
    (CFG batch 2 * clips), with pinned host copies of every input for the end-to-end variant.

    fused=True (default): ccedit_b200.sampling.FusedDPMPP2SAncestralSampler - per-step elementwise math in 3 CUDA kernels,
    the whole step (both network calls) replayed from one CUDA graph, CFG de-duplication of the layers ahead of the
    first text cross-attention.  fused=False: the PyTorch-elementwise DPMPP2SAncestralSampler.sampler_step with the
    wrapper's per-call graph (round-1 behaviour)."""

    def __init__(self, wrap, dev, kind, clips, T, h, w, sampler_steps, cfg_scale, seed, fused=True, graph=True, dedup=True):
        from ccedit_b200.sampling import (BoundDenoiser, DiscreteDenoiser, DPMPP2SAncestralSampler,
                                          FusedDPMPP2SAncestralSampler)
        self.wrap, self.dev, self.fused = wrap, dev, fused
        den = DiscreteDenoiser().to(dev)
        gcfg = {"target": "sgm.modules.diffusionmodules.guiders.VanillaCFGTV2V", "params": {"scale": cfg_scale}}
        if fused:
            self.sampler = FusedDPMPP2SAncestralSampler(num_steps=sampler_steps, device=dev, eta=1.0, s_noise=1.0,
                                                        guider_config=gcfg, use_cuda_graph=graph, cfg_dedup=dedup)
        else:
            self.sampler = DPMPP2SAncestralSampler(num_steps=sampler_steps, device=dev, eta=1.0, s_noise=1.0, guider_config=gcfg)
        self.sigmas = self.sampler.discretization(sampler_steps, device=dev)
        parts = [synthetic_clip(kind, T, h, w, seed=seed + 1000 * i) for i in range(clips)]
        cat = lambda ts: torch.cat(ts, 0).pin_memory()
        self.x_h = cat([p[0] for p in parts])
        self.c_h = {k: cat([p[1][k] for p in parts]) for k in parts[0][1]}
        self.uc_h = {k: cat([p[2][k] for p in parts]) for k in parts[0][2]}
        self.out_h = torch.empty_like(self.x_h).pin_memory()
        self.x_d, self.c_d, self.uc_d = self.x_h.to(dev), self._to_dev(self.c_h), self._to_dev(self.uc_h)
        self.denoiser = BoundDenoiser(den, wrap)
        self.s_in = torch.ones(clips, device=dev)
        self.n_sched = sampler_steps - 1                 # every step but the last makes 2 network calls
        self.x0_scale = torch.sqrt(1.0 + self.sigmas[0] ** 2)
        self.h2d = sum(t.numel() * t.element_size() for t in [self.x_h, *self.c_h.values(), *self.uc_h.values()])
        self.d2h = self.out_h.numel() * self.out_h.element_size()
        if fused:
            self.plan = self.sampler.begin(self.denoiser, self.x_d.clone(), self.c_d, self.uc_d)
            self.x_start = self.plan["x"].clone()
            self.dedup = bool(self.plan["dedup"])
        else:
            self.state = self.x_d * self.x0_scale
            self.dedup = False

    def _to_dev(self, d):
        return {k: v.to(self.dev, non_blocking=True) for k, v in d.items()}

    def _step(self, i, x, c, uc):
        j = i % self.n_sched
        return self.sampler.sampler_step(self.s_in * self.sigmas[j], self.s_in * self.sigmas[j + 1], self.denoiser, x, c, uc)

    def resident_step(self, i):
        if self.fused:
            self.sampler.fused_step(self.plan, i % self.n_sched)
            if (i + 1) % self.n_sched == 0:              # restart the schedule so values stay in range
                self.plan["x"].copy_(self.x_start)
            return
        self.state = self._step(i, self.state, self.c_d, self.uc_d)
        if (i + 1) % self.n_sched == 0:
            self.state = self.x_d * self.x0_scale

    def e2e_step(self, i):
        """Host inputs in, host result out, every step: H2D of the latent and of both conditioning dicts from pinned
        memory, one sampler step through the public API, D2H of the new latent, host synchronisation."""
        if self.fused:
            self.sampler.load_inputs(self.plan, x=self.x_h, cond=self.c_h, uc=self.uc_h)
            if i % self.n_sched == 0:
                self.plan["x"].mul_(self.x0_scale)
            self.sampler.fused_step(self.plan, i % self.n_sched)
            self.out_h.copy_(self.plan["x"], non_blocking=True)
        else:
            x = self.x_h.to(self.dev, non_blocking=True)
            c, uc = self._to_dev(self.c_h), self._to_dev(self.uc_h)
            y = self._step(i, x * self.x0_scale if i % self.n_sched == 0 else x, c, uc)
            self.out_h.copy_(y, non_blocking=True)
        torch.cuda.current_stream().synchronize()        # the caller reads the result on the host every step


def run_ours(args):
    from ccedit_b200 import ops, parallel
    from ccedit_b200.census import network_flops
    from ccedit_b200.configs import build_network

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the ccedit_b200 path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    rank, local, world = parallel.init_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    kind, T, h, w = args.kind, args.frames, args.height // 8, args.width // 8
    clips = args.clips_per_gpu

    # weights: random init of the architecture on rank 0, broadcast once (the job's only collective)
    t0 = time.time()
    wrap = build_network(kind, device=dev, use_cuda_graph=not args.no_graph, randomize_zero_init_seed=1 if rank == 0 else None)
    bcast = parallel.broadcast_weights(wrap, src=0)
    n_params = sum(p.numel() for p in wrap.parameters())
    build_s = time.time() - t0

    def timed(fn, k):
        parallel.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = ops.launch_count()
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        parallel.barrier()
        return parallel.max_over_ranks(e0.elapsed_time(e1), dev), ops.launch_count() - n0

    # `clips` clips per rank (clip-parallel, weak scaling): same shape, rank-specific seed
    fused = args.sampler == "fused"
    wl = Workload(wrap, dev, kind, clips, T, h, w, args.sampler_steps, args.cfg_scale, seed=100 + rank, fused=fused,
                  graph=not args.no_graph, dedup=not args.no_dedup)
    warm = max(args.warmup, 3)
    for i in range(warm):                               # W >= 3 untimed warm-up steps (captures the CUDA graph)
        wl.resident_step(i)
    with ClockSampler(local) as clk:
        ms, launches = timed(wl.resident_step, args.steps)
    clocks = clk.summary()
    for i in range(2):
        wl.e2e_step(i)
    ms_e2e, _ = timed(wl.e2e_step, args.steps)

    ms_per_step = ms / args.steps
    value = world * clips * args.steps / (ms / 1e3)
    flops_call = network_flops(kind, 2 * clips, T, h, w)
    calls_per_step = 2
    step_tflops = flops_call["total"] * calls_per_step / (ms_per_step / 1e3) / 1e12

    result = {
        "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": workload_config(args),
        "impl_notes": {
            "parallelism": f"clip-parallel x{world} (weights broadcast once, no step-loop collective)",
            "weights": f"random init, {n_params / 1e6:.1f} M params, fp16 kernel copies",
            "cuda_graph": not args.no_graph,
            "sampler": ("FusedDPMPP2SAncestralSampler: 3 elementwise kernels + ONE CUDA graph per sampler step (both "
                        "network calls inside)" if fused else "DPMPP2SAncestralSampler (PyTorch elementwise), one CUDA "
                        "graph per network call"),
            "cfg_dedup": wl.dedup,
            "cfg_dedup_note": "the uncond and cond halves of the CFG batch share x, t and the hint: the layers ahead of "
                              "the first text cross-attention (ControlNet hint stem + input blocks 0-1, UNet input blocks "
                              "0-1 up to the self-attention) are computed once per call and fanned out; no result is "
                              "cached across calls; network_call.executed_tflop counts what actually ran",
            "l2": "no explicit flush: every step streams 3.2 GB of weights and >100 MB activations per layer, "
                  "far beyond the 126 MB L2",
            "value_counts": "clip-steps per second: every rank advances `clips_per_gpu` clips by one sampler step per step",
        },
        "e2e": {"value": round(world * clips * args.steps / (ms_e2e / 1e3), 4), "unit": UNIT, "h2d_bytes_per_step": wl.h2d,
                "d2h_bytes_per_step": wl.d2h, "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "network_call": {"ms": round(ms_per_step / calls_per_step, 3), "algorithmic_tflop": round(flops_call["total"] / 1e12, 3),
                         "achieved_tflops": round(step_tflops, 1),
                         "frac_of_measured_sustained_bf16": round(step_tflops / peaks["sustained"], 4)},
        "init": {"build_s": round(build_s, 1), "weights_broadcast_bytes": int(bcast)},
    }

    # ---- per-kernel breakdown + roofline of the dominant kernel: one un-graphed network call, events per launch ----
    if rank == 0 and not args.no_breakdown:
        bd = kernel_breakdown(wrap, wl.x_d[:1], {k: v[:1] for k, v in wl.c_d.items()}, {k: v[:1] for k, v in wl.uc_d.items()},
                              peaks, dev, dedup=wl.dedup)
        result["network_call"].update(bd.pop("_flops"))
        result.update(bd)
    # ---- the other BASELINE.json configs, a few steps each (single GPU only: the scaling runs stay short) ----
    if rank == 0 and world == 1 and not args.no_configs and kind == "tv2v" and clips == 1:
        del wl
        torch.cuda.empty_cache()
        result["variants"] = variants(args, wrap, dev, timed, fused)
        result["configs"] = other_configs(args, wrap, dev, peaks, timed)
    if rank == 0 and world == 1 and not args.no_configs:
        torch.cuda.empty_cache()
        try:
            result["library_baseline"] = library_gpu_baseline(kind, T, h, w, dev)
        except Exception as e:                              # a baseline must never take the bench line down
            result["library_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline(kind, T, h, w, budget_s=args.cpu_budget, calls=1, warm=0)
    if rank == 0:
        print(json.dumps(result), flush=True)
    parallel.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


def variants(args, wrap, dev, timed, fused):
    """The headline workload with the optimisations of the sampler path switched off one at a time (a few steps each),
    so that the headline can be read against the plain round-1 execution: PyTorch-elementwise sampler with one CUDA
    graph per network call, and the fused step without CFG de-duplication."""
    T, h, w = args.frames, args.height // 8, args.width // 8
    out = {}
    k = max(3, min(args.steps, 5))
    for name, kw in (("plain_sampler_per_call_graph", dict(fused=False)), ("fused_step_no_cfg_dedup", dict(fused=True, dedup=False))):
        if hasattr(wrap, "reset_graphs"):
            wrap.reset_graphs()
        torch.cuda.empty_cache()
        wl = Workload(wrap, dev, args.kind, 1, T, h, w, args.sampler_steps, args.cfg_scale, seed=100, **kw)
        for i in range(3):
            wl.resident_step(i)
        ms, _ = timed(wl.resident_step, k)
        out[name] = {"steps": k, "ms_per_step": round(ms / k, 3), "value": round(k / (ms / 1e3), 4), "unit": UNIT}
        del wl
    return out


def other_configs(args, wrap, dev, peaks, timed):
    """BASELINE.json configs[2..4] next to the headline (configs[1]), each timed for a few steps on this GPU:
    config 4's per-GPU share (2 clips per GPU as one CFG batch of 4), config 3 (tvi2v, 50-step schedule, cfg 7) and the
    config-5 sweep ({9,17,33} keyframes x {384x576, 512x768, 768x1024}: ms per network call, achieved TFLOP/s and the
    fraction of the measured sustained bf16 peak for every shape)."""
    from ccedit_b200.census import network_flops
    from ccedit_b200.configs import build_network
    out = {}
    k = max(3, min(args.steps, 5))

    def steps_entry(wl, kind, clips, T, h, w, note):
        for i in range(3):
            wl.resident_step(i)
        ms, _ = timed(wl.resident_step, k)
        for i in range(2):
            wl.e2e_step(i)
        ms_e, _ = timed(wl.e2e_step, k)
        fl = network_flops(kind, 2 * clips, T, h, w)["total"]
        tf = 2 * fl / (ms / k / 1e3) / 1e12
        return {"workload": note, "steps": k, "ms_per_step": round(ms / k, 3), "value": round(clips * k / (ms / 1e3), 4),
                "e2e_value": round(clips * k / (ms_e / 1e3), 4), "unit": UNIT, "algorithmic_tflop_per_call": round(fl / 1e12, 2),
                "achieved_tflops": round(tf, 1), "frac_of_measured_sustained_bf16": round(tf / peaks["sustained"], 4)}

    T, h, w = args.frames, args.height // 8, args.width // 8
    # config 4: 16 samples over 8 GPUs = 2 clips per GPU, run as one batch (CFG batch 4)
    wrap.reset_graphs()
    wl = Workload(wrap, dev, "tv2v", 2, T, h, w, 30, 7.5, seed=300)
    out["config4_per_gpu_share"] = steps_entry(wl, "tv2v", 2, T, h, w,
                                               "tv2v 17x512x768, 2 clips per GPU as one batch (CFG batch 4), 30-step schedule, cfg 7.5")
    del wl
    # config 5: resolution / length sweep, one network call (CFG batch 2) per shape
    sweep = []
    x1 = None
    for Ts in (9, 17, 33):
        for (H, W) in ((384, 576), (512, 768), (768, 1024)):
            wrap.reset_graphs()
            torch.cuda.empty_cache()
            hs, ws = H // 8, W // 8
            x, c, uc = synthetic_clip("tv2v", Ts, hs, ws, seed=500)
            cc = {kk: torch.cat((uc[kk], c[kk]), 0).to(dev) for kk in c}
            x2 = torch.cat([x] * 2).to(dev)
            t2 = torch.full((2,), 500, dtype=torch.long, device=dev)
            for _ in range(3):
                wrap(x2, t2, cc)
            ms, _ = timed(lambda i: wrap(x2, t2, cc), 3)
            fl = network_flops("tv2v", 2, Ts, hs, ws)["total"]
            tf = fl / (ms / 3 / 1e3) / 1e12
            sweep.append({"frames": Ts, "height": H, "width": W, "ms_per_network_call": round(ms / 3, 3),
                          "algorithmic_tflop": round(fl / 1e12, 2), "achieved_tflops": round(tf, 1),
                          "frac_of_measured_sustained_bf16": round(tf / peaks["sustained"], 4),
                          "steps_per_s": round(1e3 / (2 * ms / 3), 3)})
            del x2, cc
    out["config5_sweep"] = sweep
    # config 3: tvi2v (cfca center_self + controlnet_img), 50-step schedule, cfg 7
    wrap.reset_graphs()
    torch.cuda.empty_cache()
    wrap3 = build_network("tvi2v", device=dev, use_cuda_graph=not args.no_graph, randomize_zero_init_seed=1)
    wl = Workload(wrap3, dev, "tvi2v", 1, T, h, w, 50, 7.0, seed=400)
    out["config3_tvi2v"] = steps_entry(wl, "tvi2v", 1, T, h, w,
                                       "tvi2v ref-branch (cfca center_self) depthzoe, 17x512x768, 50-step schedule, cfg 7")
    del wl, wrap3
    torch.cuda.empty_cache()
    # SURVEY 8 row f1 (next after the UNet): first-stage decode of the clip, once per clip after the last sampler step
    from ccedit_b200.autoencoder import AutoencoderKLInferenceWrapper
    from ccedit_b200.census import decoder_flops
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    with torch.device(dev):
        vae = AutoencoderKLInferenceWrapper(ddconfig=dd, embed_dim=4).eval()
    z = torch.randn(1, 4, T, h, w, generator=torch.Generator().manual_seed(9)).to(dev)
    for _ in range(2):
        vae.decode(z, scale=1.0 / 0.18215)
    ms, _ = timed(lambda i: vae.decode(z, scale=1.0 / 0.18215), 3)
    fl = decoder_flops(T, h, w)["total"]
    tf = fl / (ms / 3 / 1e3) / 1e12
    out["first_stage_decode"] = {"workload": f"AutoencoderKL.decode of {T} frames {args.height}x{args.width} (latent {h}x{w}), "
                                             "fp16 storage / fp32 accumulation, eager launches",
                                 "ms_per_clip": round(ms / 3, 3), "algorithmic_tflop": round(fl / 1e12, 2),
                                 "achieved_tflops": round(tf, 1), "frac_of_measured_sustained_bf16": round(tf / peaks["sustained"], 4)}
    return out


def ncu_traffic(kernel):
    """Average DRAM bytes (read + write) per launch of `kernel`, from the committed ncu capture of one network call at
    this workload (profiles/traffic.json, written by tools/ncu_traffic.py from `ncu --metrics dram__bytes_read.sum,
    dram__bytes_write.sum`); None if no capture is committed for this kernel."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f)
    ent = t.get("kernels", {}).get(kernel)
    return None if ent is None else {"bytes_per_launch": ent["bytes_per_launch"], "launches": ent["launches"],
                                     "algorithmic_bytes_per_launch": ent.get("algorithmic_bytes_per_launch"),
                                     "source": t.get("source")}


def kernel_breakdown(wrap, x_d, c_d, uc_d, peaks, dev, dedup=False):
    from ccedit_b200 import ops
    cc = {k: torch.cat((uc_d[k], c_d[k]), 0) for k in c_d}
    x2 = torch.cat([x_d] * 2)
    t2 = torch.full((2,), 500, dtype=torch.long, device=dev)
    call = (lambda: wrap.forward_cfg(x_d, t2[:1], cc)) if dedup else (lambda: wrap(x2, t2, cc))
    graph, wrap.use_cuda_graph = wrap.use_cuda_graph, False
    try:
        call()                                           # warm
        torch.cuda.synchronize()
        ops.profile_start()
        call()
        recs = ops.profile_stop(executed=True)
    finally:
        wrap.use_cuda_graph = graph
    agg = {}
    for name, fl, by, ms, xf in recs:
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += fl
        a[2] += by
        a[3] += ms
        a[4] += xf
    total_ms = sum(a[3] for a in agg.values())
    rows = []
    for name, (n, fl, by, ms, xf) in sorted(agg.items(), key=lambda kv: -kv[1][3]):
        rows.append({"kernel": name, "launches": n, "ms": round(ms, 3), "share": round(ms / total_ms, 4),
                     "tflops": round(fl / ms / 1e9, 1) if fl else None, "gbs": round(by / ms / 1e6, 1)})
    # FLOPs of the call three ways: the reference's algorithmic count is the roofline numerator
    # (network_call.algorithmic_tflop); `launched` = useful FLOPs of the kernels this build launches (smaller: the text
    # K/V are projected once per batch entry instead of once per frame, the time-embedding rows once per call);
    # `executed` = what the tensor cores actually multiply, padding included (K padded to 64, head dim 40 -> 48,
    # partial 128-row / 128-key tiles).
    flops3 = {"launched_tflop": round(sum(a[1] for a in agg.values()) / 1e12, 3),
              "executed_tflop": round(sum(a[4] for a in agg.values()) / 1e12, 3)}
    # dominant kernel family = the tcgen05 tap-GEMM (all gemm.* classes are the same kernel) or attention
    fam = {}
    for r in rows:
        k = "tap_gemm_kernel" if r["kernel"].startswith("gemm.") else r["kernel"]
        f = fam.setdefault(k, [0, 0.0, 0.0])
        f[0] += r["launches"]
        f[1] += agg[r["kernel"]][1]
        f[2] += agg[r["kernel"]][3]
    families = [{"kernel": k, "launches": v[0], "ms": round(v[2], 3), "share": round(v[2] / total_ms, 4),
                 "tflops": round(v[1] / v[2] / 1e9, 1) if v[1] else None} for k, v in sorted(fam.items(), key=lambda kv: -kv[1][2])]
    top, (n, fl, ms) = max(fam.items(), key=lambda kv: kv[1][2])
    top_bytes = sum(agg[r["kernel"]][2] for r in rows if (r["kernel"].startswith("gemm.") and top == "tap_gemm_kernel")
                    or r["kernel"] == top)
    tensor_bound = fl > 0
    if tensor_bound:
        achieved = fl / ms / 1e9
        roof = {"kernel": top, "bound": "tensor", "achieved": round(achieved, 1), "peak": peaks["sustained"], "unit": "TFLOP/s",
                "frac": round(achieved / peaks["sustained"], 4)}
    else:
        by = sum(agg[r["kernel"]][2] for r in rows if r["kernel"] == top)
        achieved = by / ms / 1e6
        roof = {"kernel": top, "bound": "hbm", "achieved": round(achieved, 1), "peak": peaks["hbm"], "unit": "GB/s",
                "frac": round(achieved / peaks["hbm"], 4)}
    traffic = ncu_traffic(top)
    if traffic is not None and traffic.get("algorithmic_bytes_per_launch") is None:
        traffic["algorithmic_bytes_per_launch"] = round(top_bytes / n)
    roof.update({"traffic": traffic, "algorithmic_bytes_per_launch": round(top_bytes / n),
                 "algorithmic_flop_per_launch": round(fl / n), "launches_per_network_call": n, "avg_launch_ms": round(ms / n, 4),
                 "share_of_network_call": round(ms / total_ms, 4),
                 "peak_source": f"MEASURED_PEAKS.json ({peaks['source']}; sustained figure: kernel timed inside a full "
                                "network call)",
                 "how": "CUDA events around every launch on the launching stream, one un-graphed network call"})
    return {"roofline": roof, "kernels": rows, "kernel_families": families, "profiled_call_ms": round(total_ms, 3),
            "_flops": flops3}


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs: the oracle port (the reference is pure Python/PyTorch; its own classes cannot travel to the GPU box, so the
# state-dict driven restatement pinned against them in tests/test_oracle_golden.py is what is timed here)
# ---------------------------------------------------------------------------------------------------------------------
def _oracle_setup(kind):
    from oracle import sgm_oracle as so
    from oracle.weights import load_manifest
    man = load_manifest(kind)
    g = torch.Generator().manual_seed(0)
    sd = {k: (torch.rand(shp, generator=g) - 0.5) * 0.04 for k, (shp, _) in man.items()}
    ucfg, icfg = so.TV2V_UNET_CFG, None
    if kind == "tvi2v":
        ucfg = dict(so.TV2V_UNET_CFG, enable_attention3d_crossframe=True, ST3DCA_ca_type="center_self")
        icfg = dict(so.TV2V_CONTROLNET_CFG, no_add_x=True, set_input_hint_block_as_identity=True, disable_text_ca=True)

    def call(x, t, c):
        with torch.no_grad():
            return so.wrapper_forward(sd, ucfg, so.TV2V_CONTROLNET_CFG, x, t, c, icfg)
    return call


def _oracle_inputs(kind, B, T, h, w):
    g = torch.Generator().manual_seed(1)
    c = {"crossattn": torch.randn(B, 77, 768, generator=g), "control_hint": torch.rand(B, 3, T, 8 * h, 8 * w, generator=g) * 2 - 1}
    if kind == "tvi2v":
        c["cond_feat"] = torch.randn(B, 4, h, w, generator=g)
    return torch.randn(B, 4, T, h, w, generator=g), torch.full((B,), 500, dtype=torch.long), c


def cpu_baseline(kind, T, h, w, budget_s, calls, warm):
    """Time the oracle port on a bounded sample of the workload and scale by the exact FLOP ratio.
    The sample keeps the CFG batch of 2 and the full per-frame resolution where the budget allows and cuts the number
    of keyframes (every op of the network is linear in T except temporal attention, 0.05 % of the FLOPs)."""
    from ccedit_b200.census import network_flops
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    call = _oracle_setup(kind)
    full = network_flops(kind, 2, T, h, w)["total"]
    probe = (2, 1, max(8, h // 4), max(8, w // 4))
    xp = _oracle_inputs(kind, *probe)
    call(*xp)                                            # warm (thread pool, allocator)
    t0 = time.perf_counter()
    call(*xp)
    rate = network_flops(kind, *probe)["total"] / (time.perf_counter() - t0)        # FLOP/s at the probe shape
    ladder = [(2, 2, h, w), (2, 1, h, w), (2, 2, h // 2, w // 2), (2, 1, h // 2, w // 2), probe]
    n_calls = calls + warm
    shape = next((s for s in ladder if network_flops(kind, *s)["total"] / rate * n_calls <= budget_s), probe)
    xs = _oracle_inputs(kind, *shape)
    for _ in range(warm):
        call(*xs)
    times = []
    for _ in range(calls):
        t0 = time.perf_counter()
        call(*xs)
        times.append(time.perf_counter() - t0)
    t_call = sum(times) / len(times)
    fl = network_flops(kind, *shape)["total"]
    t_full = t_call * full / fl                          # seconds per full-size network call (estimated)
    return {"value": round(1.0 / (2 * t_full), 6), "unit": UNIT, "cores": cores, "kind": "port", "estimated": True,
            "sample": f"ESTIMATE from a bounded sample - oracle port (fp32, torch {torch.__version__}, {cores} threads): {calls} network call(s) at CFG batch "
                      f"{shape[0]} x {shape[1]} keyframe(s) x latent {shape[2]}x{shape[3]} = {fl / 1e12:.2f} TFLOP in "
                      f"{t_call:.2f} s ({fl / t_call / 1e12:.3f} TFLOP/s); scaled by the FLOP ratio {full / fl:.1f} to the "
                      f"full {full / 1e12:.2f} TFLOP call, 2 calls per step",
            "seconds_per_sample_call": round(t_call, 3)}


def library_gpu_baseline(kind, T, h, w, dev):
    """The stronger baseline BASELINE.md section 3 names: the reference's op sequence run EAGERLY ON THE SAME GPU with
    PyTorch's library kernels (cuDNN convolutions, cuBLAS linears, the SDPA flash kernel) in fp16 - how the reference itself
    runs (`torch.cuda.amp.autocast`, sampling_tv2v.py:361-362).  The reference's Python sources cannot travel to the GPU
    box, so the oracle's restatement of them (same ATen calls, same rearranges; pinned to the reference in
    tests/test_oracle_golden.py) is what runs, with weights pre-cast to fp16 (no per-call autocast casts: a best case for
    the baseline).  None of this repo's kernels are involved.  One warm-up call, then 2 timed network calls."""
    from oracle import sgm_oracle as so
    from oracle.weights import load_manifest
    man = load_manifest(kind)
    g = torch.Generator(device=dev).manual_seed(0)
    sd = {k: ((torch.rand(shp, generator=g, device=dev) - 0.5) * 0.04).half() for k, (shp, _) in man.items()}
    ucfg, icfg = so.TV2V_UNET_CFG, None
    if kind == "tvi2v":
        ucfg = dict(so.TV2V_UNET_CFG, enable_attention3d_crossframe=True, ST3DCA_ca_type="center_self")
        icfg = dict(so.TV2V_CONTROLNET_CFG, no_add_x=True, set_input_hint_block_as_identity=True, disable_text_ca=True)
    x, t, c = _oracle_inputs(kind, 2, T, h, w)
    x, t = x.to(dev).half(), t.to(dev)
    c = {k: v.to(dev).half() for k, v in c.items()}
    old = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    try:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            so.wrapper_forward(sd, ucfg, so.TV2V_CONTROLNET_CFG, x, t, c, icfg)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                so.wrapper_forward(sd, ucfg, so.TV2V_CONTROLNET_CFG, x, t, c, icfg)
            e1.record()
            torch.cuda.synchronize()
    finally:
        torch.backends.cudnn.benchmark = old
    ms = e0.elapsed_time(e1) / 2
    return {"what": "oracle restatement of the reference path, eager PyTorch on this GPU under torch.autocast(fp16) (cuDNN / "
                    "cuBLAS / SDPA kernels), weights pre-cast to fp16, cudnn.benchmark on; 1 warm-up + 2 timed network calls",
            "ms_per_network_call": round(ms, 2), "value": round(1e3 / (2 * ms), 4), "unit": UNIT,
            "torch": torch.__version__}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port: state-dict driven restatement pinned to the unmodified
    reference, which is Python and cannot travel to the GPU box) on all host cores of this box, at the SAME workload
    as `--impl ours`: real full-size network calls (CFG batch 2 x T keyframes x latent h x w), no extrapolation.
    One sampler step = 2 network calls; a full-size call takes about a minute on these hosts, so the run times as many
    calls as fit `--ref-budget` seconds (at least one) after one small warm-up call, instead of K steps + W warm-up
    steps (K = 20 steps would be ~40 minutes).  `steps` / `warmup` in the line are what was actually run;
    `steps_requested` / `warmup_requested` echo the command line."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ccedit_b200.census import network_flops
    kind, T, h, w = args.kind, args.frames, args.height // 8, args.width // 8
    B = 2 * args.clips_per_gpu
    w0 = time.perf_counter()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    call = _oracle_setup(kind)
    xp = _oracle_inputs(kind, 2, 1, max(8, h // 4), max(8, w // 4))
    call(*xp)                                            # warm-up (thread pool, allocator) on a small shape
    xs = _oracle_inputs(kind, B, T, h, w)
    setup_s = time.perf_counter() - w0
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        call(*xs)
        times.append(time.perf_counter() - t0)
        spent = time.perf_counter() - t_start
        # stop when another call would not fit the budget; prefer whole steps (pairs of calls)
        if spent + times[-1] > args.ref_budget or len(times) >= 2 * max(1, args.steps):
            break
    t_call = sum(times) / len(times)
    fl = network_flops(kind, B, T, h, w)["total"]
    value = args.clips_per_gpu / (2 * t_call)
    steps_run = max(1, len(times) // 2)                  # whole sampler steps covered by the timed calls (2 calls per step)
    base = {"value": round(value, 6), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port (fp32, torch {torch.__version__}, {cores} threads): {len(times)} full-size network "
                      f"call(s) at CFG batch {B} x {T} keyframes x latent {h}x{w} = {fl / 1e12:.2f} TFLOP each, "
                      f"{t_call:.1f} s per call ({fl / t_call / 1e12:.3f} TFLOP/s); 2 calls per step; nothing extrapolated",
            "seconds_per_call": [round(t, 2) for t in times]}
    res = {
        "impl": "reference", "metric": METRIC, "value": round(value, 6), "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps_run, "warmup": 0,
        "steps_requested": args.steps, "warmup_requested": args.warmup, "network_calls_timed": len(times),
        "ms_per_step": round(2e3 * t_call, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "impl_notes": {"what": "CPU fp32 oracle port of the reference path, all host threads, full-size calls",
                       "warmup": "one small-shape call (thread pool / allocator), untimed",
                       "setup_s": round(setup_s, 1)},
        "cpu_baseline": base,
        "e2e": {"value": round(value, 6), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.perf_counter() - w0, 1),
    }
    print(json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kind", default="tv2v", choices=["tv2v", "tvi2v"])
    ap.add_argument("--frames", type=int, default=17)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=768)
    ap.add_argument("--cfg-scale", type=float, default=7.5)
    ap.add_argument("--sampler-steps", type=int, default=30)
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of CUDA-graph replay")
    ap.add_argument("--sampler", default="fused", choices=["fused", "plain"],
                    help="fused: FusedDPMPP2SAncestralSampler (step graph); plain: PyTorch-elementwise sampler_step")
    ap.add_argument("--no-dedup", action="store_true", help="fused sampler without CFG de-duplication")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs[2..4] section (tvi2v, 2 clips per GPU, sweep)")
    ap.add_argument("--clips-per-gpu", type=int, default=1, help="clips run as one batch on every GPU (BASELINE configs[3]: 2)")
    ap.add_argument("--cpu-budget", type=float, default=25.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--ref-budget", type=float, default=170.0,
                    help="seconds of timed CPU work for --impl reference (at least one full-size network call is always timed)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
