"""bench.py — denoiser steps/sec of the Stage-I temporal-3D-diffusion hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode dp|temporal]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one denoiser step of the default window: CFG batch of 2 branches x T=16 frames x
N=2048 latent tokens (+1 time token) through the 21-block DiT (width 2048, 16 heads), CFG combine (7.5) + Euler update.
5.469e14 algorithmic FLOP per step, of which the inflated self-attention QK^T+PV is 3.698e14.
Weights are seeded-random (no checkpoints offline), inputs synthetic; the per-step working set (2.9 GB weights +
GBs of activations) is far larger than the 126 MB L2, so no explicit L2 flush is needed between iterations.

Prints ONE JSON line (rank 0).  N > 1: `--mode dp` (default) runs one independent window per GPU (whole-clip data
parallel, no collective; weak scaling); `--mode temporal` shards the 16 frames of ONE window across ranks with the
temporal-attention K/V all-gathered over NCCL (strong scaling).
`--impl reference` times the reference's own CPU path (the fp32 oracle port, all host threads) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_STEP = 5.469e14          # algorithmic FLOP per denoiser step, T=16
F_ATTN_LAUNCH = 4.0 * 2 * (16 * 2049) ** 2 * 2048   # one inflated self-attention launch (QK^T + PV), B=2
METRIC = "denoiser_steps_per_sec"
UNIT = "steps/s"
WORKLOAD = "davis_camel-shaped default window: CFG x2, T=16 frames, N=2048 tokens, 21-block DiT width 2048, guidance 7.5"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1380.2), d.get("hbm_gbs", 6570.3), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm = sorted(int(float(r[0])) for r in rows if len(r) >= 7)
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) >= 7:
                for nme, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(float(rows[0][1])) if rows and len(rows[0]) >= 2 else None,
                "power_w_max": max((float(r[2]) for r in rows if len(r) >= 7), default=None),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ reference (CPU) arm
def cpu_reference_sample(threads: int, T: int = 16, N: int = 2048):
    """One DiT block (block.py:110-154) of the cond branch at the full window shape through the fp32 oracle port.
    Returns (seconds, fraction_of_step_flops)."""
    import torch

    from oracle import denoiser_oracle as do
    from oracle import synth

    torch.set_num_threads(threads)
    cfg = do.DenoiserConfig(num_layers=1, inflated_layers=(0,))
    if not hasattr(cpu_reference_sample, "_sd"):
        cpu_reference_sample._sd = {k: v for k, v in synth.make_state_dict(cfg, 1234).items()}
    sd = cpu_reference_sample._sd
    g = torch.Generator().manual_seed(0)
    L = N + 1
    h = torch.randn(T, L, cfg.width, generator=g)
    ctx = torch.randn(T, 257, cfg.cross_attention_dim, generator=g)
    pos = torch.arange(T, dtype=torch.float32)
    cos, sin = do.rotary_tables(cfg.head_dim, pos)
    rope = (cos[:, None].repeat(1, L, 1), sin[:, None].repeat(1, L, 1))
    t0 = time.perf_counter()
    with torch.no_grad():
        do.block_forward(sd, "blocks.0.", h, ctx, cfg.num_attention_heads, T, True, rope, None)
    dt = time.perf_counter() - t0
    S = T * L
    D, F_, Dc = cfg.width, int(cfg.width * cfg.mlp_ratio), cfg.cross_attention_dim
    flops = 4.0 * S * S * D + 4.0 * S * 257 * D + 2.0 * S * D * (6 * D + 2 * F_) + 2.0 * (T * 257) * Dc * 2 * D
    full = F_STEP if (T == 16 and N == 2048) else None
    return dt, flops, full


def run_reference(args, rank: int):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    T = 16
    # keep the whole run within a few minutes: probe once, shrink the frame count of the sample if needed
    t_probe, flops, _ = cpu_reference_sample(threads, T=4)
    est16 = t_probe * (4.0 * (16 * 2049) ** 2 * 2048) / (4.0 * (4 * 2049) ** 2 * 2048)
    budget = 240.0
    while T > 2 and (args.steps + args.warmup) * est16 * ((T / 16.0) ** 2) > budget:
        T //= 2
    for _ in range(args.warmup):
        cpu_reference_sample(threads, T=T)
    times, fl = [], 0.0
    for _ in range(args.steps):
        dt, fl, _ = cpu_reference_sample(threads, T=T)
        times.append(dt)
    sec = sum(times) / len(times)
    flops_per_s = fl / sec
    steps_per_s = flops_per_s / F_STEP  # sample FLOP rate extrapolated to the 5.469e14-FLOP step
    sample = (f"1 of 21 DiT blocks, cond branch, T={T} frames x 2049 tokens, fp32 oracle port of the reference modules "
              f"(oracle/denoiser_oracle.py), {sec:.2f} s/sample at {flops_per_s / 1e12:.3f} TFLOP/s, extrapolated by FLOPs "
              f"to the {F_STEP:.3e}-FLOP step")
    line = {
        "impl": "reference", "metric": METRIC, "value": steps_per_s, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / steps_per_s, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": steps_per_s, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": steps_per_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from actionmesh_b200 import ops
    from actionmesh_b200.denoiser import B200Denoiser, DenoiserConfig
    from actionmesh_b200.guidance import ClassifierFreeGuidance
    from actionmesh_b200.scheduler import B200SchedulerFlow

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        from actionmesh_b200.window_shard import configure_nccl_env

        configure_nccl_env()  # NCCL protocol / channel defaults for the sharded window's K/V all-gather (before init)
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 0)
    T, N, C, S, Dc = 16, 2048, 64, 257, 1024

    if args.mode == "temporal" and world > 1:
        from actionmesh_b200.window_shard import run_temporal_bench
        return run_temporal_bench(args, rank, local, world)

    model = B200Denoiser(DenoiserConfig()).to(dev)
    model.init_random_(seed=1234)
    cf = ClassifierFreeGuidance(guidance_at_inference=[[0, 1], [1, 1]], guidance_scales=[7.5])
    g = torch.Generator(device="cpu").manual_seed(44 + rank)
    host_lat = torch.randn(1, T, N, C, generator=g).pin_memory()
    host_ctx = torch.randn(1, T, S, Dc, generator=torch.Generator().manual_seed(5 + rank)).pin_memory()
    host_mask = torch.zeros(1, T)
    host_mask[0, 0] = 1.0  # anchor frame observed, like the first AR window
    framestep = torch.arange(T, dtype=torch.float32)[None]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`): W + K steps of one denoise() call, timed with CUDA events
    sch = B200SchedulerFlow(num_inference_steps=W + K, shift=3.0, is_additive=True)
    lat = host_lat.to(dev)
    ctx = host_ctx.to(dev)
    mask = host_mask.to(dev)
    ev = {}
    marks = {"launch0": 0}

    def cb(step, total):
        if step == W:
            ev["t0"] = torch.cuda.Event(enable_timing=True)
            ev["t0"].record()
            marks["launch0"] = ops.launch_count
            ops.event_log = []
            ops.event_tags = {"attn_self"}
        if step == total:
            ev["t1"] = torch.cuda.Event(enable_timing=True)
            ev["t1"].record()

    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if W == 0:
        cb(0, W + K)
    sch.denoise(model, cf, lat, ctx, device=dev, mask=mask, framestep=framestep, step_callback=cb)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = ev["t0"].elapsed_time(ev["t1"])
    launches = ops.launch_count - marks["launch0"]
    attn_events = ops.event_log
    ops.event_log = None
    attn_ms = [a.elapsed_time(b) for _, a, b in attn_events]
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    steps_per_s = world * K / (ms_total / 1e3)

    # ---------------- end-to-end through the public API with HOST buffers
    e2e_steps = K
    sch2 = B200SchedulerFlow(num_inference_steps=e2e_steps, shift=3.0, is_additive=True)
    host_out = torch.empty(1, T, N, C).pin_memory()
    d2h = {"bytes": 0}
    dev_lat_holder = {}

    def cb2(step, total):  # per-step device->host read of the step's result (the current latents)
        host_out.copy_(dev_lat_holder["x"], non_blocking=True)
        d2h["bytes"] += host_out.numel() * 4

    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    lat2 = host_lat.to(dev, non_blocking=True)
    ctx2 = host_ctx.to(dev, non_blocking=True)
    mask2 = host_mask.to(dev, non_blocking=True)
    dev_lat_holder["x"] = lat2
    out = sch2.denoise(model, cf, lat2, ctx2, device=dev, mask=mask2, framestep=framestep, step_callback=cb2)
    host_out.copy_(out, non_blocking=True)
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = world * e2e_steps / (float(e2e_ms.item()) / 1e3)
    h2d_bytes = (host_lat.numel() + host_ctx.numel() + host_mask.numel()) * 4

    # ---------------- N > 1: also time ONE window frame-sharded over all ranks (strong scaling, K/V all-gather)
    temporal = None
    if world > 1 and 16 % world == 0:
        try:
            from actionmesh_b200.window_shard import FrameShard

            shard = FrameShard()
            g0 = torch.Generator(device="cpu").manual_seed(44)
            lat_t = torch.randn(1, T, N, C, generator=g0).to(dev)      # identical window on every rank
            ctx_t = torch.randn(1, T, S, Dc, generator=torch.Generator().manual_seed(5)).to(dev)
            sch3 = B200SchedulerFlow(num_inference_steps=W + K, shift=3.0, is_additive=True)
            ev3 = {}

            def cb3(step, total):
                if step == W:
                    ev3["t0"] = torch.cuda.Event(enable_timing=True)
                    ev3["t0"].record()
                if step == total:
                    ev3["t1"] = torch.cuda.Event(enable_timing=True)
                    ev3["t1"].record()

            barrier()
            if W == 0:
                cb3(0, W + K)
            sch3.denoise(model, cf, lat_t, ctx_t, device=dev, mask=mask, framestep=framestep, step_callback=cb3, shard=shard)
            barrier()
            t3 = torch.tensor([ev3["t0"].elapsed_time(ev3["t1"])], device=dev, dtype=torch.float64)
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            temporal = {"value": K / (float(t3.item()) / 1e3), "unit": UNIT, "ms_per_step": float(t3.item()) / K,
                        "scaling": "strong", "frames_per_rank": T // world,
                        "note": "ONE default window, frames sharded over the ranks, temporal-attention K/V all-gathered per "
                                "layer over NCCL (window_shard.py); steps/s of that single window"}
        except Exception as exc:  # noqa: BLE001 - an optional leg must never cost the main JSON line
            temporal = {"error": f"{type(exc).__name__}: {exc}"[:400]}

    # ---------------- sec/video of the Stage-I path through the public pipeline API (N = 1 only): 16 synthetic RGB frames
    # -> CUDA preprocessing (PIL-exact bicubic resize/crop/normalise) -> DinoV2-L -> one 16-frame window, default 30 steps, CFG 7.5; then Stage II (Stage 0 out of scope)
    video = None
    if world == 1 and not args.no_video:
        try:
            import numpy as np
            from PIL import Image

            from actionmesh_b200.image_encoder import B200ImageEncoder
            from actionmesh_b200.pipeline import Stage1Pipeline, VideoInput

            enc = B200ImageEncoder().to(dev)
            enc.init_random_(seed=1235)  # DinoV2-L/14 shape, seeded random weights (no checkpoints offline)
            rng = np.random.default_rng(7)
            frames = [Image.fromarray(rng.integers(0, 255, (512, 512, 3), dtype=np.uint8), "RGB") for _ in range(T)]
            pipe = Stage1Pipeline(model, B200SchedulerFlow(num_inference_steps=30, shift=3.0, is_additive=True), cf, enc)
            anchor = torch.randn(1, N, C, generator=torch.Generator().manual_seed(99))
            vin = VideoInput(frames, torch.arange(T, dtype=torch.float32))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx_v = pipe.encode_all_frames(vin)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            bank = pipe(vin, anchor, seed=44, stage_1_steps=30, context=ctx_v)
            lat_out, _ = bank.get_ordered()
            lat_host = lat_out.cpu()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            video = {"sec_per_video_stage1": t2 - t0, "dinov2_encode_s": t1 - t0, "denoise_30_steps_s": t2 - t1,
                     "frames": T, "steps": 30, "finite": bool(torch.isfinite(lat_host).all()),
                     "note": "Stage-I path only (uint8 frames -> CUDA BitImageProcessor-equivalent preprocessing -> DinoV2 + 1 window x 30 steps, CFG 7.5) through "
                             "Stage1Pipeline; Stage 0 (TripoSG) is out of scope and not included; Stage II is timed separately below"}
            del enc, pipe
            # Stage II on the same window: 16-block trunk re-run for each of the 15 target times + the
            # fp32-grade vertex-query block for V = 20 000 anchor vertices (+ normals), B200Autoencoder.forward, host in/out.
            from actionmesh_b200.autoencoder import B200Autoencoder

            ae = B200Autoencoder().to(dev)
            ae.init_random_(seed=1236)
            gq = torch.Generator().manual_seed(13)
            pts = torch.randn(1, 20000, 3, generator=gq)
            pts = pts / pts.norm(dim=-1, keepdim=True) * 0.6
            query = torch.cat([pts, pts / 0.6], dim=-1)
            tgt = torch.linspace(0, 1, T)[None, 1:]
            ae.forward(lat_host[None, :3], torch.arange(3.0)[None], torch.zeros(1), tgt[:, :1], query[:, :512])  # warm-up
            ops.event_log, ops.event_tags = [], {"s2_attn", "s2_gemm", "s2_q"}
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            disp = ae.forward(lat_host[None], torch.arange(T, dtype=torch.float32)[None], torch.zeros(1), tgt, query)
            verts = ae.apply_displacement(query[..., :3].to(dev), disp).cpu()
            torch.cuda.synchronize()
            t4 = time.perf_counter()
            s2 = {}
            for tag, e0, e1 in ops.event_log:
                s2[tag] = s2.get(tag, 0.0) + e0.elapsed_time(e1)
            ops.event_log = None
            video.update({"stage2_decode_s": t4 - t3, "stage2_targets": int(tgt.shape[1]), "stage2_vertices": 20000,
                          "stage2_kernel_ms": {"trunk_attention": s2.get("s2_attn"), "trunk_gemm": s2.get("s2_gemm"),
                                               "query_path_gemm": s2.get("s2_q")},
                          "stage2_finite": bool(torch.isfinite(verts).all()),
                          "sec_per_video_stage1_plus_stage2": (t2 - t0) + (t4 - t3)})
            del ae
        except Exception as exc:  # noqa: BLE001 - an optional leg must never cost the main JSON line
            video = dict(video or {}, error=f"{type(exc).__name__}: {exc}"[:400])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, _, peak_src = _peaks()
    attn_avg_ms = sum(attn_ms) / max(1, len(attn_ms))
    achieved_tf = F_ATTN_LAUNCH / (attn_avg_ms * 1e-3) / 1e12 if attn_ms else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "attn_self_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        Tc = 4
        t_s, fl, _ = cpu_reference_sample(threads, T=Tc)
        if t_s < 9.0:  # a fast host: take the 3.3x larger sample so the baseline rests on ~10-30 s of CPU work
            Tc = 8
            t_s, fl, _ = cpu_reference_sample(threads, T=Tc)
        rate = fl / t_s
        cpu = {"value": rate / F_STEP, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"1 of 21 DiT blocks, cond branch, T={Tc} frames x 2049 tokens (fp32 oracle port), {t_s:.2f} s at "
                         f"{rate / 1e12:.3f} TFLOP/s, extrapolated by FLOPs to the {F_STEP:.3e}-FLOP step; "
                         f"`--impl reference` runs the longer sample"}
    line = {
        "metric": METRIC, "value": steps_per_s, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "parallelism": f"dp{world}" if world > 1 else "single", "steps_schedule": "shift 3.0",
                   "l2": "inputs larger than L2 (2.9 GB weights + >3 GB activations per step; no flush needed)",
                   "weights": "seeded random (no checkpoints offline)"},
        "step_flops": F_STEP, "model_tflops": F_STEP * steps_per_s / world / 1e12,
        "roofline": {"bound": "tensor", "kernel": "flash_attn_fwd_v4_kernel (inflated self-attention, d_h 128)",
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (achieved_tf / peak_tf) if achieved_tf else None, "traffic": traffic,
                     "peak_source": peak_src, "launches_timed": len(attn_ms), "avg_launch_ms": attn_avg_ms,
                     "flops_per_launch": F_ATTN_LAUNCH,
                     "share_of_step": (sum(attn_ms) / ms_total) if attn_ms else None},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes // e2e_steps,
                "d2h_bytes_per_step": (d2h["bytes"] + host_out.numel() * 4) // e2e_steps,
                "note": "one SchedulerFlow.denoise() call from pinned host buffers incl. per-window context K/V precompute; "
                        "window inputs are copied once (amortised per step), the latents are read back every step"},
        "gpu_launches": launches, "clocks": clocks,
    }
    if temporal is not None:
        line["temporal_shard"] = temporal
    if video is not None:
        line["video"] = video
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="dp", choices=["dp", "temporal"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-video", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return
    run_b200(args)


if __name__ == "__main__":
    main()
