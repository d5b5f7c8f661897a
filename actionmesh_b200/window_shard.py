"""Temporal (frame) sharding of ONE 16-frame denoising window across the GPUs of a box.

Everything in a DiT block is token-local — LayerNorm, QKV / out / MLP GEMMs, the cross-attention to the frame's own 257
context tokens (temporal_denoiser.py:221-226), the skip linears, the time token — except the inflated self-attention,
where every query attends to the keys of all T·L tokens (attention_processor.py:49-65).  So each rank owns T/world
consecutive frames of every CFG branch, keeps its queries local, and per layer all-gathers K and V (post RMSNorm/RoPE,
bf16) over NCCL.  The attention kernel consumes the gathered buffer in place as `kv_chunks = world` chunks (5-D TMA map
with a free chunk stride; no re-layout).  The fp32 latents are sharded the same way and all-gathered once per window.

One process per GPU (`torch.distributed`, backend "nccl" on the box, "gloo" in the CPU tests of the index logic).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


def frame_partition(n_frames: int, world: int, rank: int) -> slice:
    """Frames [rank*T/world, (rank+1)*T/world).  T must divide evenly (16 frames over 1/2/4/8 ranks)."""
    if n_frames % world:
        raise ValueError(f"{n_frames} frames do not shard evenly over {world} ranks")
    per = n_frames // world
    return slice(rank * per, (rank + 1) * per)


def gather_kv(kv_local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather the local (B*T_local*L, 2D) [K|V] rows of one layer -> (world, B*T_local*L, 2D), rank-major."""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(kv_local.shape), dtype=kv_local.dtype, device=kv_local.device)
    dist.all_gather_into_tensor(out.view(-1, kv_local.shape[-1]), kv_local.contiguous(), group=group)  # concat along dim 0
    return out


def chunked_kv_views(kv_all: torch.Tensor, B: int, s_local: int, H: int, dh: int):
    """(world, B*s_local, 2*H*dh) gathered buffer -> K, V views of logical shape (B, world, s_local, H, dh): chunk c of
    batch b is rank c's keys; concatenating the chunks in rank order restores the window's frame order."""
    world = kv_all.shape[0]
    D = H * dh
    kv5 = kv_all.view(world, B, s_local, 2 * D).permute(1, 0, 2, 3)  # (B, world, s_local, 2D), no copy
    k = kv5[..., 0:D].unflatten(-1, (H, dh))
    v = kv5[..., D:2 * D].unflatten(-1, (H, dh))
    return k, v


def configure_nccl_env() -> None:
    """Defaults for the per-layer K/V all-gather; call BEFORE `init_process_group` (NCCL caches its parameters at first use).
    Measured with tools/shard_profile.py on 8x B200 (profiles/r01_shard_profile_8gpu.log): with NCCL's default choice the
    235 MB gathers take 0.93 ms each next to the compute kernels and ~10 ms per step stay exposed; the Simple protocol on
    32 channels brings them to 0.55 ms and the exposure to ~1 ms per step (step 91 -> 81 ms under the profiler).
    Explicit user settings win (setdefault)."""
    os.environ.setdefault("NCCL_PROTO", "Simple")
    os.environ.setdefault("NCCL_MIN_NCHANNELS", "32")


class _EventWork:
    """`.wait()` with the semantics of an async NCCL work handle: the CURRENT stream waits for the recorded event."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


class PeerGather:
    """EXPERIMENTAL (off by default, `AMB_SHARD_P2P=1`; written at the end of round 1 and NOT yet run on hardware):
    the per-layer K/V all-gather over NVLink peer memory with the COPY ENGINES instead of an NCCL kernel.

    Why: the timeline of the 8-GPU sharded window (profiles/r01_shard_profile_8gpu.log) shows that what the all-gather
    costs is not its latency but its SMs — the NCCL kernel (32 channels) runs next to persistent GEMMs / multi-wave
    attention and inflates their time by ~12 ms per step.  DMA copies take no SM.

    How: every rank's K/V projection writes into a symmetric-memory buffer (torch.distributed._symmetric_memory, one
    allocation per rank, mapped into every peer).  `gather()` then, on a side stream: device-side barrier across ranks
    (all projections of this layer/branch are complete) -> `world` contiguous peer->local copies (cudaMemcpyAsync D2D,
    starting with the local chunk, peers visited in a rank-rotated order so each NVLink port sees one reader at a time)
    -> event.  The buffer is double-buffered by layer parity: a rank overwrites parity p again two layers later, after it
    passed the barrier of the layer in between, which every peer enters only after its own copies of this layer."""

    def __init__(self, shard: "FrameShard", branches: int, rows: int, cols: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm_mem

        self.world, self.rank = shard.world, shard.rank
        group = shard.group if shard.group is not None else dist.group.WORLD
        self.buf = symm_mem.empty((2, branches, rows, cols), dtype=torch.bfloat16, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        shape = (2, branches, rows, cols)
        self.peers = [self.buf if r == self.rank else self.hdl.get_buffer(r, shape, torch.bfloat16) for r in range(self.world)]
        self.stream = torch.cuda.Stream(device=device, priority=-1)

    def local(self, branch: int, parity: int) -> torch.Tensor:
        return self.buf[parity, branch]

    def gather(self, branch: int, parity: int, kv_all_b: torch.Tensor) -> _EventWork:
        """kv_all_b (world, rows, cols) <- every rank's local(branch, parity); asynchronous w.r.t. the current stream."""
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self.hdl.barrier(channel=branch)
            for k in range(self.world):
                r = (self.rank + k) % self.world
                kv_all_b[r].copy_(self.peers[r][parity, branch], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        return _EventWork(done)


class FrameShard:
    """Rank-local view of a frame-sharded window, handed to B200Denoiser._forward_packed."""

    def __init__(self, group=None):
        if group is None and dist.get_backend() == "nccl" and os.environ.get("AMB_SHARD_HIPRI", "1") != "0":
            # dedicated communicator on a HIGH-PRIORITY stream: the per-layer K/V all-gather has to run concurrently with
            # compute kernels that fill every SM (persistent GEMMs, multi-wave attention); at normal priority its CTAs
            # queue behind the pending compute CTAs and the gather is effectively serialised.
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            group = dist.new_group(ranks=list(range(dist.get_world_size())), pg_options=opts)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def frames(self, n_frames: int) -> slice:
        return frame_partition(n_frames, self.world, self.rank)

    def gather_latents(self, local: torch.Tensor) -> torch.Tensor:
        """(1, T_local, N, C) fp32 per rank -> (1, T, N, C) on every rank (once per window)."""
        out = torch.empty((self.world,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out.view(-1, *local.shape[2:]), local[0].contiguous(), group=self.group)
        return out.reshape(1, -1, *local.shape[2:])


def run_temporal_bench(args, rank: int, local: int, world: int):
    """bench.py --mode temporal: ONE default window, frames sharded over `world` ranks (strong scaling)."""
    import json

    from . import ops
    from .denoiser import B200Denoiser, DenoiserConfig
    from .guidance import ClassifierFreeGuidance
    from .scheduler import B200SchedulerFlow

    dev = torch.device("cuda", local)
    K, W = args.steps, max(args.warmup, 0)
    T, N, C, S, Dc = 16, 2048, 64, 257, 1024
    shard = FrameShard()
    model = B200Denoiser(DenoiserConfig()).to(dev)
    model.init_random_(seed=1234)  # same seed on every rank => replicated weights
    cf = ClassifierFreeGuidance(guidance_at_inference=[[0, 1], [1, 1]], guidance_scales=[7.5])
    g = torch.Generator(device="cpu").manual_seed(44)
    lat = torch.randn(1, T, N, C, generator=g).to(dev)
    ctx = torch.randn(1, T, S, Dc, generator=torch.Generator().manual_seed(5)).to(dev)
    mask = torch.zeros(1, T, device=dev)
    mask[0, 0] = 1.0
    framestep = torch.arange(T, dtype=torch.float32)[None]
    sch = B200SchedulerFlow(num_inference_steps=W + K, shift=3.0, is_additive=True)
    ev = {}
    marks = {"l0": 0}

    def cb(step, total):
        if step == W:
            ev["t0"] = torch.cuda.Event(enable_timing=True)
            ev["t0"].record()
            marks["l0"] = ops.launch_count
        if step == total:
            ev["t1"] = torch.cuda.Event(enable_timing=True)
            ev["t1"].record()

    dist.barrier()
    torch.cuda.synchronize()
    if W == 0:
        cb(0, W + K)
    sch.denoise(model, cf, lat, ctx, device=dev, mask=mask, framestep=framestep, step_callback=cb, shard=shard)
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([ev["t0"].elapsed_time(ev["t1"])], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        from bench import F_STEP, METRIC, UNIT, WORKLOAD  # type: ignore

        line = {
            "metric": METRIC, "value": K / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": f"temporal-shard x{world} (frames/rank {T // world}), "
                       "K/V all-gather per layer over NCCL", "l2": "inputs larger than L2"},
            "step_flops": F_STEP, "gpu_launches": ops.launch_count - marks["l0"],
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
