"""Kineto timeline of ONE step of the frame-sharded window on rank 0 (development tool): how long the NCCL all-gather
kernels run, how much of that is overlapped by compute kernels, and what the compute kernels cost while a gather is in
flight.  Launch with torchrun (2+ ranks).  Prints a small JSON summary on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from actionmesh_b200.denoiser import B200Denoiser  # noqa: E402
from actionmesh_b200.guidance import ClassifierFreeGuidance  # noqa: E402
from actionmesh_b200.scheduler import B200SchedulerFlow  # noqa: E402
from actionmesh_b200.window_shard import FrameShard  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    if os.environ.get("SHARD_PROFILE_NCCL_DEFAULTS", "0") == "1":
        from actionmesh_b200.window_shard import configure_nccl_env

        configure_nccl_env()
    dist.init_process_group("nccl", device_id=dev)
    model = B200Denoiser().to(dev)
    model.init_random_(1234)
    T, N, C, S, Dc = 16, 2048, 64, 257, 1024
    shard = FrameShard()
    lat = torch.randn(1, T, N, C, generator=torch.Generator().manual_seed(44)).to(dev)
    ctx = torch.randn(1, T, S, Dc, generator=torch.Generator().manual_seed(5)).to(dev)
    mask = torch.zeros(1, T, device=dev)
    mask[0, 0] = 1
    fs = torch.arange(T, dtype=torch.float32)[None]
    cf = ClassifierFreeGuidance(guidance_at_inference=[[0, 1], [1, 1]], guidance_scales=[7.5])
    sch = B200SchedulerFlow(num_inference_steps=3, shift=3.0, is_additive=True)
    sch.denoise(model, cf, lat.clone(), ctx, device=dev, mask=mask, framestep=fs, shard=shard)  # warm-up
    torch.cuda.synchronize()
    dist.barrier()
    sch1 = B200SchedulerFlow(num_inference_steps=2, shift=3.0, is_additive=True)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        sch1.denoise(model, cf, lat.clone(), ctx, device=dev, mask=mask, framestep=fs, shard=shard)
        torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
        ker = [(e.name, e.time_range.start, e.time_range.end) for e in ev if "Memcpy" not in e.name and "Memset" not in e.name]
        nccl = [(s, t) for n, s, t in ker if "nccl" in n.lower()]
        comp = sorted((s, t) for n, s, t in ker if "nccl" not in n.lower())
        # union of compute intervals
        merged = []
        for s, t in comp:
            if merged and s <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], t)
            else:
                merged.append([s, t])

        def overlap(a, b):
            tot = 0.0
            for s, t in merged:
                lo, hi = max(a, s), min(b, t)
                if hi > lo:
                    tot += hi - lo
            return tot

        nccl_total = sum(t - s for s, t in nccl)
        nccl_overlapped = sum(overlap(s, t) for s, t in nccl)
        span = max(t for _, _, t in ker) - min(s for _, s, _ in ker)
        by_name = {}
        for n, s, t in ker:
            k = n.split("<")[0][:60]
            by_name.setdefault(k, [0, 0.0])
            by_name[k][0] += 1
            by_name[k][1] += t - s
        top = sorted(by_name.items(), key=lambda kv: -kv[1][1])[:8]
        # attention kernels running while a gather is in flight vs alone
        att = [(s, t) for n, s, t in ker if "flash_attn" in n and (t - s) > 300]
        def busy(s, t):
            return any(min(t, b) - max(s, a) > 0.5 * (t - s) for a, b in nccl)
        att_busy = [t - s for s, t in att if busy(s, t)]
        att_free = [t - s for s, t in att if not busy(s, t)]
        print(json.dumps({
            "world": world, "steps_profiled": 2, "span_ms": span / 1e3, "compute_union_ms": sum(t - s for s, t in merged) / 1e3,
            "nccl_kernels": len(nccl), "nccl_total_ms": nccl_total / 1e3, "nccl_overlapped_by_compute_ms": nccl_overlapped / 1e3,
            "nccl_avg_us": nccl_total / max(1, len(nccl)),
            "attn_self_avg_us_with_gather_in_flight": sum(att_busy) / max(1, len(att_busy)), "n_with": len(att_busy),
            "attn_self_avg_us_alone": sum(att_free) / max(1, len(att_free)), "n_alone": len(att_free),
            "top_kernels_ms": {k: [v[0], round(v[1] / 1e3, 2)] for k, v in top}}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
