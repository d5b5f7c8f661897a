"""smoke(): one tiny denoiser step on cuda:0 checked against the fp32 oracle (test infrastructure: lives outside the
product package; __graft_entry__.smoke() calls it)."""
from __future__ import annotations

import os
import sys

import torch


def smoke_check(verbose: bool = True) -> float:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from actionmesh_b200.denoiser import B200Denoiser, DenoiserConfig
    from actionmesh_b200.guidance import ClassifierFreeGuidance
    from actionmesh_b200.scheduler import B200SchedulerFlow
    from oracle import denoiser_oracle as do
    from oracle import synth

    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs cuda:0")
    torch.cuda.set_device(0)
    d = dict(num_layers=3, num_attention_heads=2, width=256, cross_attention_dim=128, in_channels=64, mlp_ratio=4.0)
    cfg = DenoiserConfig(inflated_layers=(0, 1, 2), **d)
    sd = synth.make_state_dict(cfg, 7)
    model = B200Denoiser(cfg).to("cuda:0")
    model.load_state_dict(sd)
    lat, ctx, fs, mask = synth.make_inputs(1, 3, 31, 64, 9, 128, seed=3)
    sch = B200SchedulerFlow(num_inference_steps=1, shift=3.0, is_additive=True)
    cf = ClassifierFreeGuidance(guidance_at_inference=[[0, 1], [1, 1]], guidance_scales=[7.5])
    out = sch.denoise(model, cf, lat.clone().cuda(), ctx.cuda(), device="cuda:0", mask=mask.cuda(), framestep=fs).cpu()
    ocfg = do.DenoiserConfig(inflated_layers=(0, 1, 2), **d)
    ref = do.flow_denoise(do.OracleDenoiser(sd, ocfg), lat, ctx, mask, fs, num_inference_steps=1, guidance_scales=[7.5])
    err = float((out - ref).norm() / ref.norm())
    if verbose:
        print(f"smoke: 1 denoiser step (CFG x2, 3 blocks) on cuda:0 vs fp32 oracle: rel err {err:.3e}")
    if not (err < 4e-2) or not torch.equal(out[0, 0], lat[0, 0]):
        raise RuntimeError(f"smoke parity failed: rel err {err}")
    return err
