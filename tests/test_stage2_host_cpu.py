"""CPU tests of the Stage-II orchestration (no GPU): the time bookkeeping helpers against golden values written by the
reference's own embeddings.py functions, and AnimationPipeline.generate_mesh_animation's window / alpha plumbing against
the oracle restatement of pipeline.py:510-600, both driving the same recording fake decoder."""
import torch

from actionmesh_b200 import windows as W
from actionmesh_b200.pipeline import AnimationPipeline
from actionmesh_b200.windows import LatentBank, VertexBank
from conftest import load_golden
from oracle import autoencoder_oracle as ao


def test_time_helpers_match_reference_golden():
    g = load_golden("stage2_host_logic.pt")
    for (a, b, l), n in g["n_subdivisions"].items():
        assert W.get_n_subdivisions(a, b, l) == n == ao.get_n_subdivisions(a, b, l)
    names = {"w16": torch.arange(16.0)[None], "w16_off": torch.arange(15.0, 31.0)[None],
             "w5": torch.tensor([[2.0, 3.0, 4.0, 5.0, 6.0]])}
    for (name, lvl, df), ref in g["interp"].items():
        assert torch.equal(W.interpolate_timesteps(names[name], lvl, drop_first=df), ref)
        assert torch.equal(ao.interpolate_timesteps(names[name], lvl, drop_first=df), ref)
    for name, (ts, t_min, t_range, a1, a2) in g["scaling"].items():
        m, r = W.get_scaling(ts)
        assert torch.equal(m, t_min) and torch.equal(r, t_range)
        assert torch.equal(W.apply_scaling(ts[:, 0], m, r), a1) and torch.equal(W.apply_scaling(ts, m, r), a2)


class _FakeAE:
    """Duck type of B200Autoencoder that records its calls and returns a displacement depending on all its inputs."""

    class config:
        temporal_context_size = 16

    device = torch.device("cpu")

    def __init__(self):
        self.calls = []

    def __call__(self, latent, framestep, source_alpha, target_alphas, query, step_callback=None):
        self.calls.append((latent.clone(), framestep.clone(), source_alpha.clone(), target_alphas.clone(), query.clone()))
        base = query[..., :3][:, None] * (1.0 - 0.1 * target_alphas[..., None, None])
        return (base + 0.01 * latent.mean() + 0.001 * framestep.sum()).clamp(-1, 1)

    def apply_displacement(self, vertex, displacement, scale=1.0):
        return displacement.clamp(-scale, scale)


def _run(n_frames, anchor_idx):
    g = torch.Generator().manual_seed(n_frames)
    lat = torch.randn(n_frames, 6, 4, generator=g)
    ts = torch.arange(n_frames, dtype=torch.float32) + 3.0
    verts = torch.rand(11, 3, generator=g) - 0.5
    nrm = torch.nn.functional.normalize(torch.randn(11, 3, generator=g), dim=-1)
    nfn = lambda v: torch.nn.functional.normalize(v + 0.3, dim=-1)
    fake = _FakeAE()
    pipe = AnimationPipeline.__new__(AnimationPipeline)
    pipe.temporal_3D_vae, pipe.anchor_idx = fake, anchor_idx
    pipe.sliding_window_autoencoder, pipe.subsampling_level, pipe.normals_fn = 15, 1, nfn
    pipe.temporal_3D_denoiser = fake                          # .device
    bank = LatentBank(empty_dims=(6, 4))
    bank.update(timesteps=ts, latents=lat)
    vb = VertexBank()
    vb.update(timesteps=ts[anchor_idx:anchor_idx + 1], vertices=[verts])
    out = pipe.generate_mesh_animation(bank, vb, nrm)
    got_v, got_t = out.get_ordered()
    fake2 = _FakeAE()
    ref_t, ref_v = ao.generate_mesh_animation(
        lambda latent, framestep, source_alpha, target_alphas, query: fake2(latent, framestep, source_alpha, target_alphas, query),
        lat, ts, verts, nrm, anchor_idx=anchor_idx, normals_fn=nfn)
    assert [round(t, 5) for t in got_t.tolist()] == [round(t, 5) for t in ref_t]
    assert len(fake.calls) == len(fake2.calls)
    for a, b in zip(fake.calls, fake2.calls):
        for x, y in zip(a, b):
            assert torch.allclose(x, y, atol=0, rtol=0), (x, y)
    for a, b in zip(got_v, ref_v):
        assert torch.equal(a, b)
    return len(fake.calls), got_t


def test_single_window_plumbing():
    n, ts = _run(16, 0)
    assert n == 1 and len(ts) == 16


def test_multi_window_autoregressive_plumbing():
    n, ts = _run(31, 0)
    assert n == 2 and len(ts) == 31
    n, ts = _run(47, 0)
    assert n == len(W.chunk_from(0, 47, 16, 15)) and len(ts) == 47
    n, ts = _run(40, 20)                                     # anchor mid-clip: right chunks then left chunks
    assert n == len(W.chunk_from(20, 40, 16, 15))
