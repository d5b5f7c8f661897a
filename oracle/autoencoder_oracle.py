"""TEST INFRASTRUCTURE ONLY — fp32 CPU restatement of the reference's Stage-II temporal autoencoder and of the ActionBench
Chamfer metric: (1) the checker of the CUDA Stage-II path (actionmesh_b200/autoencoder.py, tests/test_autoencoder_gpu.py),
(2) the common fp32 decoder that turns Stage-I latents of both paths into per-frame vertices so that Stage-I parity can be
reported as the Chamfer distance `north_star` names (tests/test_chamfer_gpu.py).

  autoencoder_forward : actionmesh/model/temporal_autoencoder.py:163-269 (ActionMeshAutoencoder.forward)
  chamfer_score       : actionbench/chamfer.py:12-50 (compute_chamfer_score; scipy KD-tree, seeded sub-sampling)
Pinned against the reference's own module by tests/test_oracle_golden.py (fixture from oracle/gen_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from . import denoiser_oracle as do


@dataclass
class AutoencoderConfig:
    """Defaults = temporal_autoencoder.py:38-57."""
    in_channels: int = 3
    in_extra_channels: int = 3
    out_dim: int = 3
    latent_channels: int = 64
    width: int = 1024
    num_layers: int = 16
    num_attention_heads: int = 8
    embed_frequency: int = 8
    embed_include_pi: bool = False
    prediction_mode: str = "direct"

    @property
    def head_dim(self):
        return self.width // self.num_attention_heads


def frequency_embedding(x: torch.Tensor, num_freqs: int, include_pi: bool) -> torch.Tensor:
    """embeddings.py:15-53 FrequencyPositionalEmbedding(logspace=True, include_input=True): [x | sin(x f) | cos(x f)]."""
    freqs = 2.0 ** torch.arange(num_freqs, dtype=torch.float32)
    if include_pi:
        freqs = freqs * math.pi
    emb = (x[..., None] * freqs).reshape(*x.shape[:-1], -1)
    return torch.cat([x, emb.sin(), emb.cos()], dim=-1)


def alpha_embedding(size: int, *ts: torch.Tensor) -> torch.Tensor:
    """embeddings.py:56-132 TimestepEmbedder(frequency_embedding_size=size): per input [cos(t w) | sin(t w)], concatenated."""
    half = size // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    out = []
    for t in ts:
        a = t[..., None].float() * freqs
        out += [a.cos(), a.sin()]
    return torch.cat(out, dim=-1)


def _self_block(sd, prefix, h, heads, rope):
    """FlowMatchingBlock(use_cross_attention=False, attention_qk_norm=None) — block.py:136-152 with the processor's
    head-interleaved split (attention_processor.py:106-110), RoPE, no q/k norm, not inflated."""
    hn = do.layer_norm(h, sd[prefix + "norm_s_attn.weight"], sd[prefix + "norm_s_attn.bias"])
    b = hn.shape[0]
    q = hn @ sd[prefix + "s_attn.to_q.weight"].t()
    k = hn @ sd[prefix + "s_attn.to_k.weight"].t()
    v = hn @ sd[prefix + "s_attn.to_v.weight"].t()
    qkv = torch.cat([q, k, v], dim=-1)
    dh = qkv.shape[-1] // heads // 3
    q, k, v = (t.transpose(1, 2) for t in qkv.view(b, -1, heads, 3 * dh).split(dh, dim=-1))
    q, k = do.apply_rotary(q, *rope), do.apply_rotary(k, *rope)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, -1, heads * dh)
    h = h + (o @ sd[prefix + "s_attn.to_out.0.weight"].t() + sd[prefix + "s_attn.to_out.0.bias"])
    hn = do.layer_norm(h, sd[prefix + "norm_ff.weight"], sd[prefix + "norm_ff.bias"])
    ff = F.gelu(hn @ sd[prefix + "ff.net.0.proj.weight"].t() + sd[prefix + "ff.net.0.proj.bias"])
    return h + (ff @ sd[prefix + "ff.net.2.weight"].t() + sd[prefix + "ff.net.2.bias"])


def _cross_block(sd, prefix, xq, kv, heads):
    """FlowMatchingBlock(use_self_attention=False, cross_attention_norm_type='layer_norm', attention_qk_norm=None)."""
    hn = do.layer_norm(xq, sd[prefix + "norm_x_attn.weight"], sd[prefix + "norm_x_attn.bias"])
    ctx = F.layer_norm(kv, (kv.shape[-1],), sd[prefix + "x_attn.norm_cross.weight"], sd[prefix + "x_attn.norm_cross.bias"])
    b = hn.shape[0]
    q = hn @ sd[prefix + "x_attn.to_q.weight"].t()
    k = ctx @ sd[prefix + "x_attn.to_k.weight"].t()
    v = ctx @ sd[prefix + "x_attn.to_v.weight"].t()
    kvc = torch.cat([k, v], dim=-1)
    dh = kvc.shape[-1] // heads // 2
    k, v = (t.transpose(1, 2) for t in kvc.view(b, -1, heads, 2 * dh).split(dh, dim=-1))
    q = q.view(b, -1, heads, dh).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, -1, heads * dh)
    h = xq + (o @ sd[prefix + "x_attn.to_out.0.weight"].t() + sd[prefix + "x_attn.to_out.0.bias"])
    hn = do.layer_norm(h, sd[prefix + "norm_ff.weight"], sd[prefix + "norm_ff.bias"])
    ff = F.gelu(hn @ sd[prefix + "ff.net.0.proj.weight"].t() + sd[prefix + "ff.net.0.proj.bias"])
    return h + (ff @ sd[prefix + "ff.net.2.weight"].t() + sd[prefix + "ff.net.2.bias"])


@torch.no_grad()
def autoencoder_forward(sd: dict, cfg: AutoencoderConfig, latent, framestep, source_alpha, target_alphas, query):
    """temporal_autoencoder.py:163-269.  latent (B,T,N,C), framestep (B,T), source_alpha (B), target_alphas (B,T_out),
    query (B,V,3|6) -> displacement field (B,T_out,V,3) in [-1,1]."""
    B, T, N, _ = latent.shape
    T_out = target_alphas.shape[1]
    pos = do.frame_positions(framestep).reshape(B * T)
    cos, sin = do.rotary_tables(cfg.head_dim, pos)
    cos, sin = cos.reshape(B, T, -1), sin.reshape(B, T, -1)
    cos = torch.cat([cos.repeat_interleave(N, dim=1), cos], dim=1)  # (B, T*N + T, d_h): latent tokens then alpha tokens
    sin = torch.cat([sin.repeat_interleave(N, dim=1), sin], dim=1)
    lat = (latent.float() @ sd["post_quant.weight"].t() + sd["post_quant.bias"]).reshape(B, T * N, -1)
    src = source_alpha[:, None].expand_as(target_alphas)
    alpha = alpha_embedding(cfg.width // 2, src, target_alphas)  # (B, T_out, width)
    qe = frequency_embedding(query[..., :3].float(), cfg.embed_frequency, cfg.embed_include_pi)
    if cfg.in_extra_channels > 0:
        qe = torch.cat([qe, query[..., 3:].float()], dim=-1)
    out = torch.empty(B, T_out, query.shape[1], cfg.out_dim)
    for i in range(T_out):
        h = torch.cat([lat, alpha[:, i][:, None].repeat(1, T, 1)], dim=1)
        for l in range(cfg.num_layers):
            h = _self_block(sd, f"blocks.{l}.", h, cfg.num_attention_heads, (cos, sin))
        q = qe @ sd["proj_query.weight"].t() + sd["proj_query.bias"]
        o = _cross_block(sd, f"blocks.{cfg.num_layers}.", q, h, cfg.num_attention_heads)
        o = F.layer_norm(o, (o.shape[-1],), sd["norm_out.weight"], sd["norm_out.bias"])
        out[:, i] = -(o @ sd["proj_out.weight"].t() + sd["proj_out.bias"])
    return 2 * torch.sigmoid(out) - 1.0


def apply_displacement(vertex, displacement, mode="direct", scale=1.0):
    """temporal_autoencoder.py:118-141."""
    if mode == "direct":
        return displacement.clamp(-scale, scale)
    return (vertex[:, None] + displacement).clamp(-scale, scale)


# ---- Stage-II orchestration (pipeline.py:510-600 generate_mesh_animation + :316-385 _decode_displacement) ------------------
def get_n_subdivisions(start, end, level=1):
    """embeddings.py:199-214."""
    n = int(end - start + 1)
    for _ in range(1, level):
        n += n - 1
    return n


def interpolate_timesteps(timesteps, subsampling_level, drop_first=False):
    """embeddings.py:217-242."""
    t_min, t_max = timesteps.min().item(), timesteps.max().item()
    out = torch.linspace(t_min, t_max, get_n_subdivisions(t_min, t_max, subsampling_level)).reshape(1, -1)
    return out[:, 1:] if drop_first else out


def generate_mesh_animation(decode, latents, timesteps, anchor_vertices, anchor_normals, *, anchor_idx=0, context_size=16,
                            slide=15, subsampling_level=1, normals_fn=None):
    """pipeline.py:510-600 with the mesh bank reduced to {timestep: (V,3) vertices} (all meshes share the anchor faces).
    `decode(latent (1,T,N,C), framestep (1,T), source_alpha (1,), target_alphas (1,T_out), query (1,V,6)) -> (1,T_out,V,3)`
    displacement; returns (sorted timesteps, [vertices])."""
    order = torch.argsort(timesteps)
    all_ts, lat_sorted = timesteps[order], latents[order]
    bank = {float(timesteps[anchor_idx]): anchor_vertices}
    for idx in do.chunk_from(anchor_idx, len(all_ts), context_size, slide):
        wts = all_ts[idx][None]
        verts = bank[float(wts[0, 0])]
        first = abs(float(wts[0, 0]) - float(timesteps[anchor_idx])) < 1e-5
        normals = anchor_normals if first else normals_fn(verts)
        out_ts = interpolate_timesteps(wts, subsampling_level, drop_first=True)
        t_min = wts.min(dim=1).values
        t_range = wts.max(dim=1).values - t_min                                      # embeddings.py:156-196
        src = (wts[:, 0] - t_min) / t_range
        tgt = (out_ts - t_min[:, None]) / t_range[:, None]
        q = torch.cat([verts, normals], dim=-1)[None]
        v = apply_displacement(verts[None], decode(lat_sorted[idx][None], wts, src, tgt, q))[0]
        for j, t in enumerate(out_ts[0].tolist()):
            if not any(abs(t - k) < 1e-5 for k in bank):
                bank[t] = v[j]
    ts = sorted(bank)
    return ts, [bank[t] for t in ts]


def chamfer_score(pred, gt, n: int = 10_000, seed: int = 44) -> float:
    """actionbench/chamfer.py:12-50: symmetric Chamfer distance, seeded sub-sampling of the query sets, KD-tree NN."""
    from scipy.spatial import KDTree

    pred, gt = np.asarray(pred, dtype=np.float64), np.asarray(gt, dtype=np.float64)
    ip = np.random.RandomState(seed=seed).permutation(len(pred))[:n] if 0 < n < len(pred) else np.arange(len(pred))
    ig = np.random.RandomState(seed=seed + 1).permutation(len(gt))[:n] if 0 < n < len(gt) else np.arange(len(gt))
    d1, _ = KDTree(pred).query(gt[ig])
    d2, _ = KDTree(gt).query(pred[ip])
    return float(np.mean(d1) + np.mean(d2))


def make_autoencoder_state_dict(cfg: AutoencoderConfig, seed: int = 4321) -> dict:
    """Seeded, bf16-representable synthetic weights with the reference's state-dict keys (same role as synth.make_state_dict)."""
    g = torch.Generator().manual_seed(seed)
    D = cfg.width
    rs = 1.0 / math.sqrt(cfg.num_layers + 1)

    def lin(o, i, s=1.0):
        b = 1.0 / math.sqrt(i)
        return ((torch.rand(o, i, generator=g) * 2 - 1) * b * s).to(torch.bfloat16).float()

    def vec(n, lo, hi):
        return (torch.rand(n, generator=g) * (hi - lo) + lo).to(torch.bfloat16).float()

    qdim = cfg.in_channels * (2 * cfg.embed_frequency + 1) + cfg.in_extra_channels
    sd = {"post_quant.weight": lin(D, cfg.latent_channels), "post_quant.bias": vec(D, -0.1, 0.1),
          "proj_query.weight": lin(D, qdim), "proj_query.bias": vec(D, -0.1, 0.1),
          "norm_out.weight": vec(D, 0.8, 1.2), "norm_out.bias": vec(D, -0.1, 0.1),
          "proj_out.weight": lin(cfg.out_dim, D), "proj_out.bias": vec(cfg.out_dim, -0.1, 0.1)}
    for l in range(cfg.num_layers):
        p = f"blocks.{l}."
        for n in ("norm_s_attn", "norm_ff"):
            sd[p + n + ".weight"], sd[p + n + ".bias"] = vec(D, 0.8, 1.2), vec(D, -0.1, 0.1)
        for n in ("to_q", "to_k", "to_v"):
            sd[p + f"s_attn.{n}.weight"] = lin(D, D)
        sd[p + "s_attn.to_out.0.weight"], sd[p + "s_attn.to_out.0.bias"] = lin(D, D, rs), vec(D, -0.02, 0.02)
        sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"] = lin(4 * D, D), vec(4 * D, -0.02, 0.02)
        sd[p + "ff.net.2.weight"], sd[p + "ff.net.2.bias"] = lin(D, 4 * D, rs), vec(D, -0.02, 0.02)
    p = f"blocks.{cfg.num_layers}."
    for n in ("norm_x_attn", "norm_ff"):
        sd[p + n + ".weight"], sd[p + n + ".bias"] = vec(D, 0.8, 1.2), vec(D, -0.1, 0.1)
    sd[p + "x_attn.norm_cross.weight"], sd[p + "x_attn.norm_cross.bias"] = vec(D, 0.8, 1.2), vec(D, -0.1, 0.1)
    for n in ("to_q", "to_k", "to_v"):
        sd[p + f"x_attn.{n}.weight"] = lin(D, D)
    sd[p + "x_attn.to_out.0.weight"], sd[p + "x_attn.to_out.0.bias"] = lin(D, D, rs), vec(D, -0.02, 0.02)
    sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"] = lin(4 * D, D), vec(4 * D, -0.02, 0.02)
    sd[p + "ff.net.2.weight"], sd[p + "ff.net.2.bias"] = lin(D, 4 * D, rs), vec(D, -0.02, 0.02)
    return sd
