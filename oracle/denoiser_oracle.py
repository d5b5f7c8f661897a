"""TEST INFRASTRUCTURE ONLY — fp32 CPU restatement of the reference's Stage-I denoising hot path.

Functional (state-dict driven) restatement of the arithmetic of facebookresearch/actionmesh @ 66db12e; every function
cites the reference file:line it follows.  It is pinned against the reference's own modules by
tests/test_oracle_golden.py (fixtures produced by oracle/gen_golden.py from the unmodified reference code on top of
oracle/diffusers_shim.py).  "diffusers semantics unpinned": the third-party diffusers classes (Attention, RMSNorm,
FP32LayerNorm, FeedForward, Timesteps, TimestepEmbedding) are restated from diffusers >= 0.30 published behaviour.

On a CPU-only torch the reference's `autocast(device_type="cuda")` is a no-op, so the reference CPU path — and this
oracle — is plain fp32.  Never imported by the product path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------------------ config
@dataclass
class DenoiserConfig:
    """Hyper-parameters; defaults = actionmesh/model/temporal_denoiser.py:29-49."""
    in_channels: int = 64
    num_layers: int = 21
    num_attention_heads: int = 16
    width: int = 2048
    mlp_ratio: float = 4.0
    cross_attention_dim: int = 1024
    inflated_layers: tuple = field(default_factory=lambda: tuple(range(21)))

    @property
    def head_dim(self) -> int:
        return self.width // self.num_attention_heads


# ------------------------------------------------------------------------------------------------------ small pieces
def layer_norm(x, w, b, eps=1e-5):
    """diffusers FP32LayerNorm / nn.LayerNorm (block.py:64,83,98,107; temporal_denoiser.py:108)."""
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps)


def rms_norm(x, w, eps=1e-6):
    """diffusers RMSNorm(dim_head, eps=1e-6, affine) used as norm_q / norm_k (block.py:72-74 -> Attention(qk_norm='rms_norm'))."""
    var = x.float().pow(2).mean(-1, keepdim=True)
    return x * torch.rsqrt(var + eps) * w


def timestep_embedding(t: torch.Tensor, channels: int) -> torch.Tensor:
    """diffusers Timesteps(num_channels, flip_sin_to_cos=False, downscale_freq_shift=0) (temporal_denoiser.py:57-61)."""
    half = channels // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    ang = t[:, None].float() * freqs[None]
    return torch.cat([ang.sin(), ang.cos()], dim=-1)


def rotary_tables(head_dim: int, positions: torch.Tensor):
    """actionmesh/model/utils/rotary_embedding.py:10-69: theta_j = 10000^(-2j/d); cos/sin repeated per pair."""
    inv = 1.0 / (10000.0 ** (torch.arange(0, head_dim, 2, dtype=torch.float32, device=positions.device) / head_dim))
    ph = torch.outer(positions.float(), inv)
    return ph.cos().repeat_interleave(2, dim=1), ph.sin().repeat_interleave(2, dim=1)


def apply_rotary(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """rotary_embedding.py:72-124: pairs (x0,x1) -> (x0 c - x1 s, x1 c + x0 s).  x:(B,H,S,D), cos/sin:(B,S,D)."""
    x0 = x[..., 0::2]
    x1 = x[..., 1::2]
    rot = torch.stack([-x1, x0], dim=-1).flatten(-2)
    return x.float() * cos[:, None] + rot.float() * sin[:, None]


def frame_positions(framestep: torch.Tensor) -> torch.Tensor:
    """embeddings.py:135-153 scale_timestep(center=True, scale=False): subtract the per-row minimum."""
    return framestep - framestep.min(dim=1, keepdim=True).values


# --------------------------------------------------------------------------------------------------------- attention
def _sdpa_torch(q, k, v):
    return F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)


def sdpa_exact_chunked(q, k, v, rows: int = 4096):
    """The same softmax(q k^T / sqrt(d)) v evaluated with explicit fp32 matmuls in query chunks, one head at a time: used
    when the oracle runs on a CUDA device, where torch's fused SDPA kernels for fp32 inputs may route through TF32 tensor
    cores (tests/test_default_config_gpu.py sets SDPA = sdpa_exact_chunked and disables TF32 matmuls)."""
    B, H, Sq, D = q.shape
    out = torch.empty(B, H, Sq, v.shape[-1], dtype=torch.float32, device=q.device)
    scale = 1.0 / math.sqrt(D)
    for b in range(B):
        for h in range(H):
            kt = k[b, h].float().t().contiguous()
            vf = v[b, h].float()
            for r0 in range(0, Sq, rows):
                s = (q[b, h, r0:r0 + rows].float() @ kt) * scale
                out[b, h, r0:r0 + rows] = torch.softmax(s, dim=-1) @ vf
    return out


SDPA = _sdpa_torch  # hook: the attention kernel the oracle calls (exact fp32 on a CPU torch)


def attention(sd: dict, prefix: str, x: torch.Tensor, heads: int, *, context: Optional[torch.Tensor] = None,
              inflate_frames: Optional[int] = None, rope=None) -> torch.Tensor:
    """actionmesh/model/utils/attention_processor.py:36-168 (AttentionProcessor.__call__).

    x: (B*T, L, D).  Self-attention with `inflate_frames=T` first reshapes to (B, T*L, D) (:49-65); RoPE tables
    `rope=(cos,sin)` of shape (B*T, L, d_h) are reshaped alike.  Head split follows :106-119 — heads are taken from the
    *concatenation* [q|k|v] (self) or [k|v] (cross), i.e. head h = columns [3*d*h, 3*d*(h+1)) of cat(q,k,v).
    """
    if inflate_frames is not None:
        bt, l, d = x.shape
        x = x.reshape(bt // inflate_frames, inflate_frames * l, d)
        if rope is not None:
            rope = tuple(r.reshape(bt // inflate_frames, inflate_frames * l, -1) for r in rope)
    b = x.shape[0]
    q = x @ sd[prefix + "to_q.weight"].t()
    src = x if context is None else context
    k = src @ sd[prefix + "to_k.weight"].t()
    v = src @ sd[prefix + "to_v.weight"].t()
    if context is None:
        qkv = torch.cat([q, k, v], dim=-1)
        dh = qkv.shape[-1] // heads // 3
        q, k, v = qkv.view(b, -1, heads, 3 * dh).split(dh, dim=-1)
    else:
        kv = torch.cat([k, v], dim=-1)
        dh = kv.shape[-1] // heads // 2
        k, v = kv.view(b, -1, heads, 2 * dh).split(dh, dim=-1)
        q = q.view(b, -1, heads, dh)
    q, k, v = (t.transpose(1, 2) for t in (q, k, v))  # (B,H,S,dh)
    q = rms_norm(q, sd[prefix + "norm_q.weight"])
    k = rms_norm(k, sd[prefix + "norm_k.weight"])
    if rope is not None:
        q = apply_rotary(q, *rope)
        k = apply_rotary(k, *rope)
    o = SDPA(q, k, v)  # :133-139 F.scaled_dot_product_attention(dropout_p=0, is_causal=False), scale 1/sqrt(dh)
    o = o.transpose(1, 2).reshape(b, -1, heads * dh)
    o = o @ sd[prefix + "to_out.0.weight"].t() + sd[prefix + "to_out.0.bias"]
    if inflate_frames is not None:
        o = o.reshape(b * inflate_frames, -1, o.shape[-1])
    return o


# ------------------------------------------------------------------------------------------------------------- block
def block_forward(sd: dict, prefix: str, h: torch.Tensor, context: torch.Tensor, heads: int, n_frames: int,
                  inflate: bool, rope, skip: Optional[torch.Tensor]) -> torch.Tensor:
    """actionmesh/model/utils/block.py:110-154 (FlowMatchingBlock.forward)."""
    if skip is not None:  # :131-133
        cat = torch.cat([skip, h], dim=-1)
        h = cat @ sd[prefix + "linear_skip.weight"].t() + sd[prefix + "linear_skip.bias"]
        h = layer_norm(h, sd[prefix + "norm_skip.weight"], sd[prefix + "norm_skip.bias"])
    hn = layer_norm(h, sd[prefix + "norm_s_attn.weight"], sd[prefix + "norm_s_attn.bias"])
    h = h + attention(sd, prefix + "s_attn.", hn, heads, inflate_frames=n_frames if inflate else None, rope=rope)  # :137-142
    hn = layer_norm(h, sd[prefix + "norm_x_attn.weight"], sd[prefix + "norm_x_attn.bias"])
    h = h + attention(sd, prefix + "x_attn.", hn, heads, context=context)  # :146-149
    hn = layer_norm(h, sd[prefix + "norm_ff.weight"], sd[prefix + "norm_ff.bias"])
    ff = F.gelu(hn @ sd[prefix + "ff.net.0.proj.weight"].t() + sd[prefix + "ff.net.0.proj.bias"])  # exact erf GELU
    h = h + (ff @ sd[prefix + "ff.net.2.weight"].t() + sd[prefix + "ff.net.2.bias"])  # :152
    return h


# ---------------------------------------------------------------------------------------------------------- denoiser
def denoiser_forward(sd: dict, cfg: DenoiserConfig, hidden_states: torch.Tensor, context: torch.Tensor,
                     framestep: torch.Tensor, diffusion_time: torch.Tensor, mask: Optional[torch.Tensor] = None,
                     freqs_rot=None):
    """actionmesh/model/temporal_denoiser.py:151-249 (ActionMeshDenoiser.forward).  Returns (out (B,T,N,C), freqs_rot)."""
    B, T, N, _ = hidden_states.shape
    if freqs_rot is None:  # :114-149 precompute_freqs_rot
        pos = frame_positions(framestep).reshape(B * T)
        cos, sin = rotary_tables(cfg.head_dim, pos)
        freqs_rot = (cos[:, None].repeat(1, N + 1, 1), sin[:, None].repeat(1, N + 1, 1))
    h = hidden_states.reshape(B * T, N, -1).float() @ sd["proj_in.weight"].t() + sd["proj_in.bias"]  # :205-206
    dt = diffusion_time.repeat(T)  # :209  (tiles [b0,b1,b0,b1,...]; all CFG branches share t)
    if mask is not None:
        dt = dt * (1 - mask.reshape(B * T))  # :210-212 observed frames get t = 0
    emb = timestep_embedding(dt, cfg.width)  # :213
    emb = F.gelu(emb @ sd["time_proj.linear_1.weight"].t() + sd["time_proj.linear_1.bias"])
    emb = emb @ sd["time_proj.linear_2.weight"].t() + sd["time_proj.linear_2.bias"]  # :214
    h = torch.cat([emb[:, None], h], dim=1)  # :217 time token first
    ctx = context.reshape(B * T, context.shape[2], context.shape[3]).float()
    skips = []
    half = cfg.num_layers // 2
    for layer in range(cfg.num_layers):  # :222-236
        skip = None if layer <= half else skips.pop()
        h = block_forward(sd, f"blocks.{layer}.", h, ctx, cfg.num_attention_heads, T, layer in cfg.inflated_layers,
                          freqs_rot, skip)
        if layer < half:
            skips.append(h)
    h = layer_norm(h, sd["norm_out.weight"], sd["norm_out.bias"])  # :239
    h = h[:, -N:] @ sd["proj_out.weight"].t() + sd["proj_out.bias"]  # :241-242
    return h.reshape(B, T, N, -1), freqs_rot


class OracleDenoiser:
    """Duck-type of ActionMeshDenoiser for SchedulerFlow (scheduler.py:151-158): `.forward(hidden_states=, context=, ...)`."""

    def __init__(self, state_dict: dict, cfg: DenoiserConfig, device="cpu"):
        self.device = torch.device(device)
        self.sd = {k: v.detach().float().to(self.device) for k, v in state_dict.items()}
        self.cfg = cfg

    @torch.no_grad()
    def forward(self, hidden_states, context, framestep, diffusion_time, mask=None, freqs_rot=None):
        d = self.device
        return denoiser_forward(self.sd, self.cfg, hidden_states.to(d), context.to(d), framestep.to(d),
                                diffusion_time.to(d), None if mask is None else mask.to(d), freqs_rot)


# --------------------------------------------------------------------------------------------------------- scheduler
def flow_schedule(num_inference_steps: int, num_train_timesteps: int = 1000, shift: float = 3.0):
    """actionmesh/scheduler/scheduler.py:43-98: shifted-linear sigma schedule -> (timesteps (n+1), distances (n)) fp32."""
    n = num_inference_steps + 1
    full = np.linspace(1, num_train_timesteps, num_train_timesteps)[::-1] / num_train_timesteps
    full = shift * full / (1 + (shift - 1) * full)
    ts = np.linspace(full[0] * num_train_timesteps, full[-1] * num_train_timesteps, n)
    sig = ts / num_train_timesteps
    sig = shift * sig / (1 + (shift - 1) * sig)
    timesteps = torch.from_numpy((sig * num_train_timesteps).astype(np.float32))
    distances = (timesteps[:-1] - timesteps[1:]) / num_train_timesteps
    return timesteps, distances


def flow_noise(latent_shape, batch_size: int, n_timesteps: int, generator=None, corr_noise: float = 0.0, device="cpu"):
    """scheduler.py:100-137: two randn draws in this order (shared, independent); corr_noise=0 still advances the stream."""
    same = torch.randn([batch_size, 1] + list(latent_shape), generator=generator, device=device).repeat(1, n_timesteps, 1, 1)
    indep = torch.randn([batch_size, n_timesteps] + list(latent_shape), generator=generator, device=device)
    return math.sqrt(corr_noise) * same + math.sqrt(1 - corr_noise) * indep


def cfg_batch(latent, context, mask, framestep, guidance_at_inference):
    """actionmesh/scheduler/guidance.py:38-93 (cfg_at_inference)."""
    k = len(guidance_at_inference)
    lat = torch.cat([latent] * k)
    fs = torch.cat([framestep] * k) if framestep is not None else None
    ctxs, masks = [], []
    for use_img, use_lat in guidance_at_inference:
        ctxs.append(context if use_img else torch.zeros_like(context))
        if mask is not None:
            masks.append(mask if use_lat else torch.zeros_like(mask))
    return lat, torch.cat(ctxs, 0), (torch.cat(masks, 0) if mask is not None else None), fs


def cfg_aggregate(pred, guidance_scales, n_branches):
    """guidance.py:95-118 (aggregate_cfg): out = p0 + sum_i s_i (p_{i+1} - p_i)."""
    parts = pred.chunk(n_branches, dim=0)
    out = parts[0].clone()
    for i in range(n_branches - 1):
        out = out + guidance_scales[i] * (parts[i + 1] - parts[i])
    return out


@torch.no_grad()
def flow_denoise(model, init_latent, context, mask, framestep, *, num_inference_steps: int, guidance_scales,
                 guidance_at_inference=((0, 1), (1, 1)), shift: float = 3.0, is_additive: bool = True,
                 step_callback: Optional[Callable] = None):
    """scheduler.py:172-295 (_flow_sample + denoise), default actionmesh.yaml scheduler/cf_guidance settings."""
    latents = init_latent.clone()
    timesteps, distances = flow_schedule(num_inference_steps, shift=shift)
    distances = distances.to(latents.device)
    unobserved = (mask == 0) if mask is not None else None
    freqs_rot = None
    for i, t in enumerate(timesteps[:-1]):
        h_in, c_in, m_in, f_in = cfg_batch(latents, context, mask, framestep, guidance_at_inference)
        dtime = torch.tensor([float(t)], dtype=latents.dtype, device=latents.device).expand(h_in.shape[0])
        pred, freqs_rot = model.forward(hidden_states=h_in, context=c_in, framestep=f_in, mask=m_in,
                                        diffusion_time=dtime, freqs_rot=freqs_rot)
        pred = cfg_aggregate(pred, guidance_scales, len(guidance_at_inference))
        step = latents + distances[i] * pred if is_additive else latents - distances[i] * pred
        if unobserved is not None:
            latents[unobserved] = step[unobserved]
        else:
            latents = step
        if step_callback is not None:
            step_callback(i + 1, num_inference_steps)
    return latents


# ------------------------------------------------------------------------------------------------- windows and bank
def _chunk_right(start: int, end: int, size: int, slide: int):
    """actionmesh/model/utils/timesteps.py:10-48 (chunk_right): the right edge advances (first to start+size, then by
    `slide`, clamped to `end`); each window is the `size` indices left of the edge, clamped to `start`."""
    out = []
    edge = start
    while edge < end:
        edge = min(start + size, end) if not out else min(edge + slide, end)
        out.append(torch.arange(max(start, edge - size), edge))
    return out


def _chunk_left(start: int, end: int, size: int, slide: int):
    """timesteps.py:51-74 (chunk_left): chunk_right's windows in reverse order, each with descending indices."""
    return [c.flip(0) for c in reversed(_chunk_right(start, end, size, slide))]


def chunk_from(start: int, total: int, size: int, slide: int):
    """timesteps.py:77-117: AR window partition expanding from the anchor index in both directions."""
    context = size - slide
    if total == size:
        idx = torch.arange(total)
        return [torch.cat([idx[start:start + 1], idx[idx != start]])]
    if start == 0:
        return _chunk_right(0, total, size, slide)
    if start == total - 1:
        return _chunk_left(0, total, size, slide)
    if start > total - start:
        left = _chunk_left(0, start + 1, size, slide)
        right_start = min(max(0, start - context + 1), total - size)
        return left + _chunk_right(right_start, total, size, slide)
    right = _chunk_right(start, total, size, slide)
    left_end = max(min(start + context, total), size)
    return right + _chunk_left(0, left_end, size, slide)


class LatentBank:
    """actionmesh/model/utils/storage.py:20-183: per-frame latents keyed by float timestep (eps 1e-5); `get` returns
    (stacked latents, int32 mask) with zeros for missing frames; `update` never overwrites unless replace=True."""

    def __init__(self, empty_dims=(2048, 64)):
        self.empty_dims = tuple(empty_dims)
        self.items: list[torch.Tensor] = []
        self.timesteps: list[float] = []

    def _index(self, t: float, eps: float = 1e-5):
        for i, ts in enumerate(self.timesteps):
            if abs(ts - t) < eps:
                return i
        return None

    def update(self, timesteps: torch.Tensor, latents: torch.Tensor, replace: bool = False) -> None:
        ts = timesteps.flatten()
        lat = latents.reshape(ts.shape[0], *self.empty_dims)
        for i in range(ts.shape[0]):
            t = ts[i].item()
            idx = self._index(t)
            if idx is None:
                self.timesteps.append(t)
                self.items.append(lat[i])
            elif replace:
                self.items[idx] = lat[i]

    def get(self, timesteps: torch.Tensor, device="cpu", add_batch_dim: bool = False):
        lat, msk = [], []
        for t in timesteps:
            idx = self._index(float(t))
            if idx is None:
                lat.append(torch.zeros(self.empty_dims, dtype=torch.float32, device=device))
                msk.append(0)
            else:
                lat.append(self.items[idx].to(device))
                msk.append(1)
        lat = torch.stack(lat)
        msk = torch.tensor(msk, dtype=torch.int32, device=device)
        return (lat[None], msk[None]) if add_batch_dim else (lat, msk)

    def get_ordered(self):
        order = sorted(range(len(self.timesteps)), key=lambda i: self.timesteps[i])
        lat = torch.stack([self.items[i] for i in order])
        return lat, torch.tensor([self.timesteps[i] for i in order]).to(lat)


@torch.no_grad()
def generate_3d_latents(model, context, timesteps, bank: LatentBank, *, anchor_idx: int = 0, seed: int = 44,
                        window: int = 16, slide: int = 15, latent_shape=(2048, 64), noise_fn=None, **denoise_kw):
    """actionmesh/pipeline.py:435-508 (generate_3d_latents) + :247-314 (_denoise_latents): serial AR windows, seed+i per
    window, init = cond*mask + noise*(1-mask).  `noise_fn(window_idx, T)` may supply the noise (CPU/CUDA generators
    differ for the same seed) — default draws it from a CPU generator like the reference on a CPU device."""
    n_frames = timesteps.shape[0]
    for i, idx in enumerate(chunk_from(anchor_idx, n_frames, window, slide)):
        ts = timesteps[idx]
        cond, cmask = bank.get(ts, add_batch_dim=True)
        if noise_fn is not None:
            noise = noise_fn(i, len(idx))
        else:
            gen = torch.Generator(device="cpu").manual_seed(seed + i)
            noise = flow_noise(list(latent_shape), 1, len(idx), generator=gen)
        m = cmask[..., None, None].float()
        init = cond * m + noise * (1.0 - m)
        lat = flow_denoise(model, init, context[idx][None], cmask.float(), ts[None], **denoise_kw)
        bank.update(ts, lat)
    return bank
