"""TEST INFRASTRUCTURE ONLY — fp32 restatement of the Stage-0 (TripoSG) denoising loop the reference runs before Stage I
(actionmesh/pipeline.py:387-433 -> third_party/TripoSG @ fc5c409).

TripoSG's DiT (triposg/models/transformers/triposg_transformer.py:129-362,365-726) is the block family ActionMesh's
denoiser was derived from: a single-frame (T = 1) forward without rotary embedding, the same time token, the same
head-interleaved q/k/v split (triposg/models/attention_processor.py:122-133), the same long skips
(`skip_concat_front=True`, `skip_norm_last=True`).  So the restatement is a state-dict key mapping onto
oracle/denoiser_oracle.py plus the sampler: RectifiedFlowScheduler (triposg/schedulers/scheduling_rectified_flow.py:
177-215 set_timesteps, :234-308 step) inside TripoSGPipeline.__call__ (triposg/pipelines/pipeline_triposg.py:243-294).
Pinned against the reference's own modules by tests/test_oracle_golden.py (fixture tests/golden/triposg_tiny.pt)."""
from __future__ import annotations

import numpy as np
import torch

from . import denoiser_oracle as do

_BLOCK_KEYS = (("norm1.", "norm_s_attn."), ("attn1.", "s_attn."), ("norm2.", "norm_x_attn."), ("attn2.", "x_attn."),
               ("norm3.", "norm_ff."), ("skip_linear.", "linear_skip."), ("skip_norm.", "norm_skip."))


def remap_state_dict(sd: dict) -> dict:
    """TripoSGDiTModel keys (DiTBlock: norm1/attn1/norm2/attn2/norm3/ff/skip_linear/skip_norm, triposg_transformer.py:
    190-262) -> ActionMeshDenoiser keys (block.py:64-108)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("blocks."):
            _, idx, rest = k.split(".", 2)
            for a, b in _BLOCK_KEYS:
                if rest.startswith(a):
                    rest = b + rest[len(a):]
                    break
            k = f"blocks.{idx}.{rest}"
        out[k] = v
    return out


def dit_forward(sd: dict, cfg: do.DenoiserConfig, hidden_states: torch.Tensor, timestep: torch.Tensor,
                encoder_hidden_states: torch.Tensor) -> torch.Tensor:
    """TripoSGDiTModel.forward (triposg_transformer.py:633-713): (B, N, C), (B,), (B, S, Dc) -> (B, N, C).
    `sd` uses ActionMesh key names (remap_state_dict); no image_rotary_emb on this path, i.e. the rotation is the
    identity (frame position 0), and every layer's self-attention is over the single frame."""
    B = hidden_states.shape[0]
    out, _ = do.denoiser_forward(sd, cfg, hidden_states[:, None], encoder_hidden_states[:, None],
                                 torch.zeros(B, 1, device=hidden_states.device), timestep, None, None)
    return out[:, 0]


def rectified_flow_sigmas(num_inference_steps: int, shift: float = 1.0, num_train_timesteps: int = 1000):
    """RectifiedFlowScheduler.set_timesteps (scheduling_rectified_flow.py:177-215): u_i = 1 - i/n, sigma = s u / (1 + (s-1) u),
    timesteps = sigma * num_train_timesteps, sigmas gets a trailing 0."""
    u = np.array([(1.0 - i / num_inference_steps) * num_train_timesteps for i in range(num_inference_steps)]) / num_train_timesteps
    sig = shift * u / (1 + (shift - 1) * u)
    sig = torch.from_numpy(sig).to(torch.float32)
    return sig * num_train_timesteps, torch.cat([sig, torch.zeros(1)])


@torch.no_grad()
def stage0_denoise(sd: dict, cfg: do.DenoiserConfig, image_embeds: torch.Tensor, latents: torch.Tensor, *,
                   num_inference_steps: int, guidance_scale: float, shift: float = 1.0) -> torch.Tensor:
    """Denoising loop of TripoSGPipeline.__call__ (pipeline_triposg.py:243-294) with classifier-free guidance: batch
    [zeros_like(embeds), embeds] (:141-145,214-215), v = v_uncond + s (v_img - v_uncond) (:266-270), Euler step
    x <- x + (sigma_i - sigma_{i+1}) v in fp32 (scheduling_rectified_flow.py:289-296)."""
    timesteps, sigmas = rectified_flow_sigmas(num_inference_steps, shift)
    emb = torch.cat([torch.zeros_like(image_embeds), image_embeds], dim=0)
    x = latents.clone().float()
    for i, t in enumerate(timesteps):
        pred = dit_forward(sd, cfg, torch.cat([x, x]), t.expand(2 * x.shape[0]), emb)
        unc, img = pred.chunk(2)
        v = unc + guidance_scale * (img - unc)
        x = x + (sigmas[i] - sigmas[i + 1]) * v
    return x
