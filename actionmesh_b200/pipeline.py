"""Stage1Pipeline — the hot path of `ActionMeshPipeline.__call__` (reference actionmesh/pipeline.py:602-685) without the
out-of-scope stages: DinoV2 context for all frames (`encode_all_frames`, :232-245), then the autoregressive Stage-I
denoising over 16-frame windows (`generate_3d_latents` :435-508 -> `_denoise_latents` :247-314).

Stage 0 (TripoSG anchor latent), Stage II (mesh decoding), background removal, IO and rendering stay on the reference
(SURVEY 2.1 marks them out of scope); the anchor latent therefore comes in through a seeded `LatentBank`, exactly the
object `init_banks_from_anchor` hands to `generate_3d_latents` in the reference (:661,:672).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional

import torch

from .denoiser import B200Denoiser
from .guidance import ClassifierFreeGuidance
from .image_encoder import B200ImageEncoder
from .scheduler import B200SchedulerFlow
from .windows import LatentBank, chunk_from


@dataclass
class VideoInput:
    """The part of `ActionMeshInput` (actionmesh/io/video_input.py:27-55) Stage I consumes: frames + float timesteps."""
    frames: list
    timesteps: torch.Tensor  # (T,) fp32, CPU like the reference (video_input.py:53-55)

    @property
    def n_frames(self) -> int:
        return len(self.frames)

    def get(self, indices: torch.Tensor) -> "VideoInput":
        return VideoInput([self.frames[int(i)] for i in indices], self.timesteps[indices])


class Stage1Pipeline:
    def __init__(self, denoiser: B200Denoiser, scheduler: B200SchedulerFlow, cf_guidance: ClassifierFreeGuidance,
                 image_encoder: Optional[B200ImageEncoder] = None, temporal_context_size: int = 16,
                 sliding_window_denoiser: int = 15, anchor_idx: int = 0, latent_shape=(2048, 64)):
        self.temporal_3D_denoiser = denoiser
        self.scheduler = scheduler
        self.cf_guidance = cf_guidance
        self.image_encoder = image_encoder
        self.temporal_context_size = temporal_context_size
        self.sliding_window_denoiser = sliding_window_denoiser
        self.anchor_idx = anchor_idx
        self._denoiser_latent_shape = list(latent_shape)

    @property
    def device(self) -> torch.device:
        return self.temporal_3D_denoiser.device

    def encode_all_frames(self, input: VideoInput) -> torch.Tensor:
        """(T, S, D) context for all frames (pipeline.py:232-245)."""
        return self.image_encoder.encode_images(input.frames)

    def _denoise_latents(self, input: VideoInput, context: torch.Tensor, latent_bank: LatentBank, seed: int = 44,
                         step_callback: Optional[Callable[[int, int], None]] = None) -> torch.Tensor:
        """One AR window (pipeline.py:247-314)."""
        generator = torch.Generator(device=self.device).manual_seed(seed)
        cond_latents, cond_mask = latent_bank.get(timesteps=input.timesteps, device=self.device, add_batch_dim=True)
        init_noise = self.scheduler.get_noise(batch_size=1, latent_shape=self._denoiser_latent_shape,
                                              n_timesteps=input.n_frames, generator=generator, device=self.device)
        m = cond_mask[..., None, None].to(torch.float32)
        init_latent = cond_latents * m + init_noise * (1.0 - m)  # window set-up, once per window (pipeline.py:297)
        return self.scheduler.denoise(self.temporal_3D_denoiser, self.cf_guidance, init_latent=init_latent,
                                      context=context[None], mask=cond_mask.to(init_latent.dtype),
                                      framestep=input.timesteps[None], device=self.device, disable_prog=True,
                                      step_callback=step_callback)

    def generate_3d_latents(self, input: VideoInput, context: torch.Tensor, latent_bank: LatentBank, seed: int = 44,
                            step_callback: Optional[Callable[[int, int, int, int], None]] = None) -> LatentBank:
        """Serial AR windows, seed + i per window (pipeline.py:435-508)."""
        windows = chunk_from(start=self.anchor_idx, total=input.n_frames, size=self.temporal_context_size,
                             slide=self.sliding_window_denoiser)
        for i, idx in enumerate(windows):
            cb = None
            if step_callback is not None:
                cb = (lambda step, total, _i=i, _n=len(windows): step_callback(step, total, _i, _n))
            win = input.get(idx)
            lat = self._denoise_latents(win, context[idx.to(context.device)], latent_bank, seed=seed + i, step_callback=cb)
            latent_bank.update(latents=lat, timesteps=win.timesteps)
        return latent_bank

    @torch.no_grad()
    def __call__(self, input: VideoInput, anchor_latent: torch.Tensor, seed: int = 44,
                 stage_1_steps: Optional[int] = None, guidance_scales: Optional[List[float]] = None,
                 anchor_idx: Optional[int] = None, context: Optional[torch.Tensor] = None) -> LatentBank:
        """Stage-I part of ActionMeshPipeline.__call__ (pipeline.py:637-675): same override plumbing, returns the bank of
        denoised latents (what Stage II consumes)."""
        if stage_1_steps is not None:
            self.scheduler.num_inference_steps = stage_1_steps
        if guidance_scales is not None:
            self.cf_guidance.guidance_scales = guidance_scales
        if anchor_idx is not None:
            self.anchor_idx = anchor_idx
        bank = LatentBank(empty_dims=tuple(self._denoiser_latent_shape))
        bank.update(timesteps=input.timesteps[self.anchor_idx:self.anchor_idx + 1],
                    latents=anchor_latent.to(device=self.device, dtype=torch.float32))
        if context is None:
            context = self.encode_all_frames(input)
        return self.generate_3d_latents(input, context, bank, seed=seed)
