"""Stage1Pipeline / AnimationPipeline — the hot path of `ActionMeshPipeline.__call__` (reference actionmesh/pipeline.py:602-685) without the
out-of-scope stages: DinoV2 context for all frames (`encode_all_frames`, :232-245), then the autoregressive Stage-I
denoising over 16-frame windows (`generate_3d_latents` :435-508 -> `_denoise_latents` :247-314).

Stage 0 (TripoSG anchor latent + mesh), background removal, IO and rendering stay on the reference (SURVEY 2.1 marks
them out of scope); the anchor latent therefore comes in through a seeded `LatentBank`, exactly the object
`init_banks_from_anchor` hands to `generate_3d_latents` in the reference (:661,:672).  `AnimationPipeline` adds Stage II
(`generate_mesh_animation` :510-600 -> `_decode_displacement` :316-385) on the CUDA autoencoder: the anchor mesh comes in
as vertex features (positions + unit normals, mesh_processor.py:85-101) and the result is a `VertexBank` (all output
meshes share the anchor's faces, so only vertices are produced).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional

import torch

from .denoiser import B200Denoiser
from .guidance import ClassifierFreeGuidance
from .image_encoder import B200ImageEncoder
from .scheduler import B200SchedulerFlow
from .windows import (LatentBank, VertexBank, apply_scaling, chunk_from, get_scaling, interpolate_timesteps)


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



class AnimationPipeline(Stage1Pipeline):
    """Stage I + Stage II of `ActionMeshPipeline.__call__` (pipeline.py:637-683) on the CUDA path."""

    def __init__(self, denoiser, scheduler, cf_guidance, autoencoder, image_encoder=None, *,
                 sliding_window_autoencoder: int = 15, subsampling_level: int = 1,
                 normals_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, **kwargs):
        super().__init__(denoiser, scheduler, cf_guidance, image_encoder, **kwargs)
        self.temporal_3D_vae = autoencoder
        self.sliding_window_autoencoder = sliding_window_autoencoder
        self.subsampling_level = subsampling_level
        # (V, 3) vertices -> (V, 3) unit vertex normals of the deformed anchor mesh; the reference gets them from
        # trimesh (`mesh.vertex_normals`, mesh_processor.py:98).  Needed only when a clip spans more than one AR window.
        self.normals_fn = normals_fn

    def _decode_displacement(self, latents: torch.Tensor, window_timesteps: torch.Tensor, source_alpha: torch.Tensor,
                             target_alphas: torch.Tensor, vertex_features: torch.Tensor,
                             step_callback: Optional[Callable[[int, int], None]] = None) -> torch.Tensor:
        """One AR window (pipeline.py:316-385): (V, 3|6) anchor features -> (T_out, V, 3) deformed vertices."""
        q = vertex_features[None].to(self.device)
        disp = self.temporal_3D_vae(latent=latents, framestep=window_timesteps, source_alpha=source_alpha,
                                    target_alphas=target_alphas, query=q, step_callback=step_callback)
        return self.temporal_3D_vae.apply_displacement(vertex=q[..., :3], displacement=disp)[0]

    def generate_mesh_animation(self, latent_bank: LatentBank, vertex_bank: VertexBank, anchor_normals: torch.Tensor,
                                step_callback: Optional[Callable[[int, int, int, int], None]] = None) -> VertexBank:
        """Serial AR windows over the denoised latents (pipeline.py:510-600).  `vertex_bank` holds the anchor vertices at
        the anchor timestep; `anchor_normals` are its unit vertex normals."""
        windows = chunk_from(start=self.anchor_idx, total=latent_bank.n_timesteps,
                             size=self.temporal_3D_vae.config.temporal_context_size, slide=self.sliding_window_autoencoder)
        all_ts = latent_bank.get_ordered_timesteps()
        anchor_t = sorted(vertex_bank.timesteps)[0] if vertex_bank.n_timesteps == 1 else None
        for wi, idx in enumerate(windows):
            wts = all_ts[idx][None]                                             # (1, T)
            lat, _ = latent_bank.get(timesteps=wts[0], device=self.device, add_batch_dim=True)
            verts = vertex_bank.get(timesteps=wts[:, 0])[0]
            assert verts is not None, "Anchor mesh should be in the vertex bank"
            if anchor_t is not None and abs(float(wts[0, 0]) - anchor_t) < 1e-5:
                normals = anchor_normals
            elif self.normals_fn is not None:
                normals = self.normals_fn(verts)
            else:
                raise ValueError("a clip spanning several AR windows needs `normals_fn` (vertex normals of the deformed anchor)")
            feats = torch.cat([verts.to(self.device, torch.float32), normals.to(self.device, torch.float32)], dim=-1)
            out_ts = interpolate_timesteps(wts, subsampling_level=self.subsampling_level, device="cpu", drop_first=True)
            t_min, t_range = get_scaling(wts)
            cb = None
            if step_callback is not None:
                cb = (lambda step, total, _i=wi, _n=len(windows): step_callback(step, total, _i, _n))
            v = self._decode_displacement(lat, wts, apply_scaling(wts[:, 0], t_min, t_range),
                                          apply_scaling(out_ts, t_min, t_range), feats, step_callback=cb)
            vertex_bank.update(timesteps=out_ts[0], vertices=list(v))
        return vertex_bank

    def __call__(self, input: VideoInput, anchor_latent: torch.Tensor, anchor_vertices: torch.Tensor,
                 anchor_normals: torch.Tensor, seed: int = 44, stage_1_steps: Optional[int] = None,
                 guidance_scales: Optional[List[float]] = None, anchor_idx: Optional[int] = None,
                 context: Optional[torch.Tensor] = None, faces=None):
        """-> (LatentBank, VertexBank): Stage I then Stage II for every output timestep."""
        bank = super().__call__(input, anchor_latent, seed=seed, stage_1_steps=stage_1_steps,
                                guidance_scales=guidance_scales, anchor_idx=anchor_idx, context=context)
        vb = VertexBank(faces=faces)
        vb.update(timesteps=input.timesteps[self.anchor_idx:self.anchor_idx + 1],
                  vertices=[anchor_vertices.to(self.device, torch.float32)])
        return bank, self.generate_mesh_animation(bank, vb, anchor_normals)
