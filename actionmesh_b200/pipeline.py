"""ActionMeshB200Pipeline / Stage1Pipeline / AnimationPipeline — the hot path of `ActionMeshPipeline.__call__` (reference actionmesh/pipeline.py:602-685) without the
out-of-scope stages: DinoV2 context for all frames (`encode_all_frames`, :232-245), then the autoregressive Stage-I
denoising over 16-frame windows (`generate_3d_latents` :435-508 -> `_denoise_latents` :247-314).

Stage 0 (TripoSG anchor latent + mesh), background removal, IO and rendering stay on the reference (SURVEY 2.1 marks
them out of scope); the anchor latent therefore comes in through a seeded `LatentBank`, exactly the object
`init_banks_from_anchor` hands to `generate_3d_latents` in the reference (:661,:672).  `AnimationPipeline` adds Stage II
(`generate_mesh_animation` :510-600 -> `_decode_displacement` :316-385) on the CUDA autoencoder: the anchor mesh comes in
as vertex features (positions + unit normals, mesh_processor.py:85-101) and the result is a `VertexBank` (all output
meshes share the anchor's faces, so only vertices are produced).

`ActionMeshB200Pipeline` is seam 4: the reference's constructor and `__call__(input, seed, stage_0_steps,
face_decimation, floaters_threshold, stage_1_steps, guidance_scales, anchor_idx) -> list of meshes` (pipeline.py:47-53,
602-613) built from `actionmesh_b200*.yaml` through the same `_target_` plumbing, with the out-of-scope stages (TripoSG,
background removal, CPU cropping, mesh post-processing) as injected components.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, List, Optional

import torch

from ._lib import AmbError
from .config import DEFAULT_CONFIG_DIR, get_target, instantiate, load_config
from .denoiser import B200Denoiser
from .guidance import ClassifierFreeGuidance
from .image_encoder import B200ImageEncoder
from .scheduler import B200SchedulerFlow
from .windows import (LatentBank, VertexBank, apply_scaling, chunk_from, get_scaling, interpolate_timesteps)


MIN_FRAMES = 16  # actionmesh/io/video_input.py:24


@dataclass
class ActionMeshInput:
    """Same fields and checks as the reference's `ActionMeshInput` (actionmesh/io/video_input.py:27-55): N >= 16 RGB(A) PIL
    frames and their float32 CPU timesteps.  The pipeline accepts the reference's own dataclass as well (duck-typed)."""
    frames: list
    timesteps: torch.Tensor

    def __post_init__(self) -> None:
        assert len(self.frames) >= MIN_FRAMES, f"At least {MIN_FRAMES} frames are required, got {len(self.frames)}"
        assert self.timesteps.ndim == 1, f"Expected 1D timesteps, got {self.timesteps.ndim}D"
        assert len(self.frames) == self.timesteps.shape[0], \
            f"Number of frames ({len(self.frames)}) must match timesteps ({self.timesteps.shape[0]})"
        assert self.timesteps.dtype == torch.float32, f"Expected float32 timesteps, got {self.timesteps.dtype}"
        assert self.timesteps.device.type == "cpu", f"Expected CPU timesteps, got {self.timesteps.device}"

    @property
    def n_frames(self) -> int:
        return len(self.frames)

    def get(self, indices: torch.Tensor) -> "VideoInput":
        return VideoInput([self.frames[int(i)] for i in indices], self.timesteps[indices])


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


# ---------------------------------------------------------------------------------------------------- seam 4: the pipeline
class PassThroughMeshProcess:
    """Stand-in for the reference's CPU `MeshPostprocessor` (actionmesh/preprocessing/mesh_processor.py:374, decimation +
    floater removal — out of scope): same constructor arguments / mutable attributes, `process_mesh` returns its input."""

    def __init__(self, face_decimation: int = 40000, floaters_threshold: float = 0.02):
        self.face_decimation = face_decimation
        self.floaters_threshold = floaters_threshold

    def process_mesh(self, mesh, seed: int = 44):
        return mesh


@dataclass
class Mesh:
    """Minimal mesh record returned when `trimesh` is not installed: (V, 3) float32 vertices + (F, 3) int faces (numpy)."""
    vertices: "object"
    faces: "object"


def _vertex_normals(vertices: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """Unit vertex normals of a triangle mesh (area-weighted face normals).  Used for the deformed anchor of clips spanning
    several AR windows when trimesh is unavailable (the reference reads trimesh's `vertex_normals`, mesh_processor.py:98)."""
    v0, v1, v2 = (vertices[faces[:, i]] for i in range(3))
    fn = torch.cross(v1 - v0, v2 - v0, dim=-1)
    vn = torch.zeros_like(vertices)
    for i in range(3):
        vn.index_add_(0, faces[:, i], fn)
    return vn / vn.norm(dim=-1, keepdim=True).clamp_min(1e-12)


class ActionMeshB200Pipeline:
    """Drop-in for `ActionMeshPipeline` on the B200 path: same constructor arguments, `.to(device)`, and `__call__`
    signature / return value (reference actionmesh/pipeline.py:47-53,205,602-685).

    Built from `actionmesh_b200.yaml` / `actionmesh_b200_fast.yaml` (the reference's YAML with the `_target_`s re-pointed):
    scheduler, guidance, mesh post-process are instantiated at construction like the reference (:98-110); denoiser, image
    encoder and autoencoder are resolved from their `_target_`s and loaded from `weights_dir` by `.to()` (or assigned
    directly — `pipe.temporal_3D_denoiser = model` — when weights do not come from disk).  Injected, because out of scope:
      image_to_3d(image=, generator=, num_inference_steps=, guidance_scale=) -> (anchor_latent, anchor_mesh)   [TripoSG]
      background_removal.process_images(frames), image_process.process_images(frames)                          [CPU, optional]
    `anchor_mesh` needs `.vertices`, `.faces` and `.vertex_normals` (the trimesh attributes the reference reads)."""

    def __init__(self, config_name: str = "actionmesh_b200.yaml", config_dir: Optional[str] = None,
                 dtype: torch.dtype = torch.bfloat16, lazy_loading: bool = False, *, image_to_3d=None,
                 background_removal=None, image_process=None, weights_dir: str = "pretrained_weights/ActionMesh",
                 config_updates: Optional[dict] = None):
        self.cfg = load_config(config_name, config_dir or DEFAULT_CONFIG_DIR, updates=config_updates)
        self._actionmesh_weights_dir = weights_dir
        self.image_to_3d_pipe = image_to_3d
        self.background_removal = background_removal
        self.image_process = image_process
        self.temporal_3D_denoiser = None
        self.image_encoder = None
        self.temporal_3D_vae = None
        self._denoiser_latent_shape = tuple(self.cfg.denoiser_latent_shape)
        self.mesh_process = instantiate(self.cfg.model.mesh_process, _convert_="partial")()
        self.scheduler = instantiate(self.cfg.model.scheduler, _convert_="partial")()
        self.cf_guidance = instantiate(self.cfg.model.cf_guidance, _convert_="partial")()
        self._target_device = torch.device("cpu")
        self._dtype = dtype            # accepted for signature compatibility: the CUDA path owns its precision recipe
        self._lazy_loading = lazy_loading

    # ---- model lifecycle (pipeline.py:117-229)
    def _load_image_encoder(self) -> None:
        if self.image_encoder is None:
            self.image_encoder = instantiate(self.cfg.model.image_encoder, _convert_="partial")()
        self.image_encoder.to(self._target_device)

    def _load_temporal_denoiser(self) -> None:
        if self.temporal_3D_denoiser is None:
            cls = get_target(self.cfg.model.temporal_3D_denoiser["_target_"])
            self.temporal_3D_denoiser = cls.from_pretrained(os.path.join(self._actionmesh_weights_dir, "denoiser"),
                                                            device=self._target_device)
        self.temporal_3D_denoiser.to(self._target_device)

    def _load_temporal_vae(self) -> None:
        if self.temporal_3D_vae is None:
            cls = get_target(self.cfg.model.temporal_3D_vae["_target_"])
            self.temporal_3D_vae = cls.from_pretrained(os.path.join(self._actionmesh_weights_dir, "autoencoder"),
                                                       device=self._target_device)
        self.temporal_3D_vae.to(self._target_device)

    def _unload_model(self, attr: str) -> None:
        if self._lazy_loading and getattr(self, attr, None) is not None:
            setattr(self, attr, None)
            torch.cuda.empty_cache()

    def to(self, device) -> "ActionMeshB200Pipeline":
        device = torch.device(device)
        if device.type != "cuda":
            raise AmbError("ActionMeshB200Pipeline runs on CUDA (sm_100a) only; there is no CPU fallback")
        self._target_device = device
        if not self._lazy_loading:
            self._load_image_encoder()
            self._load_temporal_denoiser()
            self._load_temporal_vae()
        return self

    @property
    def device(self) -> torch.device:
        return self._target_device

    # ---- stages
    def init_banks_from_anchor(self, input, seed: int = 44):
        """Stage 0 through the injected image-to-3D component (pipeline.py:387-433) -> (LatentBank, anchor mesh)."""
        if self.image_to_3d_pipe is None:
            raise AmbError("Stage 0 (TripoSG image-to-3D) is not part of actionmesh_b200: pass image_to_3d=<callable> "
                           "returning (anchor_latent, anchor_mesh)")
        gen_dev = getattr(self.image_to_3d_pipe, "device", self._target_device)
        anchor_latent, anchor_mesh = self.image_to_3d_pipe(
            image=input.frames[self.cfg.anchor_idx], generator=torch.Generator(device=gen_dev).manual_seed(seed),
            num_inference_steps=self.cfg.model.image_to_3D_denoiser.num_inference_steps,
            guidance_scale=self.cfg.model.image_to_3D_denoiser.guidance_scale)
        anchor_mesh = self.mesh_process.process_mesh(anchor_mesh, seed=seed)
        bank = LatentBank(empty_dims=self._denoiser_latent_shape)
        bank.update(timesteps=input.timesteps[[self.cfg.anchor_idx]],
                    latents=torch.as_tensor(anchor_latent).to(device=self._target_device, dtype=torch.float32))
        return bank, anchor_mesh

    def _stage_pipeline(self) -> "AnimationPipeline":
        faces_holder = {}

        def normals_fn(verts: torch.Tensor) -> torch.Tensor:
            faces = faces_holder["faces"]
            try:
                import trimesh  # the reference's source of vertex normals

                return torch.as_tensor(trimesh.Trimesh(vertices=verts.cpu().numpy(), faces=faces.cpu().numpy(),
                                                       process=False).vertex_normals.copy(), dtype=torch.float32)
            except ImportError:
                return _vertex_normals(verts.to(torch.float32), faces.to(verts.device))

        pipe = AnimationPipeline(self.temporal_3D_denoiser, self.scheduler, self.cf_guidance, self.temporal_3D_vae,
                                 self.image_encoder, sliding_window_autoencoder=self.cfg.sliding_window_autoencoder,
                                 subsampling_level=self.cfg.subsampling_level, normals_fn=normals_fn,
                                 temporal_context_size=self.cfg.model.temporal_3D_denoiser.temporal_context_size,
                                 sliding_window_denoiser=self.cfg.sliding_window_denoiser, anchor_idx=self.cfg.anchor_idx,
                                 latent_shape=self._denoiser_latent_shape)
        pipe._faces_holder = faces_holder
        return pipe

    @torch.no_grad()
    def __call__(self, input, seed: int = 44, stage_0_steps: Optional[int] = None, face_decimation: Optional[int] = None,
                 floaters_threshold: Optional[float] = None, stage_1_steps: Optional[int] = None,
                 guidance_scales: Optional[List[float]] = None, anchor_idx: Optional[int] = None) -> list:
        """video -> 4D (pipeline.py:602-685): returns the animated meshes (fixed topology) ordered by timestep."""
        if stage_0_steps is not None:
            self.cfg.model.image_to_3D_denoiser.num_inference_steps = stage_0_steps
        if stage_1_steps is not None:
            self.scheduler.num_inference_steps = stage_1_steps
        if guidance_scales is not None:
            self.cf_guidance.guidance_scales = guidance_scales
        if face_decimation is not None:
            self.mesh_process.face_decimation = face_decimation
        if floaters_threshold is not None:
            self.mesh_process.floaters_threshold = floaters_threshold
        if anchor_idx is not None:
            self.cfg.anchor_idx = anchor_idx
        if self.background_removal is not None:
            input.frames = self.background_removal.process_images(input.frames)
        if self.image_process is not None:
            input.frames = self.image_process.process_images(input.frames)

        latent_bank, anchor_mesh = self.init_banks_from_anchor(input, seed)          # Stage 0
        self._load_image_encoder()
        vin = VideoInput(list(input.frames), input.timesteps)
        self._load_temporal_denoiser()
        self._load_temporal_vae()
        stages = self._stage_pipeline()
        context = stages.encode_all_frames(vin)                                       # DinoV2 on all frames
        self._unload_model("image_encoder")
        latent_bank = stages.generate_3d_latents(vin, context, latent_bank, seed=seed)   # Stage I
        self._unload_model("temporal_3D_denoiser")
        dev = self._target_device
        verts = torch.as_tensor(anchor_mesh.vertices, dtype=torch.float32).to(dev)
        faces = torch.as_tensor(anchor_mesh.faces).to(torch.int64)
        normals = torch.as_tensor(anchor_mesh.vertex_normals, dtype=torch.float32).to(dev)
        stages._faces_holder["faces"] = faces
        vb = VertexBank(faces=faces)
        vb.update(timesteps=input.timesteps[[self.cfg.anchor_idx]], vertices=[verts])
        vb = stages.generate_mesh_animation(latent_bank, vb, normals)               # Stage II
        self._unload_model("temporal_3D_vae")
        ordered, _ = vb.get_ordered()
        out = []
        f_np = faces.cpu().numpy()
        try:
            import trimesh

            for v in ordered:
                out.append(trimesh.Trimesh(vertices=v.cpu().numpy(), faces=f_np, process=False))
        except ImportError:
            for v in ordered:
                out.append(Mesh(vertices=v.cpu().numpy(), faces=f_np))
        return out
