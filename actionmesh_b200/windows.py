"""Autoregressive window partition and the latent bank — host-side bookkeeping of Stage I.

`chunk_from` follows actionmesh/model/utils/timesteps.py:10-117 (windows of `size` frames sliding by `slide`, expanding
from the anchor in both directions; the 1-frame overlap carries the already-denoised latent into the next window) and
`LatentBank` follows actionmesh/model/utils/storage.py:20-183 (per-frame latents keyed by float timestep, eps 1e-5).
Pure index logic, no arithmetic: it stays in Python like the reference.
"""
from __future__ import annotations

import torch


def chunk_right(start: int, end: int, size: int, slide: int) -> list[torch.Tensor]:
    """Left-to-right windows: the right edge moves to start+size, then by `slide`, clamped to `end`; every window is the
    `size` indices left of the edge, clamped to `start` (timesteps.py:10-48)."""
    assert 0 < slide <= size, f"Need slide <= size, got {slide} > {size}"
    chunks: list[torch.Tensor] = []
    edge = start
    while edge < end:
        edge = min(start + size, end) if not chunks else min(edge + slide, end)
        chunks.append(torch.arange(max(start, edge - size), edge))
    return chunks


def chunk_left(start: int, end: int, size: int, slide: int) -> list[torch.Tensor]:
    """chunk_right's windows, rightmost first, each with descending indices (timesteps.py:51-74)."""
    return [c.flip(0) for c in reversed(chunk_right(start, end, size, slide))]


def chunk_from(start: int, total: int, size: int, slide: int) -> list[torch.Tensor]:
    """AR windows starting at the anchor index `start` (timesteps.py:77-117)."""
    context = size - slide
    if total == size:
        idx = torch.arange(total)
        return [torch.cat([idx[start:start + 1], idx[idx != start]])]
    if start == 0:
        return chunk_right(0, total, size, slide)
    if start == total - 1:
        return chunk_left(0, total, size, slide)
    if start > total - start:  # more frames on the left: go left first
        left = chunk_left(0, start + 1, size, slide)
        right_start = min(max(0, start - context + 1), total - size)
        return left + chunk_right(right_start, total, size, slide)
    right = chunk_right(start, total, size, slide)
    left_end = max(min(start + context, total), size)
    return right + chunk_left(0, left_end, size, slide)


class LatentBank:
    """Timestep-indexed store of per-frame latents (storage.py:90-183).  Latents stay on the device they were produced
    on (fp32); `get` stacks the requested frames and returns an int32 presence mask, zeros where absent."""

    def __init__(self, empty_dims=(2048, 64), verbose: bool = False, tag: str = ""):
        self.empty_dims = tuple(empty_dims)
        self.items: list[torch.Tensor] = []
        self.timesteps: list[float] = []
        self.verbose, self.tag = verbose, tag

    @property
    def n_timesteps(self) -> int:
        return len(self.timesteps)

    def get_timestep_index(self, timestep: float, eps: float = 1e-5):
        for i, ts in enumerate(self.timesteps):
            if abs(ts - timestep) < eps:
                return i
        return None

    def update(self, timesteps: torch.Tensor, latents: torch.Tensor, replace: bool = False) -> None:
        ts = [float(t) for t in timesteps.flatten().tolist()]  # one host transfer instead of an .item() per frame
        lat = latents.reshape(len(ts), *self.empty_dims)
        for i, t in enumerate(ts):
            idx = self.get_timestep_index(t)
            if idx is None:
                self.timesteps.append(t)
                self.items.append(lat[i])
            elif replace:
                self.items[idx] = lat[i]

    def get(self, timesteps: torch.Tensor, device, add_batch_dim: bool = False):
        assert timesteps.ndim == 1
        lat, msk = [], []
        for t in timesteps.tolist():
            idx = self.get_timestep_index(t)
            if idx is None:
                lat.append(torch.zeros(self.empty_dims, dtype=torch.float32, device=device))
                msk.append(0)
            else:
                lat.append(self.items[idx].to(device=device, dtype=torch.float32))
                msk.append(1)
        lat_out = torch.stack(lat)
        msk_out = torch.tensor(msk, dtype=torch.int32, device=device)
        return (lat_out[None], msk_out[None]) if add_batch_dim else (lat_out, msk_out)

    def get_ordered(self):
        order = sorted(range(len(self.timesteps)), key=lambda i: self.timesteps[i])
        lat = torch.stack([self.items[i] for i in order])
        return lat, torch.tensor([self.timesteps[i] for i in order]).to(lat)

    def get_ordered_timesteps(self) -> torch.Tensor:
        """storage.py TimestepIndexedStorage.get_ordered_timesteps: all stored timesteps, ascending (fp32, CPU)."""
        return torch.tensor(sorted(self.timesteps), dtype=torch.float32)


class VertexBank:
    """Timestep-indexed store of per-frame vertex arrays — the role `MeshBank` (storage.py:187-260) plays in Stage II.
    The reference stores trimesh objects that all share the anchor's faces; only the vertices change, so this bank keeps
    (V, 3) fp32 tensors and one shared `faces` array the caller may attach."""

    def __init__(self, faces=None):
        self.items: list[torch.Tensor] = []
        self.timesteps: list[float] = []
        self.faces = faces

    @property
    def n_timesteps(self) -> int:
        return len(self.timesteps)

    def get_timestep_index(self, timestep: float, eps: float = 1e-5):
        for i, ts in enumerate(self.timesteps):
            if abs(ts - timestep) < eps:
                return i
        return None

    def update(self, timesteps: torch.Tensor, vertices, replace: bool = False) -> None:
        ts = [float(t) for t in timesteps.flatten().tolist()]
        assert len(ts) == len(vertices)
        for i, t in enumerate(ts):
            idx = self.get_timestep_index(t)
            if idx is None:
                self.timesteps.append(t)
                self.items.append(vertices[i])
            elif replace:
                self.items[idx] = vertices[i]

    def get(self, timesteps: torch.Tensor) -> list:
        assert timesteps.ndim == 1
        out = []
        for t in timesteps.tolist():
            idx = self.get_timestep_index(t)
            out.append(self.items[idx] if idx is not None else None)
        return out

    def get_ordered(self):
        order = sorted(range(len(self.timesteps)), key=lambda i: self.timesteps[i])
        return [self.items[i] for i in order], torch.tensor([self.timesteps[i] for i in order], dtype=torch.float32)


# ---- Stage-II time bookkeeping (actionmesh/model/utils/embeddings.py:156-242) ------------------------------------------
def get_scaling(timesteps: torch.Tensor):
    """embeddings.py:156-173: per-batch (min, max - min) of (B, T) timesteps."""
    t_min = timesteps.min(dim=1).values
    return t_min, timesteps.max(dim=1).values - t_min


def apply_scaling(timesteps: torch.Tensor, t_min: torch.Tensor, t_range: torch.Tensor) -> torch.Tensor:
    """embeddings.py:176-196: (t - t_min) / t_range, for (B,) or (B, T) inputs."""
    if timesteps.dim() == 1:
        return (timesteps - t_min) / t_range
    return (timesteps - t_min.unsqueeze(1)) / t_range.unsqueeze(1)


def get_n_subdivisions(start, end, level: int = 1) -> int:
    """embeddings.py:199-214: number of points after `level - 1` rounds of midpoint insertion."""
    n_points = int(end - start + 1)
    for _ in range(1, level):
        n_points += n_points - 1
    return n_points


def interpolate_timesteps(timesteps: torch.Tensor, subsampling_level: int, device="cpu", drop_first: bool = False) -> torch.Tensor:
    """embeddings.py:217-242: linspace(min, max, n_subdivisions) as (1, n) [(1, n-1) with drop_first]."""
    t_min, t_max = timesteps.min().item(), timesteps.max().item()
    out = torch.linspace(t_min, t_max, get_n_subdivisions(t_min, t_max, level=subsampling_level), device=device).reshape(1, -1)
    return out[:, 1:] if drop_first else out
