"""TEST INFRASTRUCTURE ONLY.  Minimal in-memory stand-in for the `diffusers` classes the reference hot path imports.

`diffusers` is an unpinned third-party dependency of the reference (requirements.txt:11), not vendored under
/root/reference and not installed in this image.  What follows restates the published semantics of diffusers >= 0.30
(the version TripoSG's scheduler header cites is v0.30.3) for exactly the constructor arguments the reference uses:
    block.py:12-14,64-104        Attention, FeedForward, FP32LayerNorm
    temporal_denoiser.py:16,57-68 Timesteps, TimestepEmbedding
"diffusers semantics unpinned": there is no wheel here to check this file against.
"""
from __future__ import annotations

import inspect
import math
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F


class FP32LayerNorm(nn.LayerNorm):
    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        origin_dtype = inputs.dtype
        return F.layer_norm(
            inputs.float(), self.normalized_shape,
            self.weight.float() if self.weight is not None else None,
            self.bias.float() if self.bias is not None else None, self.eps,
        ).to(origin_dtype)


class RMSNorm(nn.Module):
    def __init__(self, dim, eps: float, elementwise_affine: bool = True, bias: bool = False):
        super().__init__()
        self.eps = eps
        self.dim = (dim,) if isinstance(dim, int) else tuple(dim)
        self.weight = nn.Parameter(torch.ones(self.dim)) if elementwise_affine else None
        self.bias = None

    def forward(self, hidden_states):
        input_dtype = hidden_states.dtype
        variance = hidden_states.to(torch.float32).pow(2).mean(-1, keepdim=True)
        hidden_states = hidden_states * torch.rsqrt(variance + self.eps)
        if self.weight is not None:
            if self.weight.dtype in (torch.float16, torch.bfloat16):
                hidden_states = hidden_states.to(self.weight.dtype)
            hidden_states = hidden_states * self.weight
        else:
            hidden_states = hidden_states.to(input_dtype)
        return hidden_states


class Attention(nn.Module):
    """Parameter container + kwargs-forwarding `forward`, as diffusers.models.attention_processor.Attention."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 qk_norm=None, cross_attention_norm=None, eps=1e-5, out_bias=True, residual_connection=False,
                 rescale_output_factor=1.0, processor=None, **unused):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.query_dim = query_dim
        self.is_cross_attention = cross_attention_dim is not None
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.rescale_output_factor = rescale_output_factor
        self.residual_connection = residual_connection
        self.scale = dim_head ** -0.5
        self.group_norm = None
        self.spatial_norm = None
        if qk_norm is None:
            self.norm_q = None
            self.norm_k = None
        elif qk_norm == "rms_norm":
            self.norm_q = RMSNorm(dim_head, eps=eps)
            self.norm_k = RMSNorm(dim_head, eps=eps)
        else:
            raise ValueError(f"shim: qk_norm {qk_norm!r} not used by the reference")
        if cross_attention_norm is None:
            self.norm_cross = None
        elif cross_attention_norm == "layer_norm":
            self.norm_cross = nn.LayerNorm(self.cross_attention_dim)
        else:
            raise ValueError(f"shim: cross_attention_norm {cross_attention_norm!r} not supported")
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.processor = processor

    def norm_encoder_hidden_states(self, encoder_hidden_states):
        return self.norm_cross(encoder_hidden_states)

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        params = set(inspect.signature(self.processor.__call__).parameters.keys())
        kw = {k: v for k, v in cross_attention_kwargs.items() if k in params}
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)


class GELU(nn.Module):
    def __init__(self, dim_in, dim_out, approximate="none", bias=True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate

    def forward(self, hidden_states):
        return F.gelu(self.proj(hidden_states), approximate=self.approximate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False,
                 inner_dim=None, bias=True):
        super().__init__()
        if inner_dim is None:
            inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        if activation_fn != "gelu":
            raise ValueError("shim: only activation_fn='gelu' is used by the reference")
        self.net = nn.ModuleList([GELU(dim, inner_dim, bias=bias), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out, bias=bias)])

    def forward(self, hidden_states, *args, **kwargs):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift, scale: int = 1):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift
        self.scale = scale

    def forward(self, timesteps):
        half_dim = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(0, half_dim, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half_dim - self.downscale_freq_shift)
        emb = torch.exp(exponent)
        emb = timesteps[:, None].float() * emb[None, :]
        emb = self.scale * emb
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if self.flip_sin_to_cos:
            emb = torch.cat([emb[:, half_dim:], emb[:, :half_dim]], dim=-1)
        return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim, act_fn="silu", out_dim=None, **unused):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim, True)
        if act_fn != "gelu":
            raise ValueError("shim: only act_fn='gelu' is used by the reference")
        self.act = nn.GELU()
        self.linear_2 = nn.Linear(time_embed_dim, out_dim if out_dim is not None else time_embed_dim, True)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


def install() -> None:
    """Register the stand-in under the module names the reference imports."""
    if "diffusers" in sys.modules and not getattr(sys.modules["diffusers"], "_AMB_SHIM", False):
        return  # a real diffusers is present: use it
    def mod(name):
        m = types.ModuleType(name)
        m._AMB_SHIM = True
        sys.modules[name] = m
        return m
    d = mod("diffusers")
    dm = mod("diffusers.models")
    att = mod("diffusers.models.attention")
    ap = mod("diffusers.models.attention_processor")
    nm = mod("diffusers.models.normalization")
    em = mod("diffusers.models.embeddings")
    d.models = dm
    dm.attention, dm.attention_processor, dm.normalization, dm.embeddings = att, ap, nm, em
    att.FeedForward = FeedForward
    ap.Attention = Attention
    nm.FP32LayerNorm, nm.RMSNorm = FP32LayerNorm, RMSNorm
    em.Timesteps, em.TimestepEmbedding = Timesteps, TimestepEmbedding
    _install_triposg_extras(mod, d, dm, ap, nm, em)


# ---- what third_party/TripoSG's DiT and scheduler import on top (triposg_transformer.py:60-90, scheduling_rectified_flow.py:11-13)
class _Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as exc:
            raise AttributeError(k) from exc


def register_to_config(init):
    """diffusers.configuration_utils.register_to_config: record the constructor arguments (with defaults) in `self.config`."""
    import functools

    sig = inspect.signature(init)

    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        object.__setattr__(self, "_internal_config", _Config(cfg))
        init(self, *args, **kwargs)

    return wrapper


class ConfigMixin:
    @property
    def config(self):
        return self._internal_config


class ModelMixin(nn.Module):
    pass


class SchedulerMixin:
    pass


class PeftAdapterMixin:
    pass


class LayerNorm(nn.LayerNorm):
    """diffusers.models.normalization.LayerNorm(dim, eps=1e-5, elementwise_affine=True, bias=True) on torch >= 2.1."""

    def __init__(self, dim, eps: float = 1e-5, elementwise_affine: bool = True, bias: bool = True):
        super().__init__(dim, eps=eps, elementwise_affine=elementwise_affine, bias=bias)


def _install_triposg_extras(mod, d, dm, ap, nm, em):
    import dataclasses
    import logging as pylog

    cu = mod("diffusers.configuration_utils")
    cu.ConfigMixin, cu.register_to_config = ConfigMixin, register_to_config
    ld = mod("diffusers.loaders")
    ld.PeftAdapterMixin = PeftAdapterMixin
    mu = mod("diffusers.models.modeling_utils")
    mu.ModelMixin = ModelMixin
    ap.AttentionProcessor = object
    nm.LayerNorm = LayerNorm
    nm.AdaLayerNormContinuous = type("AdaLayerNormContinuous", (nn.Module,), {})
    em.GaussianFourierProjection = type("GaussianFourierProjection", (nn.Module,), {})
    em.apply_rotary_emb = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("apply_rotary_emb is not on the path"))
    ut = mod("diffusers.utils")
    ut.USE_PEFT_BACKEND = False
    ut.is_torch_version = lambda op, v: True
    ut.scale_lora_layers = lambda *a, **k: None
    ut.unscale_lora_layers = lambda *a, **k: None
    ut.logging = types.SimpleNamespace(get_logger=pylog.getLogger)
    ut.BaseOutput = type("BaseOutput", (), {})
    tu = mod("diffusers.utils.torch_utils")
    tu.maybe_allow_in_graph = lambda cls: cls
    ut.torch_utils = tu
    sc = mod("diffusers.schedulers")
    su = mod("diffusers.schedulers.scheduling_utils")
    su.SchedulerMixin = SchedulerMixin
    sc.scheduling_utils = su
    d.configuration_utils, d.loaders, d.utils, d.schedulers = cu, ld, ut, sc
    dm.modeling_utils = mu
