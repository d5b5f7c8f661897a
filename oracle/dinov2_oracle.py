"""TEST INFRASTRUCTURE ONLY — the reference's image encoder arithmetic: HF `transformers.Dinov2Model` in fp32
(actionmesh/model/image_encoder.py:25-55 minus the hub download).  transformers is installed (5.5.0; the reference
pins `<5`, requirements.txt:10 — the Dinov2 forward is unchanged across that boundary as far as this path goes)."""
from __future__ import annotations

import torch


def make_model(hidden_size=1024, num_layers=24, num_heads=16, seed=1235, layerscale=1.0, bf16_weights=True):
    from transformers import Dinov2Config, Dinov2Model

    torch.manual_seed(seed)
    cfg = Dinov2Config(hidden_size=hidden_size, num_hidden_layers=num_layers, num_attention_heads=num_heads,
                       image_size=518, patch_size=14, mlp_ratio=4, layerscale_value=layerscale)
    m = Dinov2Model(cfg).eval()
    rnd = (lambda t: t.to(torch.bfloat16).to(torch.float32)) if bf16_weights else (lambda t: t)
    with torch.no_grad():  # bf16_weights: bf16-representable weights, so a bf16-operand path consumes identical operands;
        for p_ in m.parameters():  # the fp32-grade path is tested with unrounded weights
            p_.copy_(rnd(p_))
        # make LayerScale / position embeddings non-trivial (HF inits them to constants / small noise)
        g = torch.Generator().manual_seed(seed + 1)
        for n, p_ in m.named_parameters():
            if "lambda1" in n:
                p_.copy_(rnd(torch.rand(p_.shape, generator=g) * 0.5 + 0.5))
            if "position_embeddings" in n or "cls_token" in n:
                p_.copy_(rnd(torch.randn(p_.shape, generator=g) * 0.2))
    return m


@torch.no_grad()
def encode(model, pixel_values: torch.Tensor) -> torch.Tensor:
    return model(pixel_values).last_hidden_state
