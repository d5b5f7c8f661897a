"""Seam 1 (-m gpu): B200AttentionProcessor plugged into a diffusers-style `Attention` container (the oracle's shim of
diffusers, which is what the reference's block.py constructs) against the fp32 restatement of
attention_processor.py:36-168.  bf16 tolerance as in test_kernels_gpu.py, two GEMM roundings deep."""
import pytest
import torch

from oracle import denoiser_oracle as do
from oracle import diffusers_shim as ds

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())


def _bf16r(m):
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(p.to(torch.bfloat16).float())


@pytest.mark.parametrize("inflate", [True, False])
def test_self_attention_processor(amb_lib, inflate):
    from actionmesh_b200.attention_processor import B200AttentionProcessor

    torch.manual_seed(0)
    D, H, T, L, B = 256, 2, 3, 33, 2
    attn = ds.Attention(query_dim=D, heads=H, dim_head=D // H, qk_norm="rms_norm", eps=1e-6, bias=False,
                        processor=B200AttentionProcessor())
    with torch.no_grad():
        attn.norm_q.weight.uniform_(0.8, 1.2)
        attn.norm_k.weight.uniform_(0.8, 1.2)
    _bf16r(attn)
    x = torch.randn(B * T, L, D).to(torch.bfloat16).float()
    pos = torch.arange(T, dtype=torch.float32).repeat(B)
    cos, sin = do.rotary_tables(D // H, pos)
    rope = (cos[:, None].repeat(1, L, 1), sin[:, None].repeat(1, L, 1))
    sd = {"a." + k: v.detach() for k, v in attn.state_dict().items()}
    ref = do.attention(sd, "a.", x, H, inflate_frames=T if inflate else None, rope=rope)
    attn = attn.cuda()
    out = attn(x.cuda(), n_frames=T, inflate_self_attention=inflate, freqs_rot=(rope[0].cuda(), rope[1].cuda()))
    assert out.shape == x.shape and out.dtype == x.dtype
    assert rel(out, ref) < 6e-3


def test_cross_attention_processor(amb_lib):
    from actionmesh_b200.attention_processor import B200AttentionProcessor

    torch.manual_seed(1)
    D, H, BT, L, S, Dc = 256, 2, 4, 33, 9, 128
    attn = ds.Attention(query_dim=D, cross_attention_dim=Dc, heads=H, dim_head=D // H, qk_norm="rms_norm", eps=1e-6,
                        bias=False, processor=B200AttentionProcessor())
    _bf16r(attn)
    x = torch.randn(BT, L, D).to(torch.bfloat16).float()
    ctx = torch.randn(BT, S, Dc).to(torch.bfloat16).float()
    sd = {"a." + k: v.detach() for k, v in attn.state_dict().items()}
    ref = do.attention(sd, "a.", x, H, context=ctx)
    attn = attn.cuda()
    out = attn(x.cuda(), encoder_hidden_states=ctx.cuda())
    assert rel(out, ref) < 6e-3
    out_bf = attn(x.cuda().bfloat16(), encoder_hidden_states=ctx.cuda().bfloat16())
    assert out_bf.dtype == torch.bfloat16 and rel(out_bf, ref) < 8e-3
