"""Structural CPU test of the denoiser's launch programs (no GPU, no arithmetic): `B200Denoiser._forward_packed` (single
GPU) and the staggered per-branch programs of the frame-sharded window are executed against shape/dtype-checking fakes of
the C-ABI wrappers in actionmesh_b200.ops.  Catches slicing / buffer-plumbing / generator-flow mistakes in the host code
that would otherwise only show on a GPU box; the numerics are covered by the -m gpu parity tests."""
import pytest
import torch

from actionmesh_b200 import denoiser as dn
from actionmesh_b200 import ops
from oracle import synth


class _Recorder:
    def __init__(self):
        self.calls = []

    def gemm(self, a, w, out, *, bias=None, a2=None, residual=None, act=0, col_scale=None, row_map=None, norm=None, out2=None, tag="gemm"):
        k = a.shape[1] + (a2.shape[1] if a2 is not None else 0)
        assert a.dtype == w.dtype == torch.bfloat16 and k == w.shape[1] and out.shape[1] == w.shape[0], (a.shape, w.shape, out.shape)
        assert a.stride(1) == 1 and w.stride(1) == 1 and out.stride(1) == 1
        if a2 is not None:
            assert a2.shape[0] == a.shape[0]
        if row_map is None:
            assert out.shape[0] >= a.shape[0]
        else:
            grp, stride, off = row_map
            last = (a.shape[0] - 1) // grp * stride + (a.shape[0] - 1) % grp + off
            assert last < out.shape[0], (row_map, a.shape, out.shape)
        if residual is not None:
            assert residual.shape[1] == out.shape[1] and residual.shape[0] >= a.shape[0]
        if bias is not None:
            assert bias.dtype == torch.float32 and bias.numel() == w.shape[0]
        if norm is not None and norm.get("rope_cols", 0):
            n_pos = (a.shape[0] + norm["rows_per_pos"] - 1) // norm["rows_per_pos"]
            assert norm["cos"].shape[0] >= n_pos and norm["cos"].shape == norm["sin"].shape, (norm["cos"].shape, n_pos)
        if out2 is not None:
            assert out2.dtype == torch.bfloat16 and out2.shape == out.shape and out.dtype == torch.float32
            self.calls.append(("gemm_out2", tuple(out2.shape)))
        self.calls.append(("gemm", tuple(a.shape), tuple(w.shape)))
        return out

    def layernorm(self, x, gamma, beta, eps, out=None):
        assert out is not None and out.shape == x.shape and gamma.numel() == x.shape[1]
        self.calls.append(("ln", tuple(x.shape)))
        return out

    def flash_attn(self, q, k, v, o, scale, kv_chunks=1, tag="attn"):
        if kv_chunks == 1:
            assert q.dim() == k.dim() == 4 and k.shape == v.shape and q.shape[0] == k.shape[0] and q.shape[2:] == k.shape[2:]
        else:
            assert k.dim() == 5 and k.shape[1] == kv_chunks and k.shape == v.shape and q.shape[2:] == k.shape[3:]
        assert o.shape == q.shape
        self.calls.append((tag, tuple(q.shape), tuple(k.shape)))
        return o

    def timestep_embedding(self, t, channels, out=None, mask=None, rows=None):
        assert out.shape == (rows, channels)
        return out

    def add_bias_rows(self, y, bias):
        assert bias.numel() == y.shape[1]
        self.calls.append(("bias_rows", tuple(y.shape)))

    def cast_bf16(self, src, out=None):
        if out is None:
            out = torch.empty(src.shape, dtype=torch.bfloat16)
        assert out.numel() == src.numel() and src.dtype == torch.float32 and out.dtype == torch.bfloat16
        self.calls.append(("cast", tuple(src.shape)))
        return out


def _model(monkeypatch, residual_fp32=True):
    rec = _Recorder()
    for name in ("gemm", "layernorm", "flash_attn", "timestep_embedding", "add_bias_rows", "cast_bf16"):
        monkeypatch.setattr(ops, name, getattr(rec, name))
    d = dict(num_layers=5, num_attention_heads=2, width=256, cross_attention_dim=128, in_channels=64, mlp_ratio=4.0)
    cfg = dn.DenoiserConfig(inflated_layers=(0, 1, 2, 3, 4), **d)
    m = dn.B200Denoiser(cfg, residual_fp32=residual_fp32)
    m._w = m._pack_state_dict(synth.make_state_dict(cfg, 1), torch.device("cpu"))
    m._loaded = True
    return m, rec, cfg


@pytest.mark.parametrize("residual_fp32", [True, False])
def test_single_gpu_program(monkeypatch, residual_fp32):
    m, rec, cfg = _model(monkeypatch, residual_fp32)
    B, T, N = 2, 4, 31
    ctx = torch.randn(B, T, 9, 128)
    ctx[0] = 0
    fs = torch.arange(T, dtype=torch.float32)[None].repeat(B, 1)
    st = m.precompute_window(ctx, fs, N)
    assert st.ctx_zero == [True, False]
    ws = m._workspace(B, T, N)
    pred = m._forward_packed(ws, st, B, T, N, torch.tensor([500.0]), torch.zeros(B * T), n_input_branches=1)
    assert pred.shape == (B * T * (N + 1), 64)
    attn = [c for c in rec.calls if c[0] == "attn_self"]
    assert len(attn) == cfg.num_layers and all(c[1] == (B, T * (N + 1), 2, 128) for c in attn)
    assert sum(1 for c in rec.calls if c[0] == "attn_cross") == cfg.num_layers      # only the non-zero-context branch
    assert sum(1 for c in rec.calls if c[0] == "bias_rows") == cfg.num_layers
    # fp32 residual stream: one bf16 operand copy per skip push and per skip pop, written as the second output of the
    # producing GEMM (no separate cast pass); none with the bf16 stream
    assert not [c for c in rec.calls if c[0] == "cast" and c[1] == (B * T * (N + 1), cfg.width)]
    assert len([c for c in rec.calls if c[0] == "gemm_out2"]) == (2 * (cfg.num_layers // 2) if residual_fp32 else 0)
    assert ws["h"].dtype == (torch.float32 if residual_fp32 else torch.bfloat16)


def test_sharded_branch_programs_interleave(monkeypatch):
    m, rec, cfg = _model(monkeypatch)
    order = []

    class _Work:
        def __init__(self, tag):
            self.tag = tag

        def wait(self):
            order.append(("wait", self.tag))

    world, B, T_all, N = 2, 2, 4, 31
    T = T_all // world

    class Shard:
        group = None

        @staticmethod
        def all_gather_kv(out, inp, channel=0):
            assert out.shape[0] == world * inp.shape[0] and out.shape[1] == inp.shape[1]
            order.append(("gather", out.data_ptr()))
            return _Work(out.data_ptr())

    Shard.world, Shard.rank = world, 0
    ctx = torch.randn(B, T_all, 9, 128)
    ctx[0] = 0
    fs = torch.arange(T_all, dtype=torch.float32)[None].repeat(B, 1)
    st = m.precompute_window(ctx, fs, N, frame_slice=slice(0, T))
    ws = m._workspace(B, T, N, world=world)
    pred = m._forward_packed(ws, st, B, T, N, torch.tensor([500.0]), torch.zeros(B * T), n_input_branches=1, shard=Shard)
    assert pred.shape == (B * T * (N + 1), 64)
    gathers = [o for o in order if o[0] == "gather"]
    assert len(gathers) == B * cfg.num_layers
    # staggering: between a branch's gather and its wait, the OTHER branch's gather/wait is issued (except at the very start)
    tags = [o for o in order]
    for i in range(len(tags) - 1):
        if tags[i][0] == "gather" and i > 0:
            assert not (tags[i + 1][0] == "wait" and tags[i + 1][1] == tags[i][1]), "a gather was waited on immediately"
    attn = [c for c in rec.calls if c[0] == "attn_self"]
    assert len(attn) == B * cfg.num_layers and all(c[1] == (1, T * (N + 1), 2, 128) and c[2][:3] == (1, world, T * (N + 1)) for c in attn)
