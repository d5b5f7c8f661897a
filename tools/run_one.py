"""Launch one hot kernel at the default-window shape (for `ncu --set full` captures under gpurun).

    python tools/run_one.py attn|attn_small|gemm_ff1|gemm_qkv|stage2 [repeats]
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from actionmesh_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "attn"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = "cuda"
g = torch.Generator().manual_seed(0)
if which in ("attn", "attn_small"):
    B, H, D = 2, 16, 128
    S = 32784 if which == "attn" else 4098
    buf = torch.randn(B, S, 3 * H * D, generator=g).to(dev).bfloat16()
    q = buf[:, :, :H * D].view(B, S, H, D)
    k = buf[:, :, H * D:2 * H * D].view(B, S, H, D)
    v = buf[:, :, 2 * H * D:].view(B, S, H, D)
    o = torch.empty(B, S, H, D, device=dev, dtype=torch.bfloat16)
    for _ in range(reps):
        ops.flash_attn(q, k, v, o, 1 / math.sqrt(D))
elif which == "gemm_ff1":
    a = torch.randn(65568, 2048, generator=g).to(dev).bfloat16()
    w = (torch.randn(8192, 2048, generator=g) / 45).to(dev).bfloat16()
    b = torch.zeros(8192, device=dev)
    c = torch.empty(65568, 8192, device=dev, dtype=torch.bfloat16)
    for _ in range(reps):
        ops.gemm(a, w, c, bias=b, act=1)
elif which == "gemm_qkv":
    a = torch.randn(65568, 2048, generator=g).to(dev).bfloat16()
    w = (torch.randn(6144, 2048, generator=g) / 45).to(dev).bfloat16()
    c = torch.empty(65568, 6144, device=dev, dtype=torch.bfloat16)
    w0 = torch.ones(128, device=dev)
    cos = torch.ones(32, 64, device=dev)
    sin = torch.zeros(32, 64, device=dev)
    for _ in range(reps):
        ops.gemm(a, w, c, norm=dict(cols=4096, seg=2048, w0=w0, w1=w0, eps=1e-6, rope_cols=4096, cos=cos, sin=sin, rows_per_pos=2049))
elif which == "stage2":  # the two HBM-bound kernels of the Stage-II fp32-grade query path at the default shape
    R, Rp, D = 32784, 32832, 1024
    s32 = torch.randn(2048, Rp, generator=g).to(dev)
    p3 = torch.empty(2048, 3 * Rp, device=dev, dtype=torch.bfloat16)
    x = torch.randn(R, D, generator=g).to(dev)
    x3 = torch.empty(R, 3 * D, device=dev, dtype=torch.bfloat16)
    for _ in range(reps):
        ops.softmax_split3(s32, R, 1 / math.sqrt(128), p3)
        ops.split3(x, x3)
torch.cuda.synchronize()
print("done", which)
