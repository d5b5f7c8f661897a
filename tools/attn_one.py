"""One launch of the attention kernel at a mid-size shape (ncu target)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from actionmesh_b200 import ops
B, H, S, D = 1, 16, int(os.environ.get("S", 16392)), 128
buf = torch.randn(B, S, 3 * H * D).cuda().bfloat16()
q = buf[:, :, :H * D].view(B, S, H, D); k = buf[:, :, H * D:2 * H * D].view(B, S, H, D); v = buf[:, :, 2 * H * D:].view(B, S, H, D)
o = torch.empty(B, S, H, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.flash_attn(q, k, v, o, 1 / math.sqrt(D))
torch.cuda.synchronize()
