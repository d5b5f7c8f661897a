"""Dump the clock64 role timeline of attention CTA (0,0,0) (v4 kernel): where each role waits.  AMB_ATTN_VER=4."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from actionmesh_b200 import ops, _lib
B, H, S, D = 1, 16, 32784, 128
buf = torch.randn(B, S, 3 * H * D).cuda().bfloat16()
q = buf[:, :, :H * D].view(B, S, H, D); k = buf[:, :, H * D:2 * H * D].view(B, S, H, D); v = buf[:, :, 2 * H * D:].view(B, S, H, D)
o = torch.empty(B, S, H, D, device="cuda", dtype=torch.bfloat16)
tr = torch.zeros(5 * 16 * 8, dtype=torch.int64, device="cuda")
ops.flash_attn(q, k, v, o, 1 / math.sqrt(D))
_lib.load_library().amb_debug_set_attn_trace(tr.data_ptr())
ops.flash_attn(q, k, v, o, 1 / math.sqrt(D))
torch.cuda.synchronize()
_lib.load_library().amb_debug_set_attn_trace(None)
t = tr.cpu().view(5, 16, 8)
t0 = int(t[4, 0, 0])
names = ["t0h0", "t0h1", "t1h0", "t1h1", "mma"]
print("softmax events: 0 enter, 1 s_full passed, 2 pass1 done, 3 max exchanged, 4 exps done, 5 arrived")
print("mma events: 0 p_a0, 1 p_b0, 2 QK0' issued, 3 p_a1, 4 p_b1, 5 QK1' issued")
for j in range(6):
    for r in range(5):
        ev = [int(x) - t0 for x in t[r, j, :6]]
        print(f"j={100 + j} {names[r]:5s} " + " ".join(f"{e:7d}" for e in ev))
    print()
# averages
import statistics
for r in range(4):
    d = t[r, 1:15]
    print(names[r], "wait_s", statistics.mean((d[:, 1] - d[:, 0]).tolist()), "pass1", statistics.mean((d[:, 2] - d[:, 1]).tolist()),
          "xchg", statistics.mean((d[:, 3] - d[:, 2]).tolist()), "exps", statistics.mean((d[:, 4] - d[:, 3]).tolist()),
          "arrive", statistics.mean((d[:, 5] - d[:, 4]).tolist()))
d = t[4, 1:15]
print("mma period", statistics.mean((d[1:, 0] - d[:-1, 0]).tolist()), "pa0->pb0", statistics.mean((d[:, 1] - d[:, 0]).tolist()),
      "pb0->qk0", statistics.mean((d[:, 2] - d[:, 1]).tolist()), "qk0->pa1", statistics.mean((d[:, 3] - d[:, 2]).tolist()),
      "pa1->pb1", statistics.mean((d[:, 4] - d[:, 3]).tolist()), "pb1->qk1", statistics.mean((d[:, 5] - d[:, 4]).tolist()))
