"""Dump the clock64 role timeline of attention CTA (0,0,0) of the CTA-pair kernel: where each role waits."""
import math, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from actionmesh_b200 import ops, _lib
if os.environ.get("AMB_PROBE_LIB"):  # a build with -DAMB_ATTN_TRACE=1 (tools/build_variant.sh)
    _lib.LIB_PATH = os.path.abspath(os.environ["AMB_PROBE_LIB"])
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
print("softmax (warp 0 of set 0 / set 1): 0 enter, 1 s_full passed, 2 half0 loaded, 3 half0 handed over, 4 half1 loaded, 5 half1 handed over")
print("mma: 0 iteration start, 1 QK(j+2) issued, 2 v_full passed, 3 p_ready passed, 4 PV(j) issued")
for j in range(5):
    for r, nme in ((0, "sm_w0"), (1, "sm_w4"), (4, "mma")):
        print(f"j={100 + j} {nme:6s} " + " ".join(f"{int(x) - t0:7d}" for x in t[r, j, :6]))
    print()
if os.environ.get("LAYOUT", "1") == "1":   # key-half layout: role r = key half, every tile; its slots are 0, 1, 2+2r, 3+2r
    for r, nme in ((0, "keys0-63", ), (1, "keys64-127", )):
        d = t[r, 1:15]
        print(nme, "period", statistics.mean((d[1:, 0] - d[:-1, 0]).tolist()), "wait_s", statistics.mean((d[:, 1] - d[:, 0]).tolist()),
              "load", statistics.mean((d[:, 2 + 2 * r] - d[:, 1]).tolist()), "half", statistics.mean((d[:, 3 + 2 * r] - d[:, 2 + 2 * r]).tolist()))
else:
    for r, nme, par in ((0, "set0", 0), (1, "set1", 1)):
        d = t[r, 2 + par:14:2]     # the set's own tiles (every second one)
        print(nme, "period(2 tiles)", statistics.mean((d[1:, 0] - d[:-1, 0]).tolist()), "wait_s", statistics.mean((d[:, 1] - d[:, 0]).tolist()),
              "load0", statistics.mean((d[:, 2] - d[:, 1]).tolist()), "half0", statistics.mean((d[:, 3] - d[:, 2]).tolist()),
              "load1", statistics.mean((d[:, 4] - d[:, 3]).tolist()), "half1", statistics.mean((d[:, 5] - d[:, 4]).tolist()))
d = t[4, 1:15]
print("mma period", statistics.mean((d[1:, 0] - d[:-1, 0]).tolist()), "k_wait+qk_issue", statistics.mean((d[:, 1] - d[:, 0]).tolist()),
      "v_wait", statistics.mean((d[:, 2] - d[:, 1]).tolist()), "p_wait", statistics.mean((d[:, 3] - d[:, 2]).tolist()),
      "pv_issue", statistics.mean((d[:, 4] - d[:, 3]).tolist()))
