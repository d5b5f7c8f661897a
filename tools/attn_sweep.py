"""Sweep attention-kernel variants (env knobs) at the default-window shape; prints TFLOP/s and the error vs torch SDPA."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import math, sys, torch
sys.path.insert(0, %r)
from actionmesh_b200 import ops
B,H,S,D = 2,16,32784,128
g = torch.Generator().manual_seed(0)
buf = torch.randn(B,S,3*H*D, generator=g).cuda().bfloat16()
q = buf[:,:,:H*D].view(B,S,H,D); k = buf[:,:,H*D:2*H*D].view(B,S,H,D); v = buf[:,:,2*H*D:].view(B,S,H,D)
o = torch.empty(B,S,H,D, device="cuda", dtype=torch.bfloat16)
sc = 1/math.sqrt(D)
for _ in range(2): ops.flash_attn(q,k,v,o,sc)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.flash_attn(q,k,v,o,sc)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/5
ref = torch.nn.functional.scaled_dot_product_attention(q.permute(0,2,1,3),k.permute(0,2,1,3),v.permute(0,2,1,3)).permute(0,2,1,3)
err = float((o.float()-ref.float()).norm()/ref.float().norm())
# sharper distribution (x4 logits) to exercise the polynomial on a wide exponent range
q2 = (q.float()*4).bfloat16()
ops.flash_attn(q2,k,v,o,sc)
ref2 = torch.nn.functional.scaled_dot_product_attention(q2.permute(0,2,1,3),k.permute(0,2,1,3),v.permute(0,2,1,3)).permute(0,2,1,3)
err2 = float((o.float()-ref2.float()).norm()/ref2.float().norm())
print("RESULT", ms, 4.0*B*H*S*S*D/ms/1e9, err, err2)
''' % ROOT

for env in sys.argv[1:] or ["AMB_ATTN_EMU=0", "AMB_ATTN_EMU=4", "AMB_ATTN_EMU=3", "AMB_ATTN_EMU=2"]:
    e = dict(os.environ)
    for kv in env.split(","):
        k, v = kv.split("=")
        e[k] = v
    r = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, timeout=300)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    if line:
        _, ms, tf, err, err2 = line[0].split()
        print(f"{env:32s} {float(ms):8.3f} ms  {float(tf):8.1f} TFLOP/s  rel_err {float(err):.2e}  sharp {float(err2):.2e}", flush=True)
    else:
        print(env, "FAILED", (r.stdout + r.stderr)[-800:], flush=True)
