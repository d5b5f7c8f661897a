"""The block's to_out GEMM exactly as the fp32-stream denoiser launches it (ncu target + timing):
(65568, 2048) x (2048, 2048)^T + bias + fp32 residual (aliasing the fp32 output) + bf16 copy."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from actionmesh_b200 import ops
M, N, K = 65568, 2048, 2048
g = torch.Generator(device="cuda").manual_seed(0)
a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
w = (torch.randn(N, K, device="cuda", generator=g) * 0.02).bfloat16()
b = torch.randn(N, device="cuda", generator=g)
h = torch.randn(M, N, device="cuda", generator=g)
hb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
variants = {"fp32_res_alias_out2": dict(out=h, residual=h, out2=hb), "fp32_res_alias": dict(out=h, residual=h),
            "bf16_out_plain": dict(out=hb)}
only = os.environ.get("ONLY")
for name, kw in variants.items():
    if only and name != only:
        continue
    for _ in range(3):
        ops.gemm(a, w, bias=b, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(a, w, bias=b, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"GEMM_ONE {name}: {ms:.4f} ms = {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s")
