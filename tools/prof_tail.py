"""Runs the fused bottleneck tail of layer1 a few times (for ncu / timing).  Usage: python tools/prof_tail.py <variant> [frames] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200"))
import torch
from vcg_b200 import ops
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
P = int(os.environ.get("P", "64"))
dev, H, T = "cuda", (56 if P == 64 else 28), 16
xs = [torch.randn(n, H, H, P, device=dev).to(torch.bfloat16) for _ in range(3)]
rs = [torch.randn(n, H, H, 4 * P, device=dev).to(torch.bfloat16) for _ in range(3)]
w2 = (torch.randn(P, 3, 3, P, device=dev) / (9 * P) ** 0.5).to(torch.bfloat16)
w3 = (torch.randn(4 * P, P, device=dev) / P ** 0.5).to(torch.bfloat16)
b2 = torch.randn(P, device=dev) * 0.1
b3 = torch.randn(4 * P, device=dev) * 0.1
tsm = torch.zeros(n, H, H, 64, device=dev, dtype=torch.bfloat16) if P == 64 else None
for i in range(3):
    ops.bottleneck_tail(xs[i % 3], w2, b2, w3, b3, rs[i % 3], 1, tsm_out=tsm, tsm_fold=32 if P == 64 else 0, clip_frames=T, variant=variant)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters):
    ops.bottleneck_tail(xs[i % 3], w2, b2, w3, b3, rs[i % 3], 1, tsm_out=tsm, tsm_fold=32 if P == 64 else 0, clip_frames=T, variant=variant)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
byts = n * H * H * (P + 4 * P * 2 + (64 if P == 64 else 0)) * 2
flops = 2.0 * n * H * H * (9 * P * P + 4 * P * P)
print(f"variant {variant}: {n} frames {ms*1e3:.1f} us  {byts/ms/1e6:.0f} GB/s algorithmic  {flops/ms/1e9:.0f} TFLOP/s")
