"""Clips/s of TwoStream.forward on fp32 image tensors: fused engine (running statistics) vs the layer-by-layer
batch-statistics BatchNorm mode (vcg_b200/bn_batch.py).  python tools/bench_bn_batch.py [B] [precision]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from test_parity_gpu import build_model  # noqa: E402  (tests/ may import the oracle's synthetic weights)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
T, L = 16, 100
model, _ = build_model(T, "mlp", precision)
img = torch.randn(B, T, 3, 224, 224, device="cuda")
ids = torch.randint(1000, 30000, (B, L), device="cuda")
mask = torch.ones(B, L, dtype=torch.long, device="cuda")
for mode in (False, True):
    model.bn_batch_stats = mode
    for _ in range(2):
        model(img, ids, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        model(img, ids, mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{precision} B={B} bn_batch_stats={mode}: {ms:.2f} ms per forward = {B / ms * 1e3:.0f} clips/s", flush=True)
