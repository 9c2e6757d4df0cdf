"""Times the token-packed attention op alone (CUDA events).  Usage: python tools/bench_attention.py [B] [Lmax] [nbuf]
nbuf = 1: the same 65 MB QKV matrix every call (L2-warm); nbuf = 4: rotating buffers (L2-cold)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200"))
import torch
from vcg_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 100
g = torch.Generator().manual_seed(0)
for mode in ("ragged", "full"):
    lens = torch.randint(10, Lmax + 1, (B,), generator=g) if mode == "ragged" else torch.full((B,), Lmax)
    cu = torch.zeros(B + 1, dtype=torch.int32); cu[1:] = torch.cumsum(lens, 0)
    total = int(cu[-1]); rows = B * Lmax + 128
    for nbuf in (1, 4):
        qs = [torch.randn(rows, 2304, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
        ok = torch.ones(rows, dtype=torch.uint8, device="cuda"); cud = cu.cuda()
        for i in range(3): ops.bert_attention_packed(qs[i % nbuf], cud, ok, Lmax)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for i in range(n): ops.bert_attention_packed(qs[i % nbuf], cud, ok, Lmax)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        fl = 4.0 * 12 * 64 * float((lens.double() ** 2).sum())
        print(f"{mode:7s} B={B} Lmax={Lmax} tokens={total} nbuf={nbuf}: {us:7.1f} us  {fl/us/1e6:7.1f} TFLOP/s (executed)  {total*(2304+768)*2/us/1e3:7.1f} GB/s", flush=True)
