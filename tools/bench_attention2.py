"""Attention op time vs clip length (all clips the same length). Usage: python tools/bench_attention2.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200"))
import torch
from vcg_b200 import ops, binding as B
lib = B.load_library()
Bc = 256
for Lc in (8, 16, 32, 33, 64, 65, 96, 128):
    lens = torch.full((Bc,), Lc)
    cu = torch.zeros(Bc + 1, dtype=torch.int32); cu[1:] = torch.cumsum(lens, 0)
    total = int(cu[-1]); rows = Bc * 128 + 128
    q = torch.randn(rows, 2304, device="cuda").to(torch.bfloat16)
    ok = torch.ones(rows, dtype=torch.uint8, device="cuda"); cud = cu.cuda()
    ctx = torch.zeros(rows, 768, dtype=torch.bfloat16, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    def run():
        B.check(lib.vcg_op_bert_attention_packed(q.data_ptr(), cud.data_ptr(), ok.data_ptr(), ctx.data_ptr(), Bc, 128, rows, s))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for _ in range(n): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"L={Lc:4d} tokens={total:6d}: {us:6.1f} us  ({us/total*1e3:6.2f} ns/token)", flush=True)
