"""Whole-pipeline throughput (configs[2] style) for a given vision chunk size, with per-layer CUDA-event profile.
Usage: python tools/bench_pipeline.py <n_frames> <chunk1,chunk2,...>"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200")); sys.path.insert(0, ROOT)
import torch
from vcg_b200 import synthetic as W
from vcg_b200.engine import Engine
T, L = 16, 100
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
chunks = [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "32").split(",")]
starts = W.clip_starts(n_frames, T); B = len(starts)
sd = W.make_state_dict(T, "mlp", seed=123)
g = torch.Generator().manual_seed(5)
frames = torch.randint(0, 256, (n_frames, 224, 224, 3), generator=g, dtype=torch.uint8).cuda()
ids, mask = W.make_text(B, L, seed=5); ids, mask = ids.cuda(), mask.cuda()
st = torch.tensor(starts, dtype=torch.int32).cuda()
for chunk in chunks:
    eng = Engine(T, "mlp", "bf16", vision=True, max_tokens=128, max_batch=chunk)
    eng.load_state_dict(sd)
    for _ in range(2): eng.score_video_u8(frames, 0, 4, ids, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); n = 3
    for _ in range(n): eng.score_video_u8(frames, 0, 4, ids, mask)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    eng.profile_begin(); eng.score_video_u8(frames, 0, 4, ids, mask); prof = eng.profile_end()
    lay = {}
    for r in prof:
        k = lay.setdefault(r["layer"], [0.0, 0.0, 0]); k[0] += r["ms"]; k[1] += r["flops"]; k[2] += r["launches"]
    tot = sum(v[0] for v in lay.values())
    print(f"chunk {chunk}: {B} clips {ms:.1f} ms -> {B/ms*1e3:.0f} clips/s (sum of kernels {tot:.1f} ms)")
    for name, (m, f, c) in sorted(lay.items(), key=lambda kv: -kv[1][0]):
        print(f"   {name:16s} {m:8.2f} ms {m/tot*100:5.1f}%  {f/m/1e9 if f else 0:8.1f} TF/s  {c} launches")
    eng.close()
