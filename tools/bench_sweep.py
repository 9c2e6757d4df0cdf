"""BASELINE.json configs[4]: shape sweep of the whole pipeline (uint8 frames -> scores), T in {8,16,32} x L in {128,256,512}
x B in {1, 32, 256, 1024} clips on a regular stride-4 grid, bf16, one B200.  Prints one JSON line per point.
Usage: python tools/bench_sweep.py [reps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200")); sys.path.insert(0, ROOT)
import torch
from vcg_b200 import synthetic as W
from vcg_b200.engine import Engine

def flops(t, l):
    return 8.174272512e9 * t + 169869312.0 * l + 36864.0 * l * l + 1179648.0 + 2.0 * (768 * 128 + t * 2048 * 128 + (t + 1) * 128 * 2)

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for T in (8, 16, 32):
    sd = W.make_state_dict(T, "mlp", seed=123)
    eng = Engine(T, "mlp", "bf16", vision=True, max_tokens=512, max_batch=64 if T <= 16 else 32)
    eng.load_state_dict(sd)
    del sd
    for B in (1, 32, 256, 1024):
        n_frames = 4 * (B - 1) + T
        g = torch.Generator().manual_seed(B)
        frames = torch.randint(0, 256, (n_frames, 224, 224, 3), generator=g, dtype=torch.uint8).cuda()
        for L in (128, 256, 512):
            ids, mask = W.make_text(B, L, seed=B + L)
            ids, mask = ids.cuda(), mask.cuda()
            for _ in range(2):
                eng.score_video_u8(frames, 0, 4, ids, mask)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = reps if B >= 256 else reps * 4
            e0.record()
            for _ in range(n):
                eng.score_video_u8(frames, 0, 4, ids, mask)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            cps = B / ms * 1e3
            print(json.dumps({"T": T, "L": L, "B": B, "ms": round(ms, 3), "clips_per_s": round(cps, 1),
                              "tflops_algorithmic": round(cps * flops(T, L) / 1e12, 1),
                              "frac_of_sustained": round(cps * flops(T, L) / 1e12 / 1394.4, 3)}), flush=True)
        del frames
    eng.close()
    torch.cuda.empty_cache()
