"""Stress: repeat whole-video scoring many times (fresh engines included) and count results that differ from the first one."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200")); sys.path.insert(0, ROOT)
import torch
from vcg_b200 import synthetic as W
from vcg_b200.engine import Engine
from oracle import two_stream_oracle as orc
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
T, L = 16, 100
sd = W.make_state_dict(T, "mlp", seed=123)
frames, scenes = W.make_video_u8(600, seed=123)
starts = W.clip_starts(600, T)
ids, mask = W.make_video_text(starts, scenes, T, L, seed=123)
pre = orc.preprocess_u8(frames)
imgs = [orc.gather_clips(pre, starts[b0:b0 + 16], T).cuda() for b0 in range(0, len(starts), 16)]
fr, idc, mkc = frames.cuda(), ids.cuda(), mask.cuda()
st = torch.tensor(starts, dtype=torch.int32).cuda()
ref = {}
bad = {"forward16": 0, "video_u8": 0, "clips_u8": 0}
t0 = time.time()
for rep in range(reps):
    if rep % 5 == 0:
        eng = Engine(T, "mlp", "bf16", vision=True, max_tokens=128, max_batch=32)
        eng.load_state_dict(sd)
    out = {"forward16": torch.cat([eng.forward(imgs[i], idc[16 * i:16 * i + 16], mkc[16 * i:16 * i + 16])[0] for i in range(len(imgs))]),
           "video_u8": eng.score_video_u8(fr, 0, 4, idc, mkc)[0].clone(),
           "clips_u8": eng.score_clips_u8(fr, st, idc, mkc)[0].clone()}
    torch.cuda.synchronize()
    print(f"rep {rep} done {time.time() - t0:.1f}s", flush=True)
    for k, v in out.items():
        if k not in ref:
            ref[k] = v.clone()
        elif not torch.equal(ref[k], v):
            bad[k] += 1
            d = (ref[k] - v).abs().max(1).values
            print(f"rep {rep} {k}: {int((d > 0).sum())} clips differ, max |diff| {float(d.max()):.3e}, clips {[int(i) for i in torch.nonzero(d > 0).flatten()[:8]]}", flush=True)
print(f"{reps} reps in {time.time() - t0:.0f} s; mismatching runs: {bad}  env: " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("VCG_")))
