"""Finds the first vision-stream kernel whose output differs between runs on identical inputs (VCG_DEBUG_CHECKSUM=1)."""
import os, sys, collections
os.environ["VCG_DEBUG_CHECKSUM"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200")); sys.path.insert(0, ROOT)
import torch
from vcg_b200 import synthetic as W
from vcg_b200.engine import Engine
from oracle import two_stream_oracle as orc
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
fresh_every = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T, L, B = 16, 100, 16
sd = {k: v.cuda() for k, v in W.make_state_dict(T, "mlp", seed=123).items()}
frames = W.make_frames_u8(4 * (B - 1) + T, seed=3)
ids, mask = W.make_text(B, L, seed=3)
img = orc.gather_clips(orc.preprocess_u8(frames), [4 * b for b in range(B)], T).cuda()
ids, mask = ids.cuda(), mask.cuda()
ref = None
firsts = collections.Counter()
eng = None
for rep in range(reps):
    if rep % fresh_every == 0:
        if eng is not None: eng.close()
        eng = Engine(T, "mlp", "bf16", vision=True, max_tokens=128, max_batch=32)
        eng.load_state_dict(sd)
    eng.forward(img, ids, mask)
    cs = eng.debug_checksums()
    if ref is None:
        ref = cs
        print(f"{len(cs)} checksums per pass", flush=True)
        continue
    diff = [i for i, (a, b) in enumerate(zip(ref, cs)) if a != b]
    if diff:
        firsts[diff[0]] += 1
        print(f"rep {rep} (call {rep % fresh_every} of its engine): first differing step {diff[0]}, {len(diff)} steps differ: {diff[:12]}", flush=True)
print("first-differing-step histogram:", dict(firsts), "of", reps, "runs")
