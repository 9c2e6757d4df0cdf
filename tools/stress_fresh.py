"""First-run glitch hunt: a FRESH engine per repetition, one forward of the same 16 clips, vision_emb compared by hash."""
import os, sys, hashlib, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200")); sys.path.insert(0, ROOT)
import torch
from vcg_b200 import synthetic as W
from vcg_b200.engine import Engine
from oracle import two_stream_oracle as orc
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
T, L, B = 16, 100, 16
sd = {k: v.cuda() for k, v in W.make_state_dict(T, "mlp", seed=123).items()}
frames = W.make_frames_u8(4 * (B - 1) + T, seed=3)
ids, mask = W.make_text(B, L, seed=3)
img = orc.gather_clips(orc.preprocess_u8(frames), [4 * b for b in range(B)], T).cuda()
ids, mask = ids.cuda(), mask.cuda()
seen = collections.Counter()
first = {}
for rep in range(reps):
    eng = Engine(T, "mlp", "bf16", vision=True, max_tokens=128, max_batch=32)
    eng.load_state_dict(sd)
    for call in range(2):
        lg, pr, ve, le = eng.forward(img, ids, mask, return_emb=True)
        torch.cuda.synchronize()
        h = hashlib.md5(ve.cpu().numpy().tobytes()).hexdigest()[:8]
        seen[(call, h)] += 1
        first.setdefault(h, ve.clone())
    eng.close()
print("distinct vision_emb results (call index, hash): count ->", dict(seen))
hs = list(first)
for h in hs[1:]:
    d = (first[hs[0]] - first[h]).abs().reshape(B * T, -1)
    fr = torch.nonzero(d.max(1).values > 0).flatten().tolist()
    print(f"  {hs[0]} vs {h}: frames {fr[:48]} max {float(d.max()):.2e}")
print("env:", " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("VCG_")))
