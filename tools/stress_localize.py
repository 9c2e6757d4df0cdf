"""Localises the run-to-run glitch: repeated TwoStream forward of the same 16 clips, comparing vision_emb, lang_emb and logits
separately with the first run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200")); sys.path.insert(0, ROOT)
import torch
from vcg_b200 import synthetic as W
from vcg_b200.engine import Engine
from oracle import two_stream_oracle as orc
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
T, L, B = 16, 100, 16
sd = W.make_state_dict(T, "mlp", seed=123)
frames = W.make_frames_u8(4 * (B - 1) + T, seed=3)
ids, mask = W.make_text(B, L, seed=3)
img = orc.gather_clips(orc.preprocess_u8(frames), [4 * b for b in range(B)], T).cuda()
ids, mask = ids.cuda(), mask.cuda()
eng = Engine(T, "mlp", "bf16", vision=True, max_tokens=128, max_batch=32)
eng.load_state_dict(sd)
ref = None
bad = {"logits": 0, "vision_emb": 0, "lang_emb": 0}
for rep in range(reps):
    lg, pr, ve, le = eng.forward(img, ids, mask, return_emb=True)
    torch.cuda.synchronize()
    cur = {"logits": lg.clone(), "vision_emb": ve.clone(), "lang_emb": le.clone()}
    if ref is None:
        ref = cur
        continue
    msg = []
    for k in cur:
        if not torch.equal(ref[k], cur[k]):
            bad[k] += 1
            d = (ref[k] - cur[k]).abs().reshape(B, -1).max(1).values
            msg.append(f"{k}: clips {[int(i) for i in torch.nonzero(d > 0).flatten()]} max {float(d.max()):.2e}")
            if k == "vision_emb":
                dd = (ref[k] - cur[k]).abs().reshape(B * T, -1)
                fr = torch.nonzero(dd.max(1).values > 0).flatten()
                ch = torch.nonzero(dd.max(0).values > 0).flatten()
                msg.append(f"frames {fr.tolist()[:40]} n_channels {len(ch)} first {ch[:6].tolist()} last {ch[-3:].tolist()}")
    if msg:
        print(f"rep {rep}: " + " | ".join(msg), flush=True)
print(f"{reps} reps: runs differing {bad}")
