"""Which stream glitches?  Repeats the text stream (precomputed embeddings) and the vision stream separately and compares
lang_emb / vision_emb bit for bit with the first result."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200")); sys.path.insert(0, ROOT)
import torch
from vcg_b200 import synthetic as W
from vcg_b200.engine import Engine
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
T, L = 16, 100
sd = W.make_state_dict(T, "mlp", seed=123)
eng = Engine(T, "mlp", "bf16", vision=True, max_tokens=128, max_batch=32)
eng.load_state_dict(sd)
for B in (16, 146):
    emb, ids, mask = W.make_precomputed_inputs(B, T, L, seed=7)
    emb, ids, mask = emb.cuda(), ids.cuda(), mask.cuda()
    ref = None; bad = 0
    for rep in range(reps):
        lg, pr, ve, le = eng.forward(None, ids, mask, return_emb=True, vision_emb=emb)
        torch.cuda.synchronize()
        if ref is None: ref = le.clone()
        elif not torch.equal(ref, le):
            bad += 1
            d = (ref - le).abs().max(1).values
            print(f"text B={B} rep {rep}: clips {[int(i) for i in torch.nonzero(d > 0).flatten()[:10]]} max diff {float(d.max()):.3e}", flush=True)
    print(f"text stream B={B}: {bad} of {reps} runs differ", flush=True)
frames = W.make_frames_u8(4 * 31 + T, seed=3).cuda()
ids, mask = W.make_text(32, L, seed=3); ids, mask = ids.cuda(), mask.cuda()
e2 = Engine(T, "mlp", "bf16", vision=True, max_tokens=128, max_batch=32, modality="embed")
e2.load_state_dict(sd)
ref = None; bad = 0
for rep in range(reps):
    ve, le = e2.embed_u8(frames, ids, mask, clip_start=None, first_start=0, clip_stride=4)
    torch.cuda.synchronize()
    if ref is None: ref = ve.clone()
    elif not torch.equal(ref, ve):
        bad += 1
        d = (ref - ve).abs().amax(dim=(1, 2))
        print(f"vision rep {rep}: clips {[int(i) for i in torch.nonzero(d > 0).flatten()[:10]]} max diff {float(d.max()):.3e}", flush=True)
print(f"vision stream (32 clips, shared stem): {bad} of {reps} runs differ")
