"""Runs the same scoring calls several times and compares the logits bit for bit (race detector for the tcgen05 pipelines)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200")); sys.path.insert(0, ROOT)
import torch
from vcg_b200 import synthetic as W
from vcg_b200.engine import Engine
from oracle import two_stream_oracle as orc
T, L = 16, 100
sd = W.make_state_dict(T, "mlp", seed=123)
frames, scenes = W.make_video_u8(600, seed=123)
starts = W.clip_starts(600, T)
ids, mask = W.make_video_text(starts, scenes, T, L, seed=123)
pre = orc.preprocess_u8(frames)
for chunk in (16, 32, 64):
    eng = Engine(T, "mlp", "bf16", vision=True, max_tokens=L, max_batch=chunk)
    eng.load_state_dict(sd)
    st = torch.tensor(starts, dtype=torch.int32)
    outs = {"forward16": [], "video_u8": [], "clips_u8": []}
    for rep in range(4):
        lg = []
        for b0 in range(0, len(starts), 16):
            sl = slice(b0, min(b0 + 16, len(starts)))
            img = orc.gather_clips(pre, starts[sl], T).cuda()
            lg.append(eng.forward(img, ids[sl].cuda(), mask[sl].cuda())[0].clone())
        outs["forward16"].append(torch.cat(lg))
        outs["video_u8"].append(eng.score_video_u8(frames.cuda(), 0, 4, ids.cuda(), mask.cuda())[0].clone())
        outs["clips_u8"].append(eng.score_clips_u8(frames.cuda(), st.cuda(), ids.cuda(), mask.cuda())[0].clone())
    torch.cuda.synchronize()
    for k, v in outs.items():
        same = all(torch.equal(v[0], x) for x in v[1:])
        dmax = max(float((v[0] - x).abs().max()) for x in v[1:])
        print(f"chunk {chunk} {k}: identical across 4 runs = {same} (max |diff| {dmax:.3e})")
    print("  forward16 vs clips_u8 max |diff|", float((outs["forward16"][0] - outs["clips_u8"][0]).abs().max()))
    eng.close()
