import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import test_video_parity_gpu as t
from oracle import two_stream_oracle as orc
from vcg_b200.engine import Engine
gd = os.path.join(ROOT, "tests", "golden")
g, T, L, B, sd, frames, starts, ids, mask = t._video_case(gd)
model = t._build_mirror_model(sd, T, "bf16")
pre = orc.preprocess_u8(frames)
out = []
for b0 in range(0, B, 16):
    sl = slice(b0, min(b0 + 16, B))
    img = orc.gather_clips(pre, starts[sl], T).cuda()
    out.append(model(img, ids[sl].cuda(), mask[sl].cuda())[0])
A = torch.cat(out).cpu()
eng = Engine(T, "mlp", "bf16", vision=True, max_tokens=L, max_batch=32)
eng.load_state_dict(sd)
st = torch.tensor(starts, dtype=torch.int32)
C = eng.score_clips_u8(frames.cuda(), st.cuda(), ids.cuda(), mask.cuda())[0].cpu()
F = torch.cat([eng.forward(orc.gather_clips(pre, starts[b0:b0 + 16], T).cuda(), ids[b0:b0 + 16].cuda(), mask[b0:b0 + 16].cuda())[0] for b0 in range(0, B, 16)]).cpu()
ref = torch.from_numpy(g["logits"])
def merr(x): return float(((x[:, 1] - x[:, 0]) - (ref[:, 1] - ref[:, 0])).abs().max())
print("flow A (mirror) margin err", merr(A), "| engine forward16", merr(F), "| engine clips_u8", merr(C))
d = (A - C).abs().max(1).values
print("A vs C: max diff", float(d.max()), "at clip", int(d.argmax()), "| A vs F max diff", float((A - F).abs().max()))
print("clips with A != C:", [int(i) for i in torch.nonzero(d > 0).flatten()[:20]])
