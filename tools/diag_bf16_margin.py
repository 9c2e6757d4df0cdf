"""Where does the bf16-mode margin error of the 146-clip golden video come from?  (diagnostic, GPU)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200")); sys.path.insert(0, ROOT)
from vcg_b200 import synthetic as W
from vcg_b200.engine import Engine

g = np.load(os.path.join(ROOT, "tests/golden/video_mlp_T16_L100_600f.npz"))
T, L, B, seed, _, _, n_frames, _ = [int(x) for x in g["meta"]]
sd = W.make_state_dict(T, "mlp", seed=seed)
sd["fusion_head.head.bias"] = torch.from_numpy(g["head_bias"]).clone()
frames, scenes = W.make_video_u8(n_frames, seed=seed)
starts = W.clip_starts(n_frames, T)
ids, mask = W.make_video_text(starts, scenes, T, L, seed=seed)
ref = torch.from_numpy(g["logits"]).double()
mref = ref[:, 1] - ref[:, 0]
fr = frames.cuda(); idc, mkc = ids.cuda(), mask.cuda()
st = torch.tensor(starts, dtype=torch.int32).cuda()
res = {}
embs = {}
for prec in ("fp32", "bf16"):
    eng = Engine(T, "mlp", prec, vision=True, max_tokens=L, max_batch=32)
    eng.load_state_dict(sd)
    # per-clip path via fp32 img? use embed-free route: score, then forward with return_emb on normalised clips is heavy; use u8 path + separate emb engine
    lg, _ = eng.score_video_u8(fr, 0, 4, idc, mkc)
    res[prec] = lg.double().cpu()
    eng.close()
    e2 = Engine(T, "mlp", prec, vision=True, max_tokens=L, max_batch=32, modality="embed")
    e2.load_state_dict(sd)
    ve, le = e2.embed_u8(fr, idc, mkc, clip_start=None, first_start=0, clip_stride=4)
    embs[prec] = (ve.clone(), le.clone())
    e2.close()
for prec in res:
    m = res[prec][:, 1] - res[prec][:, 0]
    print(prec, "logit err", float((res[prec] - ref).abs().max()), "margin err", float((m - mref).abs().max()),
          "common-mode err", float(((res[prec].sum(1) - ref.sum(1)) / 2).abs().max()))
ve32, le32 = embs["fp32"]; ve16, le16 = embs["bf16"]
print("vision emb rel err bf16 vs fp32: max", float((ve16 - ve32).abs().max() / ve32.abs().max()), "rms",
      float((ve16 - ve32).pow(2).mean().sqrt() / ve32.pow(2).mean().sqrt()))
print("lang emb rel err: max", float((le16 - le32).abs().max() / le32.abs().max()), "rms",
      float((le16 - le32).pow(2).mean().sqrt() / le32.pow(2).mean().sqrt()))
# head in fp64 on the host with the four combinations of embeddings
hw = sd["fusion_head.head.weight"].double(); hb = sd["fusion_head.head.bias"].double()
wl = sd["fusion_head.lang_proj_head.weight"].double(); wv = sd["fusion_head.vision_proj_head.weight"].double()
def head(ve, le):
    v = torch.relu(ve.double().cpu() @ wv.T)          # [B,T,128]
    l = torch.relu(le.double().cpu() @ wl.T)          # [B,128]
    x = torch.cat([v, l[:, None, :]], 1).reshape(ve.shape[0], -1)
    return x @ hw.T + hb
for nv, ve in (("v32", ve32), ("v16", ve16)):
    for nl, le in (("l32", le32), ("l16", le16)):
        o = head(ve, le); m = o[:, 1] - o[:, 0]
        print(nv, nl, "fp64 head: margin err", float((m - mref).abs().max()), "logit err", float((o - ref).abs().max()))
