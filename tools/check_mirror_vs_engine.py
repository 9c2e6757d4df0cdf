import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200")); sys.path.insert(0, ROOT)
import torch
from vcg_b200 import synthetic as W
from vcg_b200.engine import Engine
from oracle import two_stream_oracle as orc
from model.fusion import two_stream
from model.lang import bert_hugface
from model.vision import resnet50_tsm
T, L = 16, 100
sd = W.make_state_dict(T, "mlp", seed=123)
frames, scenes = W.make_video_u8(600, seed=123)
starts = W.clip_starts(600, T)
ids, mask = W.make_video_text(starts, scenes, T, L, seed=123)
pre = orc.preprocess_u8(frames)
lang = bert_hugface.BertHugface(pretrain_stage=False)
vis = resnet50_tsm.Resnet50TSM(segments_size=T, shift_div=8, pretrain_stage=False)
model = two_stream.TwoStream(lang.base_model, vis.base_model, lang.embed_size, vis.feature_dim, T, 128)
model.build_chapter_head(output_size=2, head_type="mlp")
model.load_state_dict(sd, strict=True)
model = model.to(0).eval(); model.precision = "bf16"
eng = Engine(T, "mlp", "bf16", vision=True, max_tokens=L, max_batch=32)
eng.load_state_dict(sd)
sl = slice(0, 16)
img = orc.gather_clips(pre, starts[sl], T).cuda()
a = model(img, ids[sl].cuda(), mask[sl].cuda(), return_emb=True)
b = eng.forward(img, ids[sl].cuda(), mask[sl].cuda(), return_emb=True)
for name, x, y in zip(("logits", "probs", "vision_emb", "lang_emb"), a, b):
    print(name, "max |diff|", float((x - y).abs().max()))
print("mirror engine:", model.engine.max_tokens, model.engine.max_batch, model.vision_chunk)
