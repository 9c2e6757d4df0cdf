"""Throughput of the window ("update") model: windows scored per second (CUDA events), with the split between the
backbone pass (vcg_embed over B*(2w+1) clips) and the post-backbone operators.
Usage: python tools/bench_window.py [B] [window_size] [head_type]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200")); sys.path.insert(0, ROOT)
import torch
from model.fusion import two_stream_window
from model.lang import bert_hugface
from model.vision import resnet50_tsm
from vcg_b200 import synthetic as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1
head = sys.argv[3] if len(sys.argv) > 3 else "cross_attn"
T, L, Wn = 16, 100, 2 * w + 1
sd = W.make_window_state_dict(T, w, head, seed=123)
lang = bert_hugface.BertHugface(pretrain_stage=False)
vis = resnet50_tsm.Resnet50TSM(segments_size=T, shift_div=8, pretrain_stage=False)
model = two_stream_window.TwoStream(lang.base_model, vis.base_model, 768, 2048, T, 128, w)
model.build_chapter_head(2, head); model.load_state_dict(sd, strict=True); model = model.to(0).eval()
g = torch.Generator().manual_seed(1)
img = torch.randn(B, Wn, T, 3, 224, 224, generator=g).cuda()
ids, mask = W.make_text(B * Wn, L, seed=1); ids, mask = ids.view(B, Wn, L).cuda(), mask.view(B, Wn, L).cuda()
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = timed(lambda: model(img, ids, mask, None))
eng = model.engine if hasattr(model, "engine") else model._engine
flat = (img.transpose(0, 1).reshape(Wn * B, T, 3, 224, 224), ids.transpose(0, 1).reshape(Wn * B, L), mask.transpose(0, 1).reshape(Wn * B, L))
ms_embed = timed(lambda: eng.embed(*flat))
# per-video scoring with every clip embedded once (score_video): N = B * Wn clips of one video
N = B * Wn
clips = img.reshape(N, T, 3, 224, 224); cids = ids.reshape(N, L); cmask = mask.reshape(N, L)
ms_video = timed(lambda: model.score_video(clips, cids, cmask))
print(f"score_video: {N} clips of one video, each embedded once: {ms_video:.2f} ms -> {N/ms_video*1e3:.0f} windows/s")
print(f"window model head={head} B={B} w={w} (clips per step {B*Wn}): {ms:.2f} ms/step -> {B/ms*1e3:.1f} windows/s, "
      f"{B*Wn/ms*1e3:.0f} clips/s; backbone pass {ms_embed:.2f} ms ({ms_embed/ms*100:.0f} %), heads + window stack {ms-ms_embed:.2f} ms")
