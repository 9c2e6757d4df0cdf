"""TEST INFRASTRUCTURE — VIDEO-SCALE goldens from the UNMODIFIED reference (build container only; reads /root/reference).

    python oracle/make_golden_video.py            # ~4 min of CPU; writes tests/golden/video_*.npz

Two fixtures, both produced by the reference's own modules (model/fusion/two_stream.py TwoStream + BertHugface +
Resnet50TSM, eval_utils.convert_clip_label2cut_point), exactly as the callers drive them:

  video_mlp_T16_L100_600f.npz   BASELINE.json configs[0]: one synthetic 10-minute video (600 uint8 frames ->
      range(0, 600-16, 4) = 146 clips, data/infer_youtube_video_dataset.py:117), T=16, L=100, batch 8
      (19 batches, test_whole_pipeline_per_video.py:145-161 loop), labels by topk (:156-158) and chapter timestamps by
      convert_clip_label2cut_point(labels, 16, 2) (:165-166).
  video_emb_mlp_T16_L100_B256.npz   BASELINE.json configs[1]: the same model on PRECOMPUTED vision embeddings
      (vision_model = Identity, img_clip [B,T,2048,1,1]), one batch of 256 clips.

Random-init weights give every clip of a video nearly the same margin l1 - l0, i.e. one label and no timestamps.  So
that BOTH labels (and real runs of them) occur, the two entries of ``fusion_head.head.bias`` are re-centred: the
common mode of the logits goes to zero and the decision threshold goes to the middle of the WIDEST GAP between sorted
margins in the central 10-90 % of the clips that still yields >= 3 chapter timestamps.  ``raw_logit_absmax`` keeps the
magnitude of the reference's logits BEFORE that shift (the scale the relative tolerances of BASELINE.json refer to: the
shift is one additive constant per logit, applied identically to the reference and to the CUDA path).  The adjusted bias is part of the fixture (``head_bias``); logits are
the reference's with that bias (the head is re-run through the reference's own ``model.fusion_head`` on the reference's
embeddings, and a full reference forward of the first batch confirms it is the same computation bit for bit).
``min_abs_margin`` / ``margin_gap`` record how far every clip is from a label flip: the GPU tests assert the CUDA
path's margin error is well inside it in fp32 AND bf16, which is what makes "identical chapter timestamps" a robust
statement rather than luck.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import two_stream_oracle as orc  # noqa: E402
from oracle import weights as W  # noqa: E402
from oracle.make_golden import GOLDEN, load_reference, rel  # noqa: E402

SEED = 123
CACHE = "/tmp/vcg_video_golden_cache.pt"


def recentre(logits, bias, T, convert):
    """-> (new bias, threshold gap).  See the module docstring.  Among the gaps between neighbouring sorted margins in
    the central 10-90 % of the clips, takes the WIDEST one whose threshold still yields >= 3 chapter timestamps
    (``convert`` = the reference's convert_clip_label2cut_point)."""
    l = logits.double()
    mm = l[:, 1] - l[:, 0]
    m = mm.sort().values
    n = len(m)
    lo, hi = int(0.10 * n), int(0.90 * n)
    gaps = m[lo + 1:hi + 1] - m[lo:hi]
    best = None
    for j in gaps.argsort(descending=True).tolist():
        theta = float((m[lo + j] + m[lo + j + 1]) / 2)
        labels = (mm > theta).long().tolist()
        if len(convert(labels, T, 2)) >= 3:
            best = (j, theta)
            break
    assert best is not None, "no threshold gives three chapter timestamps"
    j, theta = best
    common = float(((l[:, 0] + l[:, 1]) / 2).mean())
    nb = bias.double().clone()
    nb[0] += theta / 2 - common
    nb[1] += -theta / 2 - common
    return nb.float(), float(gaps[j])


def build_model(two_stream, bert_hugface, resnet50_tsm, sd, T, identity_vision=False):
    lang = bert_hugface.BertHugface(pretrain_stage=False)
    if identity_vision:
        from ops.basic_ops import Identity
        model = two_stream.TwoStream(lang.base_model, Identity(), lang.embed_size, 2048, T, 128)
        model.build_chapter_head(output_size=2, head_type="mlp")
        model.load_state_dict({k: v for k, v in sd.items() if not k.startswith("vision_model.")}, strict=True)
    else:
        vision = resnet50_tsm.Resnet50TSM(segments_size=T, shift_div=8, pretrain_stage=False)
        model = two_stream.TwoStream(lang.base_model, vision.base_model, lang.embed_size, vision.feature_dim, T, 128)
        model.build_chapter_head(output_size=2, head_type="mlp")
        model.load_state_dict(sd, strict=True)
    return model.eval()


def labels_of(logits):
    return logits.data.topk(1, 1, True, True)[1].view(-1).tolist()     # test_whole_pipeline_per_video.py:156-158


def main():
    torch.set_grad_enabled(False)
    torch.set_num_threads(os.cpu_count())
    two_stream, bert_hugface, resnet50_tsm, eval_utils = load_reference()
    T, L, n_frames, batch = 16, 100, 600, 8

    # ------------------------------------------------------------------ configs[0]: one 10-minute video, whole path
    sd = W.make_state_dict(T, "mlp", seed=SEED)
    frames, scenes = W.make_video_u8(n_frames, seed=SEED)
    starts = W.clip_starts(n_frames, T)
    B = len(starts)
    ids, mask = W.make_video_text(starts, scenes, T, L, seed=SEED)
    model = build_model(two_stream, bert_hugface, resnet50_tsm, sd, T)
    if os.path.exists(CACHE):
        c = torch.load(CACHE)
        vis_emb, lang_emb, logits0 = c["vis"], c["lang"], c["logits"]
    else:
        pre = orc.preprocess_u8(frames)
        vis_emb, lang_emb, logits0 = [], [], []
        t0 = time.time()
        for b0 in range(0, B, batch):                               # the callers' DataLoader loop, batch 8 (configs[0])
            sl = slice(b0, min(b0 + batch, B))
            img = orc.gather_clips(pre, starts[sl], T)
            lg, _, ve, le = model(img, ids[sl], mask[sl], return_emb=True)
            vis_emb.append(ve); lang_emb.append(le); logits0.append(lg)
            print(f"   reference forward: clips {b0}..{sl.stop} of {B}  ({time.time() - t0:.0f} s)", flush=True)
        vis_emb, lang_emb, logits0 = torch.cat(vis_emb), torch.cat(lang_emb), torch.cat(logits0)
        torch.save({"vis": vis_emb, "lang": lang_emb, "logits": logits0}, CACHE)
    bias, gap = recentre(logits0, sd["fusion_head.head.bias"], T, eval_utils.convert_clip_label2cut_point)
    sd["fusion_head.head.bias"] = bias
    model.fusion_head.head.bias.copy_(bias)
    logits = torch.cat([model.fusion_head(lang_emb[b0:b0 + batch], vis_emb[b0:b0 + batch]) for b0 in range(0, B, batch)])
    probs = torch.softmax(logits, dim=1)                             # two_stream.py:189
    # the head re-run IS the reference forward: check on the first batch through TwoStream.forward itself
    img0 = orc.gather_clips(orc.preprocess_u8(frames[:starts[batch - 1] + T]), starts[:batch], T)
    full0 = model(img0, ids[:batch], mask[:batch])[0]
    assert torch.equal(full0, logits[:batch]), (full0 - logits[:batch]).abs().max()
    labels = labels_of(logits)
    cuts = eval_utils.convert_clip_label2cut_point(labels, T, 2)     # test_whole_pipeline_per_video.py:165-166
    margin = (logits[:, 1] - logits[:, 0]).double()
    # oracle restatement on a sample of clips (the full video through the oracle is what tests/test_oracle.py times)
    o = orc.two_stream_forward(sd, img0, ids[:batch], mask[:batch], T, 128, "mlp", 8)
    assert rel(o[0], full0) <= 1e-5
    assert orc.convert_clip_label2cut_point(labels, T, 2) == cuts
    print("video: labels", "".join(map(str, labels)))
    print("video: cut points", cuts, "| positives", sum(labels), "of", B, "| gap", gap, "| min |margin|",
          float(margin.abs().min()), "| max |logit|", float(logits.abs().max()))
    np.savez_compressed(
        os.path.join(GOLDEN, "video_mlp_T16_L100_600f.npz"),
        logits=logits.numpy(), probs=probs.numpy(), labels=np.array(labels), cut_points=np.array(cuts, dtype=np.int64),
        head_bias=bias.numpy(), clip_starts=np.array(starts), scene_starts=np.array(scenes),
        lang_emb_f16=lang_emb.half().numpy(), vision_emb_clipmean=vis_emb.mean(dim=(1, 2)).numpy(),
        min_abs_margin=float(margin.abs().min()), margin_gap=gap, raw_logit_absmax=float(logits0.abs().max()),
        meta=np.array([T, L, B, SEED, 8, 128, n_frames, batch]))

    # ------------------------------------------------------------------ the same video through the ATTENTION head
    # (two_stream.py:31-48).  The backbones are the same (make_state_dict draws the head last), so the reference's
    # embeddings above are re-used and only the reference's own fusion_head is re-run; the full forward of the first batch
    # confirms it.  The decision bias re-centred is fusion_head.head.proj.bias.
    sd_a = W.make_state_dict(T, "attn", seed=SEED)
    assert all(torch.equal(sd_a[k], v) for k, v in sd.items() if not k.startswith("fusion_head.head"))
    lang_a = bert_hugface.BertHugface(pretrain_stage=False)
    vis_a = resnet50_tsm.Resnet50TSM(segments_size=T, shift_div=8, pretrain_stage=False)
    model_a = two_stream.TwoStream(lang_a.base_model, vis_a.base_model, lang_a.embed_size, vis_a.feature_dim, T, 128)
    model_a.build_chapter_head(output_size=2, head_type="attn")
    model_a.load_state_dict(sd_a, strict=True)
    model_a = model_a.eval()
    logits_a0 = torch.cat([model_a.fusion_head(lang_emb[b0:b0 + batch], vis_emb[b0:b0 + batch]) for b0 in range(0, B, batch)])
    bias_a, gap_a = recentre(logits_a0, sd_a["fusion_head.head.proj.bias"], T, eval_utils.convert_clip_label2cut_point)
    sd_a["fusion_head.head.proj.bias"] = bias_a
    model_a.fusion_head.head.proj.bias.copy_(bias_a)
    logits_a = torch.cat([model_a.fusion_head(lang_emb[b0:b0 + batch], vis_emb[b0:b0 + batch]) for b0 in range(0, B, batch)])
    full_a = model_a(img0, ids[:batch], mask[:batch])[0]
    assert torch.equal(full_a, logits_a[:batch]), (full_a - logits_a[:batch]).abs().max()
    labels_a = labels_of(logits_a)
    cuts_a = eval_utils.convert_clip_label2cut_point(labels_a, T, 2)
    margin_a = (logits_a[:, 1] - logits_a[:, 0]).double()
    o_a = orc.two_stream_forward(sd_a, img0, ids[:batch], mask[:batch], T, 128, "attn", 8)
    assert rel(o_a[0], full_a) <= 1e-5
    print("video (attn head): labels", "".join(map(str, labels_a)))
    print("video (attn head): cut points", cuts_a, "| positives", sum(labels_a), "of", B, "| gap", gap_a, "| min |margin|",
          float(margin_a.abs().min()), "| max |logit|", float(logits_a.abs().max()))
    np.savez_compressed(
        os.path.join(GOLDEN, "video_attn_T16_L100_600f.npz"),
        logits=logits_a.numpy(), probs=torch.softmax(logits_a, dim=1).numpy(), labels=np.array(labels_a),
        cut_points=np.array(cuts_a, dtype=np.int64), head_bias=bias_a.numpy(), clip_starts=np.array(starts),
        scene_starts=np.array(scenes), min_abs_margin=float(margin_a.abs().min()), margin_gap=gap_a,
        raw_logit_absmax=float(logits_a0.abs().max()), meta=np.array([T, L, B, SEED, 8, 128, n_frames, batch]))

    # ------------------------------------------------------------------ configs[1]: precomputed embeddings, B = 256
    Bq = 256
    sd2 = W.make_state_dict(T, "mlp", seed=SEED, include_vision=False)
    emb, ids2, mask2 = W.make_precomputed_inputs(Bq, T, L, seed=SEED)
    model2 = build_model(two_stream, bert_hugface, resnet50_tsm, sd2, T, identity_vision=True)
    lg0 = model2(emb.view(Bq, T, 2048, 1, 1), ids2, mask2)[0]
    bias2, gap2 = recentre(lg0, sd2["fusion_head.head.bias"], T, eval_utils.convert_clip_label2cut_point)
    sd2["fusion_head.head.bias"] = bias2
    model2.fusion_head.head.bias.copy_(bias2)
    logits2, probs2 = model2(emb.view(Bq, T, 2048, 1, 1), ids2, mask2)
    labels2 = labels_of(logits2)
    cuts2 = eval_utils.convert_clip_label2cut_point(labels2, T, 2)
    o2 = orc.two_stream_forward(sd2, None, ids2[:16], mask2[:16], T, 128, "mlp", 8, vision_emb=emb[:16])
    assert rel(o2[0], logits2[:16]) <= 1e-5
    m2 = (logits2[:, 1] - logits2[:, 0]).double()
    print("emb256: positives", sum(labels2), "of", Bq, "| cut points", len(cuts2), "| gap", gap2, "| min |margin|",
          float(m2.abs().min()), "| max |logit|", float(logits2.abs().max()))
    np.savez_compressed(
        os.path.join(GOLDEN, "video_emb_mlp_T16_L100_B256.npz"),
        logits=logits2.numpy(), probs=probs2.numpy(), labels=np.array(labels2), cut_points=np.array(cuts2, dtype=np.int64),
        head_bias=bias2.numpy(), min_abs_margin=float(m2.abs().min()), margin_gap=gap2,
        raw_logit_absmax=float(lg0.abs().max()), meta=np.array([T, L, Bq, SEED, 8, 128]))


if __name__ == "__main__":
    main()
