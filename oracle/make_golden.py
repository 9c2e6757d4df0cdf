"""TEST INFRASTRUCTURE — pins the oracle against the UNMODIFIED reference and writes tests/golden/*.npz.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

What it does, per case (head_type x clip_frames x tokens):
  1. builds the reference model exactly like the callers do (test_video_segment_point.py:69-100):
     BertHugface / Resnet50TSM / two_stream.TwoStream / build_chapter_head — with the two constructor patches
     SURVEY.md 8c lists (BertModel.from_pretrained and torchvision resnet50(pretrained=True) need the network);
  2. loads oracle.weights.make_state_dict(...) with load_state_dict(strict=True)  -> pins the key schema;
  3. runs the reference forward (fp32, CPU, .eval(), no_grad) on oracle.weights synthetic inputs;
  4. runs oracle.two_stream_oracle on the same inputs and asserts agreement (<= 1e-5 relative);
  5. stores the REFERENCE's outputs (logits, probs, lang_emb, vision_emb, tap statistics) as the golden fixture.
Also pins the peak-picker / precision-recall restatements on the vectors of SURVEY.md section 4.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/video_chapter_generation"
sys.path.insert(0, ROOT)

from oracle import two_stream_oracle as orc  # noqa: E402
from oracle import weights as W  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# (name, head_type, T, L, B)
CASES = [
    ("mlp_T16_L100_B2", "mlp", 16, 100, 2),
    ("attn_T8_L32_B2", "attn", 8, 32, 2),
]


def load_reference():
    """Import the reference modules with the network-touching constructors patched to random init."""
    import torchvision
    import transformers
    from transformers import BertConfig, BertModel

    def from_pretrained(name, *a, **kw):
        kw.pop("config", None)
        cfg = BertConfig(**{k: v for k, v in kw.items() if k in ("output_attentions",)})
        return BertModel(cfg)

    BertModel.from_pretrained = staticmethod(from_pretrained)
    transformers.BertModel = BertModel
    orig_resnet50 = torchvision.models.resnet50
    torchvision.models.resnet50 = lambda pretrained=False, **kw: orig_resnet50(weights=None)
    sys.path.insert(0, REF)
    from model.fusion import two_stream  # noqa
    from model.lang import bert_hugface  # noqa
    from model.vision import resnet50_tsm  # noqa
    from eval_utils import eval_utils  # noqa
    return two_stream, bert_hugface, resnet50_tsm, eval_utils


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def tap_stats(t):
    """Small, layout-independent summary of an activation tensor (channel-first reference layout)."""
    t = t.detach().float()
    flat = t.reshape(-1)
    return np.array([flat.mean().item(), flat.abs().mean().item(), flat.abs().max().item(),
                     flat.double().pow(2).sum().sqrt().item()], dtype=np.float64)


def main():
    two_stream, bert_hugface, resnet50_tsm, eval_utils = load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_grad_enabled(False)

    for name, head_type, T, L, B in CASES:
        print(f"== {name}")
        sd = W.make_state_dict(clip_frames=T, head_type=head_type, seed=123)
        lang = bert_hugface.BertHugface(pretrain_stage=False)
        vision = resnet50_tsm.Resnet50TSM(segments_size=T, shift_div=8, pretrain_stage=False)
        model = two_stream.TwoStream(lang.base_model, vision.base_model, lang.embed_size, vision.feature_dim, T, 128)
        model.build_chapter_head(output_size=2, head_type=head_type)
        missing = model.load_state_dict(sd, strict=True)
        print("   load_state_dict(strict=True):", missing)
        model = model.eval()

        ids, mask = W.make_text(B, L, seed=123)
        frames = W.make_frames_u8(4 * (B - 1) + T, seed=123)
        starts = [4 * b for b in range(B)]
        img = orc.gather_clips(orc.preprocess_u8(frames), starts, T)        # [B,T,3,224,224] fp32

        # reference taps through forward hooks on the real modules
        ref_taps = {}
        hooks = [model.vision_model.relu.register_forward_hook(
            lambda m, i, o: ref_taps.setdefault("vision.stem", o))]
        for s in range(1, 5):
            hooks.append(getattr(model.vision_model, f"layer{s}").register_forward_hook(
                lambda m, i, o, s=s: ref_taps.__setitem__(f"vision.layer{s}", o)))
        hooks.append(model.lang_model.embeddings.register_forward_hook(
            lambda m, i, o: ref_taps.__setitem__("bert.embeddings", o)))
        for li, layer in enumerate(model.lang_model.encoder.layer):
            hooks.append(layer.register_forward_hook(
                lambda m, i, o, li=li: ref_taps.__setitem__(f"bert.layer{li}", o[0] if isinstance(o, tuple) else o)))

        logits, probs, vis_emb, lang_emb = model(img, ids, mask, return_emb=True)
        for h in hooks:
            h.remove()

        taps = {}
        o_logits, o_probs, o_vis, o_lang = orc.two_stream_forward(sd, img, ids, mask, T, 128, head_type, 8, taps=taps)
        errs = {"logits": rel(o_logits, logits), "probs": rel(o_probs, probs), "vision_emb": rel(o_vis, vis_emb),
                "lang_emb": rel(o_lang, lang_emb)}
        for k in ref_taps:
            errs[k] = rel(taps[k], ref_taps[k])
        print("   oracle vs reference (rel):", {k: f"{v:.2e}" for k, v in errs.items()})
        assert max(errs.values()) <= 1e-5, errs

        # precomputed-embedding configuration (BASELINE.json config 2): Identity vision model, [B,T,2048,1,1] input
        from ops.basic_ops import Identity
        model2 = two_stream.TwoStream(lang.base_model, Identity(), lang.embed_size, vision.feature_dim, T, 128)
        model2.fusion_head = model.fusion_head
        model2 = model2.eval()
        l2, p2 = model2(vis_emb.view(B, T, 2048, 1, 1), ids, mask)
        o2 = orc.two_stream_forward(sd, None, ids, mask, T, 128, head_type, 8, vision_emb=vis_emb)
        assert rel(o2[0], l2) <= 1e-5 and rel(l2, logits) <= 1e-6

        labels = logits.topk(1, 1, True, True)[1].view(-1).tolist()
        assert labels == orc.predict_labels(logits)
        out = {
            "logits": logits.numpy(), "probs": probs.numpy(), "lang_emb": lang_emb.numpy(),
            "vision_emb": vis_emb.numpy(), "labels": np.array(labels), "text_ids": ids.numpy(),
            "attention_mask": mask.numpy(), "clip_starts": np.array(starts),
            "meta": np.array([T, L, B, 123, 8, 128]),   # T, L, B, seed, shift_div, hidden
        }
        for k, v in ref_taps.items():
            out["tap/" + k] = tap_stats(v)
        np.savez_compressed(os.path.join(GOLDEN, f"two_stream_{name}.npz"), **out)
        print("   logits", logits.tolist(), "labels", labels)

    # ---- peak picker / PR known answers, computed with the reference functions
    vec = [1, 0, 0, 0, 1, 1, 0, 0, 1, 1, 1, 1, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]   # eval_utils.py:97
    rng = np.random.RandomState(123)
    extra = [rng.randint(0, 2, size=n).tolist() for n in (1, 2, 7, 50, 146, 896)]
    cases = []
    for arr in [vec, [0, 1, 1], [], [1], [0], [1, 0], [1, 1, 1, 0]] + extra:
        for T in (8, 16, 32):
            ref = eval_utils.convert_clip_label2cut_point(arr, T, 2)
            assert ref == orc.convert_clip_label2cut_point(arr, T, 2)
            cases.append((arr, T, ref))
    assert eval_utils.convert_clip_label2cut_point(vec, 16, 2) == [8, 26, 48, 64]
    pr = eval_utils.calculate_pr([10, 50, 100], [10, 52, 96, 200])
    assert pr == orc.calculate_pr([10, 50, 100], [10, 52, 96, 200])
    np.savez_compressed(
        os.path.join(GOLDEN, "cut_points.npz"),
        labels=np.array([np.array(c[0], dtype=np.int64) for c in cases], dtype=object),
        T=np.array([c[1] for c in cases]), cuts=np.array([np.array(c[2], dtype=np.int64) for c in cases], dtype=object),
        pr=np.array(pr, dtype=np.float64), allow_pickle=True)
    print("cut-point cases:", len(cases), "pr:", pr)


if __name__ == "__main__":
    main()
