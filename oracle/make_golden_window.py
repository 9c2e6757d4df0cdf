"""TEST INFRASTRUCTURE — golden fixtures of the reference's window ("update") model.

Run in the build container only (reads /root/reference):   python oracle/make_golden_window.py

Builds the UNMODIFIED two_stream_window.TwoStream exactly like test_video_segment_update.py:90-107, loads
oracle.weights.make_window_state_dict with strict=True (pins the key schema), runs the reference forward on CPU, checks
oracle/window_oracle.py against it (<= 1e-5 relative) and stores the REFERENCE's outputs in tests/golden/window_*.npz.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import two_stream_oracle as orc  # noqa: E402
from oracle import weights as W  # noqa: E402
from oracle import window_oracle as worc  # noqa: E402
from oracle.make_golden import GOLDEN, load_reference, rel  # noqa: E402

# (name, head_type, T, window_size, L, B)
CASES = [("cross_attn_T8_w1_L24_B2", "cross_attn", 8, 1, 24, 2), ("mlp_T8_w1_L24_B2", "mlp", 8, 1, 24, 2),
         # the reference's experimental fusion heads (two_stream_window.py:187-237)
         ("bilinear_T8_w1_L24_B2", "bilinear", 8, 1, 24, 2), ("multiplication_T8_w1_L24_B2", "multiplication", 8, 1, 24, 2),
         ("self_attn_T8_w1_L24_B2", "self_attn", 8, 1, 24, 2)]


def make_inputs(T, window, L, B, seed):
    Wn = 2 * window + 1
    frames = W.make_frames_u8(4 * (B + Wn - 2) + T, seed=seed)
    norm = orc.preprocess_u8(frames)
    # window b = clips starting at 4*(b + i), i = 0..2w  (neighbouring clips overlap, as in the dataset)
    img = torch.stack([orc.gather_clips(norm, [4 * (b + i) for i in range(Wn)], T) for b in range(B)])   # [B,W,T,3,224,224]
    ids, mask = W.make_text(B * Wn, L, seed=seed)
    ids, mask = ids.view(B, Wn, L).clone(), mask.view(B, Wn, L).clone()
    # window 0 sits at the start of its video: its first clip is the dataset's padding (zero frames, zero ids, ZERO mask;
    # infer_youtube_video_dataset.py:488-499)
    img[0, 0] = 0
    ids[0, 0] = 0
    mask[0, 0] = 0
    return img, ids, mask


def main():
    _, bert_hugface, resnet50_tsm, _ = load_reference()
    from model.fusion import two_stream_window  # noqa  (reference module)
    torch.set_grad_enabled(False)
    for name, head_type, T, window, L, B in ([] if os.environ.get('ONLY_VARIANTS') else CASES):
        print(f"== {name}")
        sd = W.make_window_state_dict(T, window, head_type, seed=123)
        lang = bert_hugface.BertHugface(pretrain_stage=False)
        vision = resnet50_tsm.Resnet50TSM(segments_size=T, shift_div=8, pretrain_stage=False)
        model = two_stream_window.TwoStream(lang.base_model, vision.base_model, lang.embed_size, vision.feature_dim, T,
                                            128, window)
        model.build_chapter_head(output_size=2, head_type=head_type)
        print("   load_state_dict(strict=True):", model.load_state_dict(sd, strict=True))
        model = model.eval()
        img, ids, mask = make_inputs(T, window, L, B, seed=77)
        clip_info = {"clip_start_frame": torch.zeros(B, 2 * window + 1, dtype=torch.long),
                     "total_frames": torch.full((B,), 100), "target_clip_idx": torch.full((B,), window),
                     "total_num_clips": torch.full((B,), 20)}
        logits, probs = model(img, ids, mask, clip_info)
        o_logits, o_probs = worc.window_forward(sd, img, ids, mask, T, head_type)
        errs = {"logits": rel(o_logits, logits), "probs": rel(o_probs, probs)}
        print("   oracle vs reference (rel):", errs, "logits", logits.tolist())
        assert max(errs.values()) <= 1e-5, errs
        np.savez_compressed(os.path.join(GOLDEN, f"window_{name}.npz"), logits=logits.numpy(), probs=probs.numpy(),
                            labels=logits.topk(1, 1, True, True)[1].view(-1).numpy(), text_ids=ids.numpy(),
                            meta=np.array([T, window, L, B, 77]))
    # ---- the two unused variants: model/fusion/two_stream_domain_specific.py and model/fusion/window_self_attention.py
    from model.fusion import two_stream_domain_specific, window_self_attention  # noqa  (reference modules)
    name, T, window, L, B = "domain_T8_w1_L24_B2", 8, 1, 24, 2
    print(f"== {name}")
    sd = W.make_domain_state_dict(T, window, seed=123)
    lang = bert_hugface.BertHugface(pretrain_stage=False)
    vision = resnet50_tsm.Resnet50TSM(segments_size=T, shift_div=8, pretrain_stage=False)
    model = two_stream_domain_specific.TwoStream(lang.base_model, vision.base_model, lang.embed_size, vision.feature_dim,
                                                 T, 128, window)
    model.build_chapter_head(output_size=2)
    print("   load_state_dict(strict=True):", model.load_state_dict(sd, strict=True))
    model = model.eval()
    img, ids, mask = make_inputs(T, window, L, B, seed=77)
    logits, probs = model(img, ids, mask, None)
    o_logits, o_probs = worc.domain_specific_forward(sd, img, ids, mask, T)
    errs = {"logits": rel(o_logits, logits), "probs": rel(o_probs, probs)}
    print("   oracle vs reference (rel):", errs, "logits", logits.tolist())
    assert max(errs.values()) <= 1e-5, errs
    np.savez_compressed(os.path.join(GOLDEN, f"window_{name}.npz"), logits=logits.numpy(), probs=probs.numpy(),
                        labels=logits.topk(1, 1, True, True)[1].view(-1).numpy(), text_ids=ids.numpy(),
                        meta=np.array([T, window, L, B, 77]))
    for window in (1, 2):
        name = f"single_block_w{window}_B5"
        print(f"== {name}")
        cfg = type("Config", (), {"hidden_size": 128, "num_attention_heads": 16, "attention_probs_dropout_prob": 0.1,
                                  "window_size": window})
        clf = window_self_attention.VideoChapterClassifier(cfg)
        sd = W.make_single_block_state_dict(window, seed=123)
        print("   load_state_dict(strict=True):", clf.load_state_dict(sd, strict=True))
        clf = clf.eval()
        x = torch.randn(5, 2 * window + 1, 128, generator=torch.Generator().manual_seed(9 + window))
        logits, probs = clf(x, None)
        o_logits, o_probs = worc.single_block_classifier(sd, x)
        errs = {"logits": rel(o_logits, logits), "probs": rel(o_probs, probs)}
        print("   oracle vs reference (rel):", errs, "logits", logits.tolist())
        assert max(errs.values()) <= 1e-5, errs
        np.savez_compressed(os.path.join(GOLDEN, f"window_{name}.npz"), logits=logits.numpy(), probs=probs.numpy(),
                            x=x.numpy(), meta=np.array([window, 5]))
    try:   # the reference's MemoryManager starts a monitoring thread per model
        model.memory_manager.tracker._tracking = False
    except Exception:
        pass


if __name__ == "__main__":
    main()
