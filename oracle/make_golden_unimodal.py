"""TEST INFRASTRUCTURE — golden fixtures of the reference's single-modality scorers (--data_mode image / text).

Run in the build container only (reads /root/reference):   python oracle/make_golden_unimodal.py

Per case it builds the UNMODIFIED reference module (Resnet50TSM / Resnet50 / BertHugface with pretrain_stage=False,
test_video_segment_point.py:72-93), loads oracle.weights.make_unimodal_state_dict with strict=True (pins the key
schema), runs the reference forward on CPU, checks the oracle restatement against it (<= 1e-5 relative) and stores the
REFERENCE's outputs under tests/golden/unimodal_*.npz.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import two_stream_oracle as orc  # noqa: E402
from oracle import weights as W  # noqa: E402
from oracle.make_golden import GOLDEN, load_reference, rel  # noqa: E402

# (name, kind, T, L, B)
CASES = [("r50tsm_T8_B2", "r50tsm", 8, 0, 2), ("r50_T8_B2", "r50", 8, 0, 2), ("bert_L48_B3", "bert", 0, 48, 3)]


def main():
    _, bert_hugface, resnet50_tsm, _ = load_reference()
    from model.vision import resnet50  # noqa  (reference module; load_reference() put it on sys.path)
    torch.set_grad_enabled(False)
    for name, kind, T, L, B in CASES:
        print(f"== {name}")
        sd = W.make_unimodal_state_dict(kind, clip_frames=max(T, 1), seed=123)
        if kind == "bert":
            model = bert_hugface.BertHugface(pretrain_stage=False)
            model.build_chapter_head()
            print("   load_state_dict(strict=True):", model.load_state_dict(sd, strict=True))
            model = model.eval()
            ids, mask = W.make_text(B, L, seed=321)
            logits, probs = model(ids, mask)
            o_logits, o_probs, o_emb = orc.text_only_forward(sd, ids, mask)
            out = {"text_ids": ids.numpy(), "attention_mask": mask.numpy(), "emb": o_emb.numpy(),
                   "meta": np.array([T, L, B, 321])}
        else:
            if kind == "r50tsm":
                model = resnet50_tsm.Resnet50TSM(segments_size=T, shift_div=8, pretrain_stage=False)
            else:
                model = resnet50.Resnet50(segments_size=T, pretrain_stage=False)
            model.build_chapter_head()
            print("   load_state_dict(strict=True):", model.load_state_dict(sd, strict=True))
            model = model.eval()
            frames = W.make_frames_u8(4 * (B - 1) + T, seed=321)
            starts = [4 * b for b in range(B)]
            img = orc.gather_clips(orc.preprocess_u8(frames), starts, T)
            logits, probs = model(img)
            o_logits, o_probs, o_emb = orc.vision_only_forward(sd, img, T, 8 if kind == "r50tsm" else 0)
            out = {"clip_starts": np.array(starts), "emb": o_emb.numpy(), "meta": np.array([T, L, B, 321])}
        errs = {"logits": rel(o_logits, logits), "probs": rel(o_probs, probs)}
        print("   oracle vs reference (rel):", errs, "logits", logits.tolist())
        assert max(errs.values()) <= 1e-5, errs
        out.update({"logits": logits.numpy(), "probs": probs.numpy(),
                    "labels": logits.topk(1, 1, True, True)[1].view(-1).numpy()})
        np.savez_compressed(os.path.join(GOLDEN, f"unimodal_{name}.npz"), **out)


if __name__ == "__main__":
    main()
